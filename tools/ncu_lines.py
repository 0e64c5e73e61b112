"""Per-source-line summary of an ncu report (SASS page joined with nvdisasm -g line info).
usage: ncu_lines.py <report.ncu-rep> <cubin-name e.g. diffuse> <mangled-kernel-substring> <source.cu>"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, cub, ksub, srcfile = sys.argv[1:5]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", cub, os.path.join(root, "laris_b200", "liblaris_b200.so")], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(txt) if ".text." in l and ksub in l][0]
line, addr2line = None, {}
for l in txt[start + 1:]:
    if l.startswith("//-----") and addr2line:
        break
    m = re.search(r'//## File "(.*?)", line (\d+)', l)
    if m:
        line = int(m.group(2)) if m.group(1).endswith(os.path.basename(srcfile)) else -int(m.group(2))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
    if m:
        addr2line[int(m.group(1), 16)] = line
kflt = ["--kernel-name", "regex:" + os.environ["NCU_KERNEL"]] if os.environ.get("NCU_KERNEL") else []
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + kflt, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]
ai, ie, smp = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples")
ws = h.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in h else smp
wsi = h.index("L1 Wavefronts Shared Ideal") if "L1 Wavefronts Shared Ideal" in h else smp
per, pers, pw, pwi = (collections.Counter() for _ in range(4))
base, tot, ts, tw = None, 0, 0, 0
for r in rows[hi + 1:]:
    if r and r[0] in ("Kernel Name", "Address"):
        break
    try:
        a = int(r[ai], 16) if r[ai].startswith("0x") else int(r[ai]); n = float(r[ie]); s = float(r[smp] or 0)
        w = float(r[ws] or 0); wi = float(r[wsi] or 0)
    except Exception:
        continue
    if base is None:
        base = a
    ln = addr2line.get(a - base)
    per[ln] += n; pers[ln] += s; pw[ln] += w; pwi[ln] += wi; tot += n; ts += s; tw += w
src = open(srcfile).read().split("\n")
print(f"total warp instr {tot:.3g}  samples {ts:.0f}  shared wavefronts {tw:.3g}")
for ln, v in per.most_common(int(sys.argv[5]) if len(sys.argv) > 5 else 24):
    s = src[ln - 1].strip()[:84] if ln and ln > 0 else f"(inlined/other {ln})"
    print(f"{v/tot*100:5.1f}% instr {pers[ln]/max(ts,1)*100:5.1f}% smp  shW {pw[ln]/1e6:6.2f}M (ideal {pwi[ln]/1e6:5.2f}M)  L{ln}: {s}")

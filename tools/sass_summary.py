"""Blackwell-native instruction counts per kernel of the built library (run where cuobjdump is: the build container).

    python tools/sass_summary.py > profiles/sass_summary.txt

Mnemonics (B200_PROFILING.md): UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor copies
(cp.async.bulk.tensor, incl. tile::gather4), UBLKCP = 1-D cp.async.bulk, UBLKPF = cp.async.bulk.prefetch.L2, LDGSTS = cp.async, SYNCS = mbarrier ops."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "laris_b200", "liblaris_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pats = collections.OrderedDict([("UTC*MMA", r"\bUTC\w*MMA\b"), ("LDTM", r"\bLDTM\b"), ("STTM", r"\bSTTM\b"), ("UTMALDG", r"\bUTMALDG\b"),
                                ("UTMASTG", r"\bUTMASTG\b"), ("UBLKCP", r"\bUBLKCP\b"), ("UBLKPF", r"\bUBLKPF\b"), ("LDGSTS", r"\bLDGSTS\b"),
                                ("SYNCS", r"\bSYNCS\b"), ("HMMA", r"\bHMMA\b"), ("DFMA", r"\bDFMA\b"), ("FFMA", r"\bFFMA\b")])
arch = re.findall(r"arch = (sm_\w+)", out)
print(f"# {os.path.relpath(lib, ROOT)}: {len(out.splitlines())} SASS lines, arch {sorted(set(arch))}")
print("# kernel".ljust(64) + "".join(k.rjust(9) for k in pats))
tot = collections.Counter()
for block in out.split("Function : ")[1:]:
    name = block.split("\n", 1)[0].strip()
    demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0]
    demangled = re.sub(r"^void ", "", demangled).replace("laris::", "")
    counts = {k: len(re.findall(p, block)) for k, p in pats.items()}
    tot.update(counts)
    if any(counts[k] for k in ("UTC*MMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UBLKPF", "LDGSTS", "SYNCS", "HMMA")):
        print(demangled[:62].ljust(64) + "".join(str(counts[k]).rjust(9) for k in pats))
print("TOTAL (all kernels)".ljust(64) + "".join(str(tot[k]).rjust(9) for k in pats))

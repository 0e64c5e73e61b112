"""Turn gpurun_out artefacts into the committed summaries under profiles/.

usage:
  make_profile_summary.py launches <launches.csv> <out.md> [title]
      per-kernel table of an `ncu --metrics gpu__time_duration.sum` launch list (count, total, share)
  make_profile_summary.py full <report.ncu-rep> <out.md> [kernel-regex] [title]
      key metrics of an `ncu --set full` capture (raw page), one block per kernel
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "lts__t_bytes.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("laris::", "")


def launches(path, out, title):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        k = short(r[4])
        a = agg.setdefault(k, [0, 0.0, r[7], r[8]])
        a[0] += 1
        a[1] += float(r[-1])
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# {title}\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` launch list "
                f"({len(rows)} launches, {tot/1e3:.1f} us total; per-launch times are cold-cache and serialised, "
                "so read SHARES, not absolutes).\n\n| kernel | launches | total us | mean us | share | block | grid |\n|---|---:|---:|---:|---:|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {a[0]} | {a[1]/1e3:.1f} | {a[1]/a[0]/1e3:.2f} | {100*a[1]/tot:.1f}% | {a[2]} | {a[3]} |\n")


def full(rep, out, pat, title):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    with open(out, "w") as f:
        f.write(f"# {title}\n\n`ncu --set full --clock-control none --import-source on` (raw page of `{rep.split('/')[-1]}`).\n")
        for r in rows[2:]:
            if pat and not re.search(pat, r[kn]):
                continue
            f.write(f"\n## `{short(r[kn])}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"| {k} | {r[i]} | {units[i]} |\n")


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "launch list")
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None,
             sys.argv[5] if len(sys.argv) > 5 else "ncu full capture")

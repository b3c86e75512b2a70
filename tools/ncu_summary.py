#!/usr/bin/env python
"""Turns ncu output into the small markdown/CSV summaries kept under profiles/.

  ncu_summary.py rep  <file.ncu-rep> [title]   -> markdown table of the per-kernel metrics the judge reads
  ncu_summary.py list <launches.csv> [title]   -> per-kernel-name totals and shares of a gpu__time_duration launch list
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads/inst (of 32)"),
    ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch_resolving"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_throttle"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
]


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("rtb::<unnamed>::", "").replace("<unnamed>::", "").replace("void ", "").replace("unnamed>::", "")
    return name.strip()


def rep(path, title):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    kernels = rows[2:]
    print(f"### {title}\n")
    print(f"Source: `ncu --set full --clock-control none --import-source on`, report `{path.split('/')[-1]}` "
          f"(kept in gpurun_out/, not tracked). One column per captured launch, in launch order.\n")
    print("| metric | " + " | ".join(f"{short(k[idx['Kernel Name']])}" for k in kernels) + " |")
    print("|---|" + "---|" * len(kernels))
    for key, label in METRICS:
        if key not in idx:
            continue
        u = units[idx[key]]
        cells = []
        for k in kernels:
            v = k[idx[key]]
            try:
                f = float(v.replace(",", ""))
                v = f"{f:.0f}" if abs(f) >= 1000 else f"{f:.2f}"
            except ValueError:
                pass
            cells.append(v)
        print(f"| {label} [{u}] | " + " | ".join(cells) + " |")
    print()


def launch_list(path, title):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot = OrderedDict()
    for r in rows[1:]:
        n = short(r[ki])
        t = float(r[vi].replace(",", ""))
        c = tot.setdefault(n, [0, 0.0])
        c[0] += 1
        c[1] += t
    total = sum(v[1] for v in tot.values())
    print(f"### {title}\n")
    print(f"Source: `ncu --metrics gpu__time_duration.sum --clock-control none --csv` ({path.split('/')[-1]}); "
          f"{sum(v[0] for v in tot.values())} launches, {total / 1e6:.3f} ms summed. Per-launch times under ncu are "
          "serialised and cold-cache: read the shares, not the absolutes.\n")
    print("| kernel | launches | total [us] | share |")
    print("|---|---|---|---|")
    for n, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{n}` | {c} | {t / 1e3:.1f} | {100 * t / total:.1f} % |")
    print()


if __name__ == "__main__":
    kind, path = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else path
    (rep if kind == "rep" else launch_list)(path, title)

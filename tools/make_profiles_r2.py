#!/usr/bin/env python
"""Turns raw-page CSV exports of ncu captures (gpurun_out/prof_r2_<workload>.raw.csv, one frame's wavefront launches,
made by tools/r2_profile.sh) into profiles/r2_<workload>.md and the per-workload entries of profiles/issue.json and
profiles/traffic.json that bench.py attaches to its roofline object.

  issue  = warp instructions / s / (SMs x 4 schedulers x SM clock) x threads per instruction / 32
         = the fraction of the machine's thread-instruction slots the traversal launches used — what binds the
           kernel (the BVH is L2-resident, DRAM carries a tenth of the algorithmic bytes).
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
SMS, SCHED = 148, 4

ROWS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads/inst (of 32)"),
    ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("sm__cycles_active.avg", "SM active cycles"),
    ("sm__cycles_elapsed.avg", "SM elapsed cycles"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}


def short(name):
    for k in ("k_wf_traverse", "k_wf_shade", "k_wf_generate", "k_wf_resolve", "k_skin", "k_refit", "k_refresh", "k_tlas"):
        if k in name:
            return k
    return name.split("(")[0][-24:]


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, units, ks = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}

    def val(k, m, scale=True):
        if m not in idx or k[idx[m]] in ("", "n/a"):
            return None
        v = float(k[idx[m]].replace(",", ""))
        return v * UNIT.get(units[idx[m]], 1.0) if scale else v
    return idx, units, ks, val


def main():
    tags = sys.argv[1:] or ["K3", "K3headline", "K3glass", "K2", "K4", "K5"]
    issue_path, traffic_path = os.path.join(P, "issue.json"), os.path.join(P, "traffic.json")
    issue = json.load(open(issue_path)) if os.path.isfile(issue_path) else {}
    traffic = json.load(open(traffic_path)) if os.path.isfile(traffic_path) else {}
    for tag in tags:
        path = os.path.join(G, f"prof_r2_{tag}.raw.csv")
        if not os.path.isfile(path):
            print("missing", path)
            continue
        idx, units, ks, val = load(path)
        names = [short(k[idx["Kernel Name"]]) for k in ks]
        trav = [k for k, n in zip(ks, names) if n == "k_wf_traverse"]
        t = sum(val(k, "gpu__time_duration.sum") for k in trav)
        winst = sum(val(k, "smsp__inst_executed.sum") for k in trav)
        tinst = sum(val(k, "smsp__inst_executed.sum") * val(k, "smsp__thread_inst_executed_per_inst_executed.ratio") for k in trav)
        # SM clock during the capture: elapsed cycles / duration of the same launches
        clock = sum(val(k, "sm__cycles_elapsed.avg") for k in trav) / t
        slots = SMS * SCHED * clock * t
        dram = [val(k, "dram__bytes_read.sum") + val(k, "dram__bytes_write.sum") for k in trav]
        issue[tag] = {
            "kernel": "k_wf_traverse", "launches_captured": len(trav), "time_ms": round(t * 1e3, 3),
            "warp_instructions": int(winst), "threads_per_instruction": round(tinst / winst, 2),
            "issue_slot_frac": round(winst / slots, 4), "thread_instruction_frac": round(tinst / (32 * slots), 4),
            "sm_clock_mhz": round(clock / 1e6, 1),
            "unit": "fraction of SMs x 4 schedulers x SM clock x 32 lanes",
            "source": f"profiles/r2_{tag}.md (ncu, one frame of the workload, k_wf_traverse launches summed)"}
        traffic[tag] = {"kernel": "k_wf_traverse", "dram_bytes_per_launch": round(sum(dram) / len(dram)),
                        "launches_captured": len(trav),
                        "source": f"profiles/r2_{tag}.md (ncu, dram__bytes_read.sum + dram__bytes_write.sum, mean over the "
                                  "k_wf_traverse launches of one frame)"}
        with open(os.path.join(P, f"r2_{tag}.md"), "w") as f:
            f.write(f"# Round 2 — ncu capture of one frame of workload {tag} (wavefront launches in launch order)\n\n")
            f.write("Source: `tools/r2_profile.sh` (`ncu --clock-control none`, bench.py's frame loop, pipeline_lanes=1 so that "
                    "launches do not overlap under the profiler); raw-page CSV exported on the GPU box. Times under ncu are "
                    "serialised and cold-cache: read shares and ratios.\n\n")
            f.write("| metric | " + " | ".join(names) + " |\n|---|" + "---|" * len(names) + "\n")
            for m, label in ROWS:
                if m not in idx:
                    continue
                cells = []
                for k in ks:
                    v = k[idx[m]]
                    try:
                        x = float(v.replace(",", ""))
                        cells.append(f"{x:.3g}" if abs(x) < 1e6 else f"{x:.4g}")
                    except ValueError:
                        cells.append(v)
                f.write(f"| {label} [{units[idx[m]]}] | " + " | ".join(cells) + " |\n")
            i = issue[tag]
            f.write(f"\n**k_wf_traverse over the frame:** {len(trav)} launches, {i['time_ms']} ms, {i['warp_instructions'] / 1e9:.2f} G warp "
                    f"instructions at {i['threads_per_instruction']} threads per instruction = **{100 * i['issue_slot_frac']:.1f} % of the issue "
                    f"slots x {i['threads_per_instruction']}/32 lanes = {100 * i['thread_instruction_frac']:.1f} % of the thread-instruction peak** "
                    f"(148 SMs x 4 schedulers at {i['sm_clock_mhz']} MHz); DRAM {sum(dram) / 1e9:.2f} GB over those launches "
                    f"= {sum(dram) / t / 1e9:.0f} GB/s.\n")
        print(tag, json.dumps(issue[tag]))
    json.dump(issue, open(issue_path, "w"), indent=1)
    json.dump(traffic, open(traffic_path, "w"), indent=1)


if __name__ == "__main__":
    main()

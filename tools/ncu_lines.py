#!/usr/bin/env python
"""Per-source-line view of one kernel launch in an .ncu-rep: warp instructions, threads per instruction and stall
samples, aggregated from ncu's source page (needs -lineinfo and --import-source on).
  ncu_lines.py <rep> <kernel-regex> <launch-skip> [top]"""
import csv, io, subprocess, sys, os, re

def load(rep, kernel, skip):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                          f"regex:{kernel}", "--launch-skip", str(skip), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    out = []  # (file, line, text, inst, threads, samples)
    f = None; hdr = None; cur = None
    for r in rows:
        if not r: continue
        if r[0] == "File Path": f = os.path.basename(r[1]); continue
        if r[0] == "Function Name": continue
        if r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; idx_src = 1; continue
        if r[0] != "":
            cur = [f, int(r[0]), r[1], 0, 0, 0, 0]; out.append(cur)
        elif cur is not None:
            try:
                cur[3] += int(r[hdr["Instructions Executed"]]); cur[4] += int(r[hdr["Thread Instructions Executed"]])
                cur[5] += int(r[hdr["# Samples"]]); cur[6] += 1
            except (ValueError, KeyError):
                pass
    return out

if __name__ == "__main__":
    rep, kernel, skip = sys.argv[1], sys.argv[2], int(sys.argv[3])
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    lines = load(rep, kernel, skip)
    ti = sum(l[3] for l in lines); tt = sum(l[4] for l in lines); ts = sum(l[5] for l in lines)
    print(f"total warp inst {ti}, threads/inst {tt / max(1, ti):.2f}, samples {ts}")
    print(f"{'file:line':28s} {'inst%':>6s} {'thr/inst':>8s} {'smp%':>6s} {'sass':>5s}  source")
    for l in sorted(lines, key=lambda l: -l[3])[:top]:
        print(f"{l[0][:20] + ':' + str(l[1]):28s} {100 * l[3] / ti:6.2f} {l[4] / max(1, l[3]):8.2f} {100 * l[5] / max(1, ts):6.2f} {l[6]:5d}  {l[2].strip()[:110]}")

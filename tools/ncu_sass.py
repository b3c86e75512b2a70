#!/usr/bin/env python
"""SASS listing of one kernel launch with executed counts: ncu_sass.py <rep> <kernel-regex> <launch-skip> > out.txt
columns: address  warp-instructions  threads/inst  stall-samples  instruction"""
import csv, io, subprocess, sys
rep, kernel, skip = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name",
                      f"regex:{kernel}", "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = {h: i for i, h in enumerate(rows[1])}
rows = [rows[0], rows[1]] + [r for r in rows[2:] if len(r) > hdr["Thread Instructions Executed"] and r[hdr["Instructions Executed"]].isdigit()]
tot = sum(int(r[hdr["Instructions Executed"]]) for r in rows[2:])
print(f"# {rows[0][1]}  total warp instructions {tot}")
for r in rows[2:]:
    ie, te, sm = int(r[hdr["Instructions Executed"]]), int(r[hdr["Thread Instructions Executed"]]), int(r[hdr["# Samples"]])
    print(f"{r[0][-5:]} {ie:10d} {100*ie/tot:5.2f}% {te / max(1, ie):5.1f} {sm:5d}  {r[1]}")

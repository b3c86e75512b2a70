"""Aggregate an `ncu --page source --csv` export (per SASS instruction) by source line.

The CSV export has no line column, so the lines come from `nvdisasm -g -c` of the same cubin (addresses match):
    cuobjdump -xelf all wavefront_shade.o; nvdisasm -g -c *.cubin > shade.dis
    python tools/ncu_source_lines.py gpurun_out/src_r2z_shade.sass.csv shade.dis [kernel-index] [top]
Prints, per file:line, warp instructions executed, thread instructions, stall samples and the dominant stall reasons."""
import csv, re, sys, collections

def load_dis(path):
    funcs, cur, line = {}, None, None
    for raw in open(path, errors="replace"):
        m = re.match(r"\s*\.section\s+\.text\.(\S+),", raw)
        if m:
            cur = funcs.setdefault(m.group(1), {}); line = None; continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', raw)
        if m:
            line = (m.group(1).split("/")[-1], int(m.group(2)), m.group(3).strip()); continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", raw)
        if m and cur is not None:
            cur[int(m.group(1), 16)] = (line, m.group(2).strip())
    return funcs

def load_csv(path):
    kernels, rows, header, name = [], None, None, None
    for rec in csv.reader(open(path, errors="replace")):
        if not rec: continue
        if rec[0] == "Kernel Name":
            if rows is not None: kernels.append((name, header, rows))
            name, rows, header = rec[1], [], None; continue
        if rec[0] == "Address":
            header = rec; continue
        if rows is not None and header is not None: rows.append(rec)
    if rows is not None: kernels.append((name, header, rows))
    return kernels

def main():
    csv_path, dis_path = sys.argv[1], sys.argv[2]
    which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    name, header, rows = load_csv(csv_path)[which]
    col = {h: i for i, h in enumerate(header)}
    mangled_hint = re.sub(r"[^A-Za-z0-9_]", "", name.split("(")[0].split("::")[-1].split("<")[0])
    funcs = load_dis(dis_path)
    # pick the function whose instruction count matches
    cands = [(abs(len(v) - len(rows)), k) for k, v in funcs.items() if mangled_hint in k]
    cands.sort()
    fn = funcs[cands[0][1]]
    print("#", name); print("# matched", cands[0][1][-60:], "instructions", len(fn), "csv rows", len(rows))
    base = int(rows[0][col["Address"]], 16) if rows[0][col["Address"]].startswith("0x") else int(rows[0][col["Address"]])
    stalls = [h for h in header if h.startswith("stall_") and "Not Issued" not in h]
    agg = collections.defaultdict(lambda: collections.Counter())
    tot = collections.Counter()
    for r in rows:
        a = r[col["Address"]]
        addr = (int(a, 16) if a.startswith("0x") else int(a)) - base
        line, text = fn.get(addr, (None, "?"))
        key = (line[0], line[1]) if line else ("?", 0)
        def num(h):
            try: return float(r[col[h]] or 0)
            except ValueError: return 0.0
        c = agg[key]
        c["inst"] += num("Instructions Executed"); c["thr"] += num("Thread Instructions Executed")
        c["samples"] += num("Warp Stall Sampling (All Samples)"); c["n"] += 1
        c["local"] += num("L2 Theoretical Sectors Local")
        for s in stalls: c[s] += num(s)
        tot["inst"] += num("Instructions Executed"); tot["thr"] += num("Thread Instructions Executed"); tot["samples"] += num("Warp Stall Sampling (All Samples)")
    print("# total warp inst %.4g thread inst %.4g (%.1f thr/inst) samples %d" % (tot["inst"], tot["thr"], tot["thr"] / max(tot["inst"], 1), tot["samples"]))
    byfile = collections.Counter()
    for (f, l), c in agg.items(): byfile[f] += c["samples"]
    print("# samples by file:", {k: "%.1f%%" % (100 * v / max(tot["samples"], 1)) for k, v in byfile.most_common()})
    print("%-28s %6s %8s %8s %6s  %s" % ("line", "sass", "inst%", "samples%", "thr/i", "top stalls"))
    for key, c in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        st = sorted(((c[s], s[6:]) for s in stalls), reverse=True)[:3]
        print("%-28s %6d %7.2f%% %7.2f%% %6.1f  %s" % ("%s:%d" % key, c["n"], 100 * c["inst"] / max(tot["inst"], 1), 100 * c["samples"] / max(tot["samples"], 1),
              c["thr"] / max(c["inst"], 1), ", ".join("%s %.0f%%" % (n, 100 * v / max(c["samples"], 1)) for v, n in st)))

if __name__ == "__main__":
    main()

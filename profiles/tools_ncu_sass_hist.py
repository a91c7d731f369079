"""Opcode histogram (executed warp instructions) of one kernel from an ncu source-page csv:
  ncu -i X.ncu-rep --page source --print-source cuda,sass --csv -k regex:<kernel> > lines.csv
  python profiles/tools_ncu_sass_hist.py lines.csv <units per launch> [first_line last_line]
With a line range, lists every SASS instruction attributed to those source lines instead."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
lo, hi = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (None, None)
hdr, cur_line, cur_file = None, None, None
ops, per_line, seen = collections.Counter(), [], set()
for r in rows:
    if not r:
        continue
    if r[0] == "Function Name":
        continue
    if r[0] in ("File Path", "File Name"):
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    if r[0].isdigit():
        cur_line = int(r[0])
        continue
    if r[0] == "" and r[2].startswith("0x") and r[2] not in seen:  # the csv lists the kernel once per source file view
        seen.add(r[2])
        try:
            inst = int(r[hdr.index("Instructions Executed")])
        except ValueError:
            continue
        text = r[3].strip()
        toks = text.split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        ops[op.split(".")[0]] += inst
        per_line.append((cur_file, cur_line, text, inst, int(r[hdr.index("# Samples")] or 0)))
tot = sum(ops.values())
print("total %d warp instructions, %.1f per unit" % (tot, tot / units))
if lo is None:
    for op, c in ops.most_common(45):
        print("%-12s %8.1f/unit %5.1f%%" % (op, c / units, 100.0 * c / tot))
else:
    for f, l, t, i, s in per_line:
        if f and f.startswith("sot_pairs") and lo <= (l or 0) <= hi:
            print("%4d %7.2f/unit samp %5d  %s" % (l, i / units, s, t[:110]))

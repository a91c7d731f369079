"""Per-source-line instruction counts and stall samples from an ncu report:
  ncu -i X.ncu-rep --page source --print-source cuda,sass --csv -k regex:<kernel> > lines.csv
  python profiles/tools_ncu_lines.py lines.csv <units per launch> [min share]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.004
cur, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] in ("File Path", "File Name"):
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr and r[0].isdigit() and len(r) == len(hdr) and r[2] == "-":
        try:
            inst, samp = int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("# Samples")])
        except ValueError:
            continue
        st = {h: int(r[i]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h and r[i].isdigit()}
        top = sorted(st.items(), key=lambda x: -x[1])[:2]
        out.append((cur, int(r[0]), r[1].strip()[:88], inst, samp, " ".join("%s:%d" % (k[6:], v) for k, v in top if v)))
ti, ts = sum(o[3] for o in out), sum(o[4] for o in out)
print("total warp instructions %d (%.1f per unit), samples %d" % (ti, ti / units, ts))
for f, l, s, i, sm, top in out:
    if i / ti > min_share or sm / ts > min_share * 1.5:
        print("%-14s:%4d  %7.1f/unit %5.1f%%  samp %5.1f%%  [%s]  %s" % (f[:14], l, i / units, 100 * i / ti, 100 * sm / ts, top, s))

"""Development aid: executed warp instructions and stall samples per source line of one ncu capture
(ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > X.csv; python tools/ncu_lines.py X.csv [top])."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg, fname, cur, ie, isamp = {}, None, None, None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        ie, isamp = r.index("Instructions Executed"), r.index("# Samples")
        continue
    if ie is None or len(r) <= ie:
        continue
    if r[0] != "":
        cur = (fname, int(r[0]), r[1])
        continue
    a = agg.setdefault(cur, [0, 0])
    try:
        a[0] += int(r[ie]); a[1] += int(r[isamp])
    except ValueError:
        pass
tot = sum(a[0] for a in agg.values()); tots = sum(a[1] for a in agg.values())
print("total inst", tot, "samples", tots)
for k, a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print("%5.1f%% inst %5.1f%% samples  %s:%d  %s" % (100 * a[0] / tot, 100 * a[1] / max(tots, 1), k[0], k[1], k[2].strip()[:100]))

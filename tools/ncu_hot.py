"""Developer tool: print the hottest SASS lines (warp-stall samples) of an ncu report's source page.
    ncu -i rep.ncu-rep --page source --csv > src.csv ; python tools/ncu_hot.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20
hdr = rows[1]
c = hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows[2:] if len(r) > c]
tot = sum(float(r[c] or 0) for r in data) or 1.0
print("total samples", tot)
for r in sorted(data, key=lambda r: -float(r[c] or 0))[:N]:
    st = sorted(((float(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
    print("%5.1f%%  %-70s %s" % (100 * float(r[c]) / tot, r[1].strip()[:70],
                                 " ".join("%s=%d" % (n[6:], v) for v, n in st if v > 0)))

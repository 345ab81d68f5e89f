"""Hot SASS instructions of an ncu --page source --csv dump: python tools/ncu_hot.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ci = hdr.index("# Samples"); si = hdr.index("Source"); ei = hdr.index("Instructions Executed")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for k, r in enumerate(rows[2:]):
    try:
        n = int(r[ci])
    except Exception:
        continue
    top = sorted(((int(r[i] or 0), hdr[i]) for i in stall), reverse=True)[:2]
    data.append((n, k, r[si].strip()[:90], r[ei], top))
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
for d in sorted(data, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{d[0]:6d} {100*d[0]/tot:5.1f}% #{d[1]:5d} x{d[3]:>8} {d[2]:90s} {d[4]}")

"""Hot SASS instructions of an ncu report's source page (csv): python tools_sass_hot.py src.csv [min_pct]
Prints every instruction holding >= min_pct of the warp-state samples with its top stall reasons, in program order."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
minp = float(sys.argv[2]) if len(sys.argv) > 2 else 0.7
hdr = rows[1]; data = rows[2:]
iS = hdr.index('Source'); iN = hdr.index('# Samples'); iA = hdr.index('Address') if 'Address' in hdr else None
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[iN] or 0) for r in data)
for k, r in enumerate(data):
    n = int(r[iN] or 0)
    if n * 100.0 / max(tot, 1) >= minp:
        st = sorted(((int(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:3]
        print(f"{k:5d} {100.0 * n / tot:5.1f}%  {r[iS].strip()[:70]:70s} {[(h, v) for v, h in st if v]}")
print("total samples", tot)

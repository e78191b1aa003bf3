"""Reads an `ncu --page source --csv` export and prints the SASS lines with the most stall samples
together with the dominant stall reasons (developer tool; not part of the product path)."""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
hdr = rows[1]
col = {n: i for i, n in enumerate(hdr)}
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
data = []
total = 0
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    if not r[col["# Samples"]].isdigit():
        continue
    n = int(r[col["# Samples"]] or 0)
    total += n
    data.append((n, r))
data.sort(key=lambda t: -t[0])
print("total samples", total)
agg = {n: 0 for n in stall_cols}
for n, r in data:
    for s in stall_cols:
        agg[s] += int(r[col[s]] or 0)
print("by reason:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
for n, r in data[:top]:
    reasons = sorted(((int(r[col[s]] or 0), s[6:]) for s in stall_cols), reverse=True)[:3]
    print("%6d %5.1f%%  %-70s %s" % (n, 100.0 * n / max(1, total), r[col["Source"]].strip()[:70], reasons))

"""Top stall-sample instructions of one kernel from `ncu --page source --csv` output (first kernel instance)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = his[0]
end = his[1] - 1 if len(his) > 1 else len(rows)
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
S = idx['# Samples']
tot = sum(int(r[S]) for r in data)
print("instructions", len(data), "samples", tot)
for r in sorted(data, key=lambda r: -int(r[S]))[:n]:
    print(f"{int(r[S]):6d} {100*int(r[S])/max(tot,1):5.1f}%  exec={r[idx['Instructions Executed']]:>8s}  {r[idx['Source']].strip()[:100]}")

"""Condense `ncu --page source --csv --print-source cuda` (per CUDA line) or sass output: the lines with the most warp-stall samples."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
# find header row (contains 'Source' and a sampling column)
hi = next(i for i, r in enumerate(rows) if any("Sampling" in c for c in r))
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
samp = next(h for h in hdr if "Warp Stall Sampling (All" in h) if any("Warp Stall Sampling (All" in h for h in hdr) else next(h for h in hdr if "Sampling" in h)
src = "Source" if "Source" in col else hdr[1]
data = []
for r in rows[hi + 1:]:
    if len(r) != len(hdr):
        continue
    try:
        n = float(r[col[samp]].replace(",", "") or 0)
    except ValueError:
        continue
    data.append((n, r))
tot = sum(n for n, _ in data) or 1
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print("total samples", tot, "column:", samp)
keep = [h for h in hdr if h.startswith("stall_") or "Instructions Executed" == h][:0]
for n, r in sorted(data, key=lambda t: -t[0])[:top]:
    line = r[col["#"]] if "#" in col else ""
    print(f"{100 * n / tot:5.1f}%  {line:>5s}  {r[col[src]].strip()[:150]}")

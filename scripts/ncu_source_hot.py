"""Condense `ncu --page source --csv --print-source sass`: SASS instructions with the most warp-stall samples, the dominant stall
reason of each, and totals per opcode class.  usage: ncu_source_hot.py file.csv[.gz] [top]"""
import csv, gzip, sys, collections
fn = sys.argv[1]
f = gzip.open(fn, "rt", errors="replace") if fn.endswith(".gz") else open(fn, errors="replace")
rows = list(csv.reader(f))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    n = float(r[col["Warp Stall Sampling (All Samples)"]] or 0)
    st = {h: float(r[col[h]] or 0) for h in stall_cols}
    data.append((n, r[col["Source"]].strip(), st, float(r[col["Instructions Executed"]] or 0)))
tot = sum(d[0] for d in data) or 1
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
print(rows[0][1][:100], "| samples", int(tot), "| SASS instructions", len(data))
agg = collections.Counter()
for n, src, st, ex in data:
    for h, v in st.items():
        agg[h] += v
print("stall reasons:", " ".join(f"{h[6:]}={100 * v / tot:.0f}%" for h, v in agg.most_common(8)))
ops = collections.Counter(); opx = collections.Counter()
for n, src, st, ex in data:
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.split(".")[0]
    ops[op] += n; opx[op] += ex
print("by opcode (samples% / warp-instructions):", " ".join(f"{o}={100 * v / tot:.0f}%/{int(opx[o])}" for o, v in ops.most_common(12)))
for i, (n, src, st, ex) in enumerate(data):
    pass
order = sorted(range(len(data)), key=lambda i: -data[i][0])[:top]
for i in sorted(order):
    n, src, st, ex = data[i]
    dom = max(st.items(), key=lambda kv: kv[1])
    print(f"{100 * n / tot:5.1f}%  #{i:5d} x{int(ex):8d}  {dom[0][6:]:14s} {src[:110]}")

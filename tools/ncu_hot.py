"""python tools/ncu_hot.py source_page.csv [N]: hottest SASS instructions of an `ncu --page source --csv` export with their
dominant stall reasons, plus totals per stall reason."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
tot = {h: 0 for h in stall_cols}
total_samples = 0
total_inst = 0
for idx, r in enumerate(rows[hi + 1:]):
    if len(r) < len(hdr) or r[0] == "Address": continue
    s = int(r[col["# Samples"]] or 0)
    total_samples += s
    total_inst += int(r[col["Instructions Executed"]] or 0)
    st = {h: int(r[col[h]] or 0) for h in stall_cols}
    for h in stall_cols: tot[h] += st[h]
    data.append((s, idx, r[col["Source"]], st, int(r[col["Instructions Executed"]] or 0)))
print("total samples", total_samples, "warp-instructions", total_inst)
print("stall totals:", ", ".join(f"{h[6:]} {100*v/max(1,total_samples):.1f}%" for h, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v))
for s, idx, src, st, ie in sorted(data, reverse=True)[:n]:
    top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"{100*s/total_samples:5.1f}% #{idx:5d} x{ie:8d} {src[:90]:90s} {top[0][0][6:]}={top[0][1]} {top[1][0][6:]}={top[1][1]}")

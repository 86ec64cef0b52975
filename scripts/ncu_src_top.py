"""Top stall sites from `ncu -i X.ncu-rep --page source --csv` output (SASS view).
usage: ncu -i rep --page source --csv > src.csv; python scripts/ncu_src_top.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[hi + 1:] if len(r) >= len(hdr) - 2 and r[0].startswith("0x")]
key = ci["Warp Stall Sampling (All Samples)"]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(float(r[key] or 0) for r in body)
print("total samples", tot, "instructions", len(body))
agg = {s: sum(float(r[ci[s]] or 0) for r in body) for s in stalls}
print("by reason:", ", ".join(f"{k[6:]}={v/tot*100:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0))
for idx, r in sorted(enumerate(body), key=lambda ir: -float(ir[1][key] or 0))[:n]:
    why = sorted(((float(r[ci[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
    print(f"{float(r[key])/tot*100:5.1f}%  #{idx:5d} {r[ci['Source']].strip()[:90]:90s} {why[0][1]}:{why[0][0]:.0f} {why[1][1]}:{why[1][0]:.0f}")

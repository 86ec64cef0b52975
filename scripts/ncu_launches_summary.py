"""Per-kernel table for one training step from an ncu --csv launch list.
usage: python scripts/ncu_launches_summary.py gpurun_out/launches.csv [out.json]"""
import collections, csv, json, re, sys

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
recs = collections.OrderedDict()
for row in csv.DictReader(lines):
    d = recs.setdefault(int(row["ID"]), {"name": row["Kernel Name"]})
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except ValueError:
        v = 0.0
    d[row["Metric Name"]] = (v, row["Metric Unit"])
ids = sorted(recs)
pat = [k for k, i in enumerate(ids) if "patchify" in recs[i]["name"]]
sel = ids[pat[0]:pat[1]] if len(pat) >= 2 else ids
T = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
B = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
agg = collections.defaultdict(lambda: dict(n=0, us=0.0, rd=0.0, wr=0.0, tens=0.0))
for i in sel:
    d = recs[i]
    nm = re.sub(r"\(.*", "", d["name"]).replace("void ", "")[:78]
    a = agg[nm]
    t = d["gpu__time_duration.sum"]; us = t[0] * T.get(t[1], 1.0)
    a["n"] += 1; a["us"] += us
    for k, key in (("rd", "dram__bytes_read.sum"), ("wr", "dram__bytes_write.sum")):
        if key in d:
            a[k] += d[key][0] * B.get(d[key][1], 1)
    tp = d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
    if tp:
        a["tens"] += tp[0] * us
tot = sum(a["us"] for a in agg.values())
print(f"one step (patchify -> next patchify): {len(sel)} launches, {tot / 1e3:.2f} ms serialized (ncu replays each kernel alone, cold L2)")
print(f"{'kernel':78s} {'n':>4s} {'ms':>7s} {'%':>5s} {'avg us':>8s} {'rd GB':>7s} {'wr GB':>7s} {'GB/s':>6s} {'tensor%':>7s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"])[:24]:
    print(f"{k:78s} {a['n']:4d} {a['us'] / 1e3:7.3f} {100 * a['us'] / tot:5.1f} {a['us'] / a['n']:8.1f} {a['rd'] / 1e9:7.2f} {a['wr'] / 1e9:7.2f} "
          f"{(a['rd'] + a['wr']) / a['us'] / 1e3:6.0f} {a['tens'] / a['us'] if a['us'] else 0:7.1f}")
g = [a for k, a in agg.items() if "gemm" in k]
n = sum(a["n"] for a in g); by = sum(a["rd"] + a["wr"] for a in g); us = sum(a["us"] for a in g)
print(f"GEMM family: {n} launches, {us / 1e3:.2f} ms = {100 * us / tot:.1f}% of the step, DRAM {by / 1e9:.1f} GB/step = {by / n / 1e6:.1f} MB/launch")
if len(sys.argv) > 2:
    json.dump({"source": sys.argv[1].split("/")[-1] + " (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over one training step)",
               "gemm_launches_per_step": n, "dram_bytes_per_launch": by / n, "dram_bytes_per_step_gemm": by,
               "serialized_ms_gemm": us / 1e3, "gemm_share_of_step_serialized": us / tot}, open(sys.argv[2], "w"), indent=1)

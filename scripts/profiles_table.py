"""Markdown table of the committed bench lines (profiles/r02_bench_*.json)."""
import glob, json, os, sys
rows = []
for f in sorted(glob.glob("profiles/r02_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception:
        continue
    ex = d.get("extras") or {}
    cb = d.get("cpu_baseline") or {}
    rows.append((d["config"].get("name", "?"), d["n_gpus"], os.path.basename(f), d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"],
                 d["attn_mlp_frac_of_peak"], d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"].get("share_of_step") or 0,
                 d["host_enqueue_ms_per_step"], d["gpu_launches"] / max(1, d["steps"]), cb.get("value"), cb.get("kind"), ex))
print("| config | GPUs | file | value | ms/step | e2e | attn+MLP frac of peak | GEMM TFLOP/s (frac, share of step) | host enqueue ms | launches/step | CPU arm (kind) |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for r in rows:
    cpu = f"{r[13]:.3g} ({r[14]})" if r[13] else "—"
    print(f"| {r[0]} | {r[1]} | `{r[2]}` | {r[3]:.1f} {r[4]} | {r[5]:.2f} | {r[6]:.1f} | {r[7]:.3f} | {r[8]:.0f} ({r[9]:.2f}, {r[10]:.2f}) | {r[11]:.2f} | {r[12]:.0f} | {cpu} |")
print()
for r in rows:
    if r[15]:
        print(f"* `{r[2]}` extras: " + ", ".join(f"{k} = {v:.4g}" if isinstance(v, float) else f"{k}: {v}" for k, v in r[15].items() if k != "decoder_note"))

"""Text summary of `ncu -i X.ncu-rep --page details --csv` files for profiles/ (one block per kernel capture).
usage: python scripts/ncu_details_summary.py gpurun_out/r02_*.details.csv > profiles/r02_ncu_set_full_summary.txt"""
import csv, sys
KEEP = ("Duration", "Elapsed Cycles", "SM Frequency", "DRAM Throughput", "Memory Throughput", "Compute (SM) Throughput",
        "Registers Per Thread", "Dynamic Shared Memory Per Block", "Block Size", "Grid Size", "Achieved Occupancy",
        "Executed Ipc Active", "Issue Slots Busy", "No Eligible", "Eligible Warps Per Scheduler", "L2 Cache Throughput",
        "L1/TEX Cache Throughput", "Mem Busy", "Max Bandwidth", "Local Load", "Local Store", "Theoretical Occupancy")
for path in sys.argv[1:]:
    rows = list(csv.DictReader(l for l in open(path, errors="replace") if l.startswith('"')))
    if not rows:
        continue
    print(f"== {path.split('/')[-1]}: {rows[0].get('Kernel Name', '')[:110]}")
    seen = set()
    for r in rows:
        name = r.get("Metric Name", "")
        if any(name.startswith(k) for k in KEEP) and (r.get("Section Name"), name) not in seen:
            seen.add((r.get("Section Name"), name))
            print(f"   {r.get('Section Name', '')[:34]:34s} {name[:44]:44s} {r.get('Metric Value', ''):>14s} {r.get('Metric Unit', '')}")
    print()

"""Turn an ncu CSV launch list (gpu__time_duration + dram bytes per launch) into the
committed summary under profiles/ and the per-step DRAM traffic bench.py reports.

    python tools/ncu_summary.py gpurun_out/launches_grid.csv grid10x10 profiles/r01
"""
import csv
import json
import os
import sys

src, workload, prefix = sys.argv[1], sys.argv[2], sys.argv[3]
rows = []
with open(src) as f:
    lines = [l for l in f if not l.startswith("==")]
reader = csv.DictReader(lines)
per = {}
order = []
for r in reader:
    key = r["ID"]
    if key not in per:
        per[key] = {"kernel": r["Kernel Name"], "grid": r.get("Grid Size", ""), "block": r.get("Block Size", "")}
        order.append(key)
    name, val, unit = r["Metric Name"], r["Metric Value"].replace(",", ""), r["Metric Unit"]
    v = float(val)
    if name == "gpu__time_duration.sum":
        v = v * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1.0)
        per[key]["us"] = v
    elif name.startswith("dram__bytes"):
        v = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        per[key]["rd" if "read" in name else "wr"] = v

launches = [per[k] for k in order]
step = [l for l in launches if "sbn_" in l["kernel"]]
total_us = sum(l["us"] for l in step)
by_kernel = {}
for l in step:
    fam = l["kernel"].replace("void ", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "").split("<")[0]
    d = by_kernel.setdefault(fam, {"launches": 0, "us": 0.0, "dram_bytes": 0.0})
    d["launches"] += 1
    d["us"] += l["us"]
    d["dram_bytes"] += l.get("rd", 0) + l.get("wr", 0)
os.makedirs(os.path.dirname(prefix) or ".", exist_ok=True)
with open(prefix + f"_launches_{workload}.csv", "w") as f:
    f.write("launch,kernel,grid,block,duration_us,dram_read_bytes,dram_write_bytes\n")
    for i, l in enumerate(step):
        f.write(f"{i},\"{l['kernel']}\",\"{l['grid']}\",\"{l['block']}\",{l['us']:.3f},{l.get('rd', 0):.0f},{l.get('wr', 0):.0f}\n")
summary = {
    "workload": workload,
    "note": "ncu --clock-control none, one pass of the device program with plain launches (no CUDA graph); "
            "per-launch times are cold-cache and serialised: compare shares, not absolutes",
    "launches": len(step),
    "total_us": total_us,
    "by_kernel": {k: {**v, "share": v["us"] / total_us} for k, v in by_kernel.items()},
    "dram_bytes_per_step": sum(l.get("rd", 0) + l.get("wr", 0) for l in step),
}
with open(prefix + f"_summary_{workload}.json", "w") as f:
    json.dump(summary, f, indent=1)
traffic_path = os.path.join(os.path.dirname(prefix) or ".", "traffic.json")
traffic = {}
if os.path.exists(traffic_path):
    traffic = json.load(open(traffic_path))
dom = max(by_kernel.items(), key=lambda kv: kv[1]["us"])
steps_only = {k: v for k, v in by_kernel.items() if k.startswith(("sbn_step_tiled", "sbn_step_batched", "sbn_pair", "sbn_triple"))}
traffic[workload] = {
    "dram_bytes_per_step": sum(v["dram_bytes"] for v in steps_only.values()),
    "launches_per_step": sum(v["launches"] for v in steps_only.values()),
    "by_kernel": {k: {"launches": v["launches"], "dram_bytes": v["dram_bytes"], "us": v["us"]} for k, v in steps_only.items()},
    "dominant_kernel": dom[0], "dominant_dram_bytes_per_step": dom[1]["dram_bytes"],
    "dominant_launches_per_step": dom[1]["launches"],
    "source": "profiles/" + os.path.basename(prefix) + f"_launches_{workload}.csv (ncu dram__bytes_read.sum + dram__bytes_write.sum)"}
json.dump(traffic, open(traffic_path, "w"), indent=1)
print(json.dumps(summary, indent=1))

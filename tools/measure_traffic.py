"""DRAM traffic of the two stage kernels at the benchmark size -> profiles/traffic.json, keyed on the kernel sources.

On the GPU box (one GPU; ncu replays only the two metrics):
    python tools/measure_traffic.py run       # ncu ... > gpurun_out/traffic_67M.csv
Here, afterwards:
    python tools/measure_traffic.py summarise gpurun_out/traffic_67M.csv   # writes profiles/traffic.json

bench.py quotes `roofline.traffic` from that file only while its `source_hash` equals the hash of the kernel sources it
runs (bench.source_hash), so a stale capture is dropped, never shown.
"""
import csv
import io
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
RECIPE = ("ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:tma -c 4 -s 8 "
          "python tools/prof_case.py 512 512 256 1 1 0 2")
CELLS = 512 * 512 * 256


def run():
    out = os.path.join(ROOT, "gpurun_out", "traffic_67M.csv")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = RECIPE.split() 
    i = cmd.index("python")
    cmd = cmd[:i] + ["--csv", "--log-file", out] + [sys.executable] + cmd[i + 1:]
    return subprocess.call(cmd, cwd=ROOT)


def summarise(path):
    import bench
    rows = list(csv.DictReader(io.StringIO("".join(l for l in open(path) if l.startswith('"')))))
    acc = {}
    for r in rows:
        k = "flux_rk_o2" if "flux_rk" in r["Kernel Name"] else "grad_limiter" if "grad_limiter" in r["Kernel Name"] else None
        if k and r["Metric Name"] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            a = acc.setdefault(k, {"bytes": 0.0, "launches": set()})
            a["bytes"] += float(r["Metric Value"])
            a["launches"].add(r["ID"])
    out = {"recipe": RECIPE, "captured": time.strftime("%Y-%m-%d"), "source_hash": bench.source_hash(),
           "what": "sod_o2_visc at the benchmark size (67.1 M cells), mean over the captured launches, per owned cell"}
    for k, a in acc.items():
        out[k] = {"dram_bytes_per_cell": a["bytes"] / len(a["launches"]) / CELLS,
                  "algorithmic_bytes_per_cell": 728 if k == "flux_rk_o2" else 424, "launches": len(a["launches"])}
    with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "run":
        sys.exit(run())
    summarise(sys.argv[2])

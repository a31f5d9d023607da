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
RECIPE = ("ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:tma -c 8 -s 8 "
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
    # launches in order; a STAGE is one gradient launch followed by its flux launches (one, or the two passes of the
    # shared-cut-face scheme): only whole stages count, whatever the capture window cut off at either end is dropped
    launches = {}
    for r in rows:
        k = "flux_rk_o2" if "flux_rk" in r["Kernel Name"] else "grad_limiter" if "grad_limiter" in r["Kernel Name"] else None
        if k and r["Metric Name"] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            l = launches.setdefault(int(r["ID"]), {"kind": k, "bytes": 0.0})
            l["bytes"] += float(r["Metric Value"])
    seq = [launches[i] for i in sorted(launches)]
    starts = [i for i, l in enumerate(seq) if l["kind"] == "grad_limiter"]
    stages = [seq[a:b] for a, b in zip(starts, starts[1:])]   # the last (possibly cut) stage is dropped
    if not stages:
        raise SystemExit("no whole stage in the capture")
    out = {"recipe": RECIPE, "captured": time.strftime("%Y-%m-%d"), "source_hash": bench.source_hash(),
           "what": "sod_o2_visc at the benchmark size (67.1 M cells), per RK stage and owned cell (a stage's flux launches "
                   "together), mean over %d whole stages" % len(stages)}
    for k in ("flux_rk_o2", "grad_limiter"):
        per_stage = [sum(l["bytes"] for l in st if l["kind"] == k) for st in stages]
        out[k] = {"dram_bytes_per_cell": sum(per_stage) / len(per_stage) / CELLS,
                  "algorithmic_bytes_per_cell": 728 if k == "flux_rk_o2" else 424,
                  "launches_per_stage": sum(1 for l in stages[0] if l["kind"] == k)}
    with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "run":
        sys.exit(run())
    summarise(sys.argv[2])

"""Run every experiment build of the library (tools/build_variants.py) on the GPU: the smoke parity check, then a
throughput probe at 256x256x128; the two fastest are repeated at 512x512x256.  Not the benchmark of record.
    python tools/variant_sweep.py [tag ...] > gpurun_out/variants.jsonl
"""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VDIR = os.path.join(ROOT, "miniaero_b200", "variants")


def run(tag, args, timeout=600):
    env = dict(os.environ, MINIAERO_B200_LIB=os.path.join(VDIR, "libminiaero_b200_%s.so" % tag))
    try:
        p = subprocess.run([sys.executable] + args, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
    except subprocess.TimeoutExpired:
        return None, "timeout"
    return p.returncode, p.stdout + p.stderr


def main():
    tags = sys.argv[1:] or sorted(os.path.basename(f)[len("libminiaero_b200_"):-3] for f in glob.glob(VDIR + "/*.so"))
    results = {}
    for tag in tags:
        if tag.startswith("x"):   # timing experiments: results are wrong by construction
            rc, out = 0, "smoke OK (skipped)"
        else:
            rc, out = run(tag, ["-c", "import __graft_entry__ as g; g.smoke()"], 300)
        ok = rc == 0 and "smoke OK" in out
        print(json.dumps(dict(tag=tag, smoke=ok, detail=[l for l in out.splitlines() if "smoke" in l][-3:] if ok else out[-1500:])), flush=True)
        if not ok:
            continue
        rc, out = run(tag, ["tools/quickbench.py", "one", "256", "256", "128", tag])
        for l in out.splitlines():
            if l.startswith("{"):
                print(l, flush=True)
                results[tag] = json.loads(l)["ms_per_step"]
        if rc != 0:
            print(json.dumps(dict(tag=tag, error=out[-1500:])), flush=True)
    if os.environ.get("SWEEP_NO_BIG"):
        return
    best = sorted((t for t in results if not t.startswith("x")), key=results.get)[:2]
    if "t128b3" in results and "t128b3" not in best:
        best.append("t128b3")
    for tag in best:
        rc, out = run(tag, ["tools/quickbench.py", "one", "512", "512", "256", tag + " 67M"], 900)
        for l in out.splitlines():
            if l.startswith("{"):
                print(l, flush=True)
        if rc != 0:
            print(json.dumps(dict(tag=tag, error=out[-1500:])), flush=True)


if __name__ == "__main__":
    main()

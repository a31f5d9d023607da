SWEEP_NO_BIG=1 python tools/variant_sweep.py default > gpurun_out/variants7.jsonl 2> gpurun_out/variants7.err; cut -c1-330 gpurun_out/variants7.jsonl | grep ms_per_step
python tools/quickbench.py big 2>&1 | grep "^{" | cut -c1-330
python -m pytest tests -x -q -m gpu 2>&1 | tail -3

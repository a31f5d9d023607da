#!/bin/bash
# round-2 evidence for the final kernels: ncu launch list of the bench command, --set full of the stage kernels
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02k_launches.csv python bench.py --steps 2 --warmup 3 --no-also --no-cpu-baseline --no-parity --no-strong --e2e-steps 0 > $O/r02k_launches_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tma -s 6 -c 3 -o $O/r02k_full python tools/prof_case.py 256 256 128 1 1 0 2 > $O/r02k_full.log 2>&1
ls -la $O | tail -6

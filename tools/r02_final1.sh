#!/bin/bash
# round-2 single-GPU evidence: bench line, ncu launch list of the bench command, --set full of the two stage kernels,
# DRAM traffic at the benchmark size
cd "$(dirname "$0")/.."
O=gpurun_out
python bench.py --steps 20 --warmup 3 > $O/r02g_bench.json 2> $O/r02g_bench.err
tail -c 400 $O/r02g_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > $O/r02g_bench_reference.json 2>> $O/r02g_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02g_launches.csv python bench.py --steps 2 --warmup 3 --no-also --no-cpu-baseline --no-parity --no-strong --e2e-steps 0 > $O/r02g_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tma -s 6 -c 3 -o $O/r02g_full python tools/prof_case.py 256 256 128 1 1 0 2 > $O/r02g_full.log 2>&1
python tools/measure_traffic.py run > $O/r02g_traffic.log 2>&1
ls -la $O | tail -8

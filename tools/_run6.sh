python -m pytest tests/test_gpu_functions.py tests/test_gpu_fullsize.py -x -q -m gpu -s 2>&1 | tail -15
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01e_launches.csv python bench.py --steps 2 --warmup 1 --no-also --no-cpu-baseline --e2e-steps 1 > gpurun_out/r01e_launches_bench.log 2>&1
tail -2 gpurun_out/r01e_launches_bench.log | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:tma -c 2 -s 4 -o gpurun_out/r01e_full python tools/prof_case.py 256 256 128 1 1 0 1 > gpurun_out/r01e_full.log 2>&1
tail -2 gpurun_out/r01e_full.log
ls -la gpurun_out

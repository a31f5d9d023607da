python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r01e_2gpu.json 2> gpurun_out/bench_r01e_2gpu.err
tail -2 gpurun_out/bench_r01e_2gpu.err; cut -c1-1500 gpurun_out/bench_r01e_2gpu.json

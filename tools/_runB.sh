python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench_r01f_4gpu.json 2> gpurun_out/bench_r01f_4gpu.err
tail -2 gpurun_out/bench_r01f_4gpu.err | cut -c1-300; cut -c1-400 gpurun_out/bench_r01f_4gpu.json

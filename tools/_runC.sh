python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 3 --warmup 3 --workload flatplate_strong --e2e-steps 1 > gpurun_out/bench_r01f_strong_2gpu.json 2> gpurun_out/bench_r01f_strong_2gpu.err
tail -3 gpurun_out/bench_r01f_strong_2gpu.err | cut -c1-300; cut -c1-700 gpurun_out/bench_r01f_strong_2gpu.json

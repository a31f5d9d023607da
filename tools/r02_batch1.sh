#!/bin/bash
# round-2 GPU batch 1: first measurement of the pipelined flux kernel (correctness, then timing)
cd "$(dirname "$0")/.."
O=gpurun_out/r02a
mkdir -p $O
V=miniaero_b200/variants
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/gpu.txt
echo "== smoke tma" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== smoke pipe" ; MINIAERO_FLUX_KERNEL=pipe timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== smoke pipe pg1" ; MINIAERO_B200_LIB=$V/libminiaero_b200_pg1.so MINIAERO_FLUX_KERNEL=pipe timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== parity tests with the pipe kernel"
MINIAERO_FLUX_KERNEL=pipe timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_structured.py -m gpu -x -q 2>&1 | tail -5
for tag in tma pipe; do
  echo "== quick $tag"
  MINIAERO_FLUX_KERNEL=$tag timeout 600 python tools/quickbench.py one 256 256 128 $tag 2>&1 | tail -1 | tee -a $O/quick.jsonl
done
echo "== quick pipe pg1"
MINIAERO_B200_LIB=$V/libminiaero_b200_pg1.so MINIAERO_FLUX_KERNEL=pipe timeout 600 python tools/quickbench.py one 256 256 128 pipe_pg1 2>&1 | tail -1 | tee -a $O/quick.jsonl
for tag in tma pipe; do
  echo "== big $tag"
  MINIAERO_FLUX_KERNEL=$tag timeout 900 python tools/quickbench.py big 2>&1 | tail -1 | tee -a $O/quick.jsonl
done
echo "== big pipe pg1"
MINIAERO_B200_LIB=$V/libminiaero_b200_pg1.so MINIAERO_FLUX_KERNEL=pipe timeout 900 python tools/quickbench.py big 2>&1 | tail -1 | tee -a $O/quick.jsonl

#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r02e
mkdir -p $O
echo "== new test"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "shared_cut" 2>&1 | tail -15
echo "== whole suite with sharing forced on"; MINIAERO_SHARE_CUT_FACES=1 timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for sh in 0 1; do
  echo "== quick share=$sh"
  MINIAERO_SHARE_CUT_FACES=$sh timeout 600 python tools/quickbench.py one 256 256 128 share$sh 2>&1 | tail -1 | tee -a $O/quick.jsonl
done
for sh in 0 1; do
  echo "== big share=$sh"
  MINIAERO_SHARE_CUT_FACES=$sh timeout 900 python tools/quickbench.py big 2>&1 | tail -1 | tee -a $O/quick.jsonl
done

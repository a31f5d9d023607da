#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r02e
mkdir -p $O
echo "== suite with sharing forced on"; MINIAERO_SHARE_CUT_FACES=1 timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for sh in 0 1; do
  echo "== big share=$sh"
  MINIAERO_SHARE_CUT_FACES=$sh timeout 900 python tools/quickbench.py big 2>&1 | tail -1 | tee -a $O/quick4.jsonl
done

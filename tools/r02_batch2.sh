#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r02a
mkdir -p $O
for tile in 4,4,4 4,4,8 4,8,8; do
for tag in tma pipe; do
  echo "== quick $tag tile $tile"
  MINIAERO_TILE=$tile MINIAERO_FLUX_KERNEL=$tag timeout 600 python tools/quickbench.py one 256 256 128 "$tag-$tile" 2>&1 | tail -1 | tee -a $O/quick2.jsonl
done
done

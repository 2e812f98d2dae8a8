#!/bin/bash
# step time of config 2 with the conv pair fused from different tree levels upward (rows per window >= N)
cd "$(dirname "$0")/.."
for n in 1 3 9; do
  echo -n "R3D_TC_FUSE_MIN_ROWS=$n  "
  R3D_TC_FUSE_MIN_ROWS=$n timeout 120 python scripts/gpu_diag.py forward bf16x3 243 1024 1 | cut -c1-120
done

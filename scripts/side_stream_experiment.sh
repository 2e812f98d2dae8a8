#!/bin/bash
# step time of config 2 with the GlobalInfo chain serialised / on the side stream at different priorities
cd "$(dirname "$0")/.."
for cfg in "R3D_SIDE_STREAM=0" "R3D_SIDE_PRIO=0" "R3D_SIDE_PRIO=1" "R3D_SIDE_PRIO=-1"; do
  echo -n "$cfg  "
  env $cfg timeout 120 python scripts/gpu_diag.py forward bf16x3 243 1024 1 | cut -c1-120
done

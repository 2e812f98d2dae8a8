#!/bin/bash
# timing experiments on the TC GEMM epilogue (R3D_TC_DEBUG: 1 skip stores, 2 skip residual, 4 skip epilogue math)
for shape in "82944 256 256 6" "27648 256 256 6"; do
  for dbg in 0 1 2 3 7; do
    echo -n "shape=$shape dbg=$dbg "; R3D_TC_DEBUG=$dbg timeout 100 python scripts/gpu_diag.py selftest $shape bf16x3 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms_tc=%.4f tflops=%.1f err=%.2e'%(d['ms_tc'],d['tflops_tc'],d['err']))"
  done
done

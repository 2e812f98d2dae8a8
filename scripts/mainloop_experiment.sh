#!/bin/bash
# main-loop rate experiments: R3D_TC_DEBUG=7 removes epilogue math/stores/residual; R3D_TC_CLUSTER=0 -> 1-CTA tiles
for cl in 1 0; do
 for prec in bf16x3 bf16; do
  for shape in "27648 256 768 6" "82944 256 256 6"; do
   for dbg in 7 3 0; do
    echo -n "cluster=$cl prec=$prec shape=$shape dbg=$dbg "; R3D_TC_CLUSTER=$cl R3D_TC_DEBUG=$dbg timeout 100 python scripts/gpu_diag.py selftest $shape $prec | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms_tc=%.4f tflops=%.1f'%(d['ms_tc'],d['tflops_tc']))"
   done
  done
 done
done

"""cuBLAS bf16 GEMM under the profiler: the calibration point for `sm__pipe_tensor_cycles_active` (what does the counter read
for a kernel that runs at the measured peak of MEASURED_PEAKS.json?).  Run under
    ncu --set full --clock-control none -k regex:'nvjet|gemm|cutlass|xmma' -c 2 -o gpurun_out/r2_cublas python scripts/cublas_calibration.py
and without ncu for the plain timing line."""
import json
import sys

import torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
a = torch.randn(n, n, device="cuda", dtype=torch.bfloat16)
b = torch.randn(n, n, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    c = a @ b
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for _ in range(5):
    e0.record()
    c = a @ b
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(json.dumps({"n": n, "ms_best": best, "tflops": 2.0 * n ** 3 / best / 1e9}))

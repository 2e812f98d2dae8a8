"""A few serial forwards of one BASELINE configuration, for profiling under ncu:
    ncu --set full --clock-control none --import-source on -k regex:"gemm_tc|prologue|assemble|tail_tc" -s 38 -c 19 -o gpurun_out/prof python scripts/one_forward.py
    python scripts/one_forward.py [cfg2|cfg3|cfg5] [forwards]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ray3d_b200 import Lifter, NetSpec, synth

CFG = {"cfg2": ((3, 3, 3, 3, 3), 1, 1024, "bf16x3"), "cfg3": ((3, 3, 3, 3), 1, 4096, "bf16"), "cfg5": ((3, 3, 3, 3, 3), 3, 512, "bf16x3")}
widths, stage, B, prec = CFG[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
spec = NetSpec(filter_widths=widths, stage=stage)
sp, st = synth.make_state_dicts(spec)
lf = Lifter(spec, sp, st, precision=prec, device=0)
lf.plan.set_option("side_stream", 0)             # serial launch order: the profiler serialises kernels anyway
uv, cam = synth.make_inputs(spec, B, seed=1)
uvc, camc = torch.from_numpy(uv).cuda(), torch.from_numpy(cam).cuda()
for _ in range(n):
    out = lf.forward_uv(uvc, camc)[2]
torch.cuda.synchronize()
print("launches per forward:", lf.plan.kernel_launches, "finite:", bool(torch.isfinite(out).all()))

"""Cycle accounting of the chained tail launch (experiment builds: R3D_BUILD_EXPERIMENTS=1 python -m ray3d_b200.build --force).
python scripts/tail_stats.py [B]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ray3d_b200 import Lifter, NetSpec, synth, _capi

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
spec = NetSpec(filter_widths=(3, 3, 3, 3, 3))
sp, st = synth.make_state_dicts(spec)
W = int(sys.argv[2]) if len(sys.argv) > 2 else 128
lf = Lifter(spec, sp, st, precision="bf16x3", device=0, options={"tail_width": W})
uv, cam = synth.make_inputs(spec, B, seed=5)
uvc, camc = torch.from_numpy(uv).cuda(), torch.from_numpy(cam).cuda()
for _ in range(5):
    lf.forward_uv(uvc, camc)
buf = (C.c_uint64 * 8)()
_capi.check(_capi.lib().r3d_debug_tail_stats(buf))
N = 20
for _ in range(N):
    lf.forward_uv(uvc, camc)
_capi.check(_capi.lib().r3d_debug_tail_stats(buf))
v = [int(x) for x in buf]
ncta = 74
names = ["dep spin", "queue slot wait", "ring slot wait (producer)", "mma: operands not landed", "mma: accumulator not drained", "mma: next unit not published", "units", "kernel cycles per leader CTA"]
import time
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(50):
    lf.forward_uv(uvc, camc)
torch.cuda.synchronize(); print(f"width {W}: {(time.perf_counter()-t0)/50*1e3:.3f} ms per forward (serial)")
print(f"B={B}: per forward, per leader CTA (74 clusters assumed), cycles")
for n, x in zip(names, v):
    print(f"  {n:34s} {x / N / ncta:12.0f}")

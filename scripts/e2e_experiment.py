"""End-to-end (host buffers) timing for different chunk counts: R3D_HOST_CHUNKS=n python scripts/e2e_experiment.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ray3d_b200 import Lifter, NetSpec, synth
spec = NetSpec(filter_widths=(3, 3, 3, 3, 3))
sp, st = synth.make_state_dicts(spec)
lf = Lifter(spec, sp, st, precision="bf16x3")
B = 1024
uv, cam = synth.make_inputs(spec, B, seed=1)
uvh, camh = torch.from_numpy(uv).pin_memory(), torch.from_numpy(cam).pin_memory()
out = torch.empty((B, 1, 17, 3), dtype=torch.float32).pin_memory()
for _ in range(3):
    lf.forward_uv_host(uvh, camh, out=out)
t0 = time.perf_counter()
n = 10
for _ in range(n):
    lf.forward_uv_host(uvh, camh, out=out)
dt = (time.perf_counter() - t0) / n
# plain H2D rate for reference
d = torch.empty_like(uvh, device="cuda")
torch.cuda.synchronize(); t1 = time.perf_counter()
for _ in range(10):
    d.copy_(uvh, non_blocking=True)
torch.cuda.synchronize(); h2d = (time.perf_counter() - t1) / 10
print(f"chunks={os.environ.get('R3D_HOST_CHUNKS','4')} e2e {dt*1e3:.3f} ms/step {B/dt:.0f} seq/s ; plain H2D of uv {h2d*1e3:.3f} ms ({uvh.numel()*4/h2d/1e9:.1f} GB/s)")

"""Streaming end-to-end diagnostics: r3d_submit_host / r3d_wait throughput for different depths / lane counts, next to the plain
pinned H2D rate of the same buffers.  python scripts/e2e_diag.py [B]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ray3d_b200 import Lifter, NetSpec, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
spec = NetSpec(filter_widths=(3, 3, 3, 3, 3))
sp, st = synth.make_state_dicts(spec)
lf = Lifter(spec, sp, st, precision="bf16x3", device=0)
sets = []
for i in range(4):
    uv, cam = synth.make_inputs(spec, B, seed=5 + i)
    sets.append((torch.from_numpy(uv).pin_memory(), torch.from_numpy(cam).pin_memory()))
outs = [torch.empty((B, 1, 17, 3), dtype=torch.float32).pin_memory() for _ in range(4)]
d = torch.empty_like(sets[0][0], device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(20):
    d.copy_(sets[i % 4][0], non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 20
print(f"plain pinned H2D: {sets[0][0].numel()*4/dt/1e9:.1f} GB/s ({dt*1e3:.3f} ms per {sets[0][0].numel()*4/1e6:.1f} MB)")
for i in range(3):
    lf.forward_uv_host(*sets[i], out=outs[i])
for lanes in (2, 1):
    lf.plan.set_option("lanes", lanes)
    for depth in (1, 2, 3, 4):
        pend = []
        steps = 40
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tsub = 0.0
        for i in range(steps):
            if len(pend) == depth:
                lf.wait(pend.pop(0))
            a = time.perf_counter()
            pend.append(lf.submit_uv_host(*sets[i % 4], out=outs[i % 4]))
            tsub += time.perf_counter() - a
        for tk in pend:
            lf.wait(tk)
        dt = (time.perf_counter() - t0) / steps
        print(f"lanes={lanes} depth={depth}: {dt*1e3:.3f} ms/step {B/dt:.0f} seq/s  (host time inside submit: {tsub/steps*1e3:.3f} ms/step)")
    t0 = time.perf_counter()
    for i in range(20):
        lf.forward_uv_host(*sets[i % 4], out=outs[i % 4])
    dt = (time.perf_counter() - t0) / 20
    print(f"lanes={lanes} blocking: {dt*1e3:.3f} ms/step {B/dt:.0f} seq/s")

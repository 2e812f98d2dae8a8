"""Pipeline-fill diagnostics of the streaming host path: how long until the FIRST result of a burst of k submissions is in
host memory (k = 1: copy + kernels + read-back of one step, nothing else in flight).  python scripts/e2e_ramp.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ray3d_b200 import Lifter, NetSpec, synth

B = 1024
spec = NetSpec(filter_widths=(3, 3, 3, 3, 3))
sp, st = synth.make_state_dicts(spec)
lf = Lifter(spec, sp, st, precision="bf16x3", device=0)
sets = []
for i in range(4):
    uv, cam = synth.make_inputs(spec, B, seed=5 + i)
    sets.append((torch.from_numpy(uv).pin_memory(), torch.from_numpy(cam).pin_memory()))
outs = [torch.empty((B, 1, 17, 3), dtype=torch.float32).pin_memory() for _ in range(4)]
for i in range(6):
    lf.wait(lf.submit_uv_host(*sets[i % 4], out=outs[i % 4]))
for k in (1, 2, 3, 4):
    res = []
    for rep in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tks = [lf.submit_uv_host(*sets[i], out=outs[i]) for i in range(k)]
        t_sub = time.perf_counter() - t0
        lf.wait(tks[0])
        t_first = time.perf_counter() - t0
        for t in tks[1:]:
            lf.wait(t)
        t_all = time.perf_counter() - t0
        res.append((t_sub * 1e3, t_first * 1e3, t_all * 1e3))
    res.sort(key=lambda r: r[1])
    print(f"burst of {k}: submit {res[2][0]:.2f} ms, first result {res[2][1]:.2f} ms, all {res[2][2]:.2f} ms")

"""Tiny forward for compute-sanitizer runs (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python scripts/sanitize_small.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from ray3d_b200 import Lifter, NetSpec, synth

spec = NetSpec(filter_widths=(3, 3, 3))
sp, st = synth.make_state_dicts(spec)
# default plan; the GlobalInfo chain forced into the chained launch (tail_tc_kernel); the main tail chained as well
for opts in ({}, {"side_chain": 2}, {"tail_fusion": 1, "side_chain": 2}):
    lf = Lifter(spec, sp, st, options=opts)
    for B in ((1500, 300, 7) if not opts else (300,)):      # 1500 windows: multi-wave launches claim their units dynamically
        uv, cam = synth.make_inputs(spec, B, seed=3)
        uvc, camc = torch.from_numpy(uv).cuda(), torch.from_numpy(cam).cuda()
        a = lf.forward_uv(uvc, camc)[2]
        p0, p1 = lf.submit_uv(uvc, camc), lf.submit_uv(uvc, camc)
        b, c = lf.join(p0)[2], lf.join(p1)[2]
        torch.cuda.synchronize()
        print(opts, B, bool(torch.equal(a, b)), bool(torch.equal(a, c)), float(a.abs().mean()))
    del lf

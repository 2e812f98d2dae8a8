"""A/B timing of plan-level knobs inside ONE process (same clocks, same thermal state): every variant is a Lifter
built under its own environment settings; timing blocks are interleaved A B C A B C ... and the per-variant median
and minimum over the rounds are printed.  Usage (under gpurun):
    python scripts/ab_compare.py [--T 243 --B 1024 --stage 1 --prec bf16x3 --rounds 7 --iters 100] "" "R3D_TC_FUSE_MIN_ROWS=1" ...
An empty string is the default configuration.  Only knobs read at plan construction time can be compared this way."""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--T", type=int, default=243)
    ap.add_argument("--B", type=int, default=1024)
    ap.add_argument("--stage", type=int, default=1)
    ap.add_argument("--prec", default="bf16x3")
    ap.add_argument("--rounds", type=int, default=7)
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("variants", nargs="+")
    a = ap.parse_args()
    import torch
    from ray3d_b200 import Lifter, NetSpec, synth
    widths = {9: (3, 3), 27: (3, 3, 3), 81: (3, 3, 3, 3), 243: (3, 3, 3, 3, 3)}[a.T]
    spec = NetSpec(filter_widths=widths, stage=a.stage)
    sp, st = synth.make_state_dicts(spec)
    uv, cam = synth.make_inputs(spec, a.B, seed=1)
    uvc, camc = torch.from_numpy(uv).cuda(), torch.from_numpy(cam).cuda()
    lifters = []
    for v in a.variants:
        kv = dict(x.split("=", 1) for x in v.split(",") if x)
        old = {k: os.environ.get(k) for k in kv}
        os.environ.update(kv)
        lf = Lifter(spec, sp, st, precision=a.prec)
        for _ in range(5):
            lf.forward_uv(uvc, camc)        # binds the workspace under this variant's environment
        torch.cuda.synchronize()
        for k, o in old.items():
            if o is None:
                os.environ.pop(k)
            else:
                os.environ[k] = o
        lifters.append(lf)
    ref = lifters[0].forward_uv(uvc, camc)[2].clone()
    times = [[] for _ in lifters]
    for r in range(a.rounds):
        for i, lf in enumerate(lifters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            lf.forward_uv(uvc, camc)
            e0.record()
            for _ in range(a.iters):
                out = lf.forward_uv(uvc, camc)[2]
            e1.record()
            torch.cuda.synchronize()
            times[i].append(e0.elapsed_time(e1) / a.iters)
    for v, t, lf in zip(a.variants, times, lifters):
        same = bool(torch.equal(lf.forward_uv(uvc, camc)[2], ref))
        print(json.dumps(dict(variant=v or "default", median_ms=round(statistics.median(t), 4), min_ms=round(min(t), 4),
                              seq_per_s=round(a.B / statistics.median(t) * 1e3), equal_to_first=same)))


if __name__ == "__main__":
    main()

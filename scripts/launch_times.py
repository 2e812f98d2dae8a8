"""Per-launch CUDA-event durations of one forward (launches serialised on one stream), median over the profiled runs.
    [R3D_TC_DEBUG=n ...] python scripts/launch_times.py [--T 243 --B 1024 --stage 1 --prec bf16x3] [--only expand_conv,...]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--T", type=int, default=243)
    ap.add_argument("--B", type=int, default=1024)
    ap.add_argument("--stage", type=int, default=1)
    ap.add_argument("--prec", default="bf16x3")
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    import torch
    from ray3d_b200 import Lifter, NetSpec, synth
    widths = {9: (3, 3), 27: (3, 3, 3), 81: (3, 3, 3, 3), 243: (3, 3, 3, 3, 3)}[a.T]
    spec = NetSpec(filter_widths=widths, stage=a.stage)
    sp, st = synth.make_state_dicts(spec)
    lf = Lifter(spec, sp, st, precision=a.prec)
    uv, cam = synth.make_inputs(spec, a.B, seed=1)
    uvc, camc = torch.from_numpy(uv).cuda(), torch.from_numpy(cam).cuda()
    for _ in range(5):
        lf.forward_uv(uvc, camc)
    lf.plan.set_profiling(True)
    for _ in range(40):
        lf.forward_uv(uvc, camc)
    torch.cuda.synchronize()
    times, runs = lf.plan.launch_times()
    only = set(x for x in a.only.split(",") if x)
    out = {n: round(ms * 1e3, 1) for n, ms in times if not only or n in only}
    out["_total_us"] = round(sum(ms for _, ms in times) * 1e3, 1)
    out["_env"] = {k: v for k, v in os.environ.items() if k.startswith("R3D_")}
    print(json.dumps(out))


if __name__ == "__main__":
    main()

"""Does keeping two independent batches in flight on two streams raise throughput?  The tail of a step (upper tree
levels, FC heads: one-wave launches) leaves SMs idle that the head of the next step could use.  Two Lifters (two
plans: own workspace, own side stream) alternate steps on two torch streams; compared with one Lifter on one stream.
    python scripts/two_lane_experiment.py [--B 1024 --T 243 --iters 200 --rounds 5]"""
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
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--rounds", type=int, default=5)
    ap.add_argument("--lanes", type=int, default=2)
    a = ap.parse_args()
    import torch
    from ray3d_b200 import Lifter, NetSpec, synth
    widths = {9: (3, 3), 27: (3, 3, 3), 81: (3, 3, 3, 3), 243: (3, 3, 3, 3, 3)}[a.T]
    spec = NetSpec(filter_widths=widths)
    sp, st = synth.make_state_dicts(spec)
    sets = []
    for i in range(4):
        uv, cam = synth.make_inputs(spec, a.B, seed=10 + i)
        sets.append((torch.from_numpy(uv).cuda(), torch.from_numpy(cam).cuda()))
    lifters = [Lifter(spec, sp, st) for _ in range(a.lanes)]
    streams = [torch.cuda.Stream() for _ in range(a.lanes)]
    for lf in lifters:
        for i in range(3):
            lf.forward_uv(*sets[i])
    torch.cuda.synchronize()

    def run_single(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            lifters[0].forward_uv(*sets[i % 4], want_pos=False)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    def run_lanes(n):
        main_s = torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in streams:
            s.wait_stream(main_s)
        for i in range(n):
            with torch.cuda.stream(streams[i % a.lanes]):
                lifters[i % a.lanes].forward_uv(*sets[i % 4], want_pos=False)
        for s in streams:
            main_s.wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    t1, t2 = [], []
    for r in range(a.rounds):
        t1.append(run_single(a.iters))
        t2.append(run_lanes(a.iters))
    print(json.dumps(dict(B=a.B, T=a.T, lanes=a.lanes, single_ms=round(statistics.median(t1), 4), lanes_ms=round(statistics.median(t2), 4),
                          single_seq_s=round(a.B / statistics.median(t1) * 1e3), lanes_seq_s=round(a.B / statistics.median(t2) * 1e3))))


if __name__ == "__main__":
    main()

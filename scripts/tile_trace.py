"""Per-tile SM-clock trace of ONE tensor-core GEMM launch of the forward pass (CTA 0: TMA producer, MMA thread, first
epilogue warp), via r3d_debug_tc_trace.  Prints, per traced tile, cycle offsets from the launch's first stamp.
    python scripts/tile_trace.py --op 0 [--T 243 --B 1024]        # op index within the launch graph (0 = expand_conv)"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--T", type=int, default=243)
    ap.add_argument("--B", type=int, default=1024)
    ap.add_argument("--op", type=int, default=0)
    ap.add_argument("--tiles", type=int, default=30)
    a = ap.parse_args()
    import numpy as np
    import torch
    from ray3d_b200 import Lifter, NetSpec, synth, _capi
    widths = {9: (3, 3), 27: (3, 3, 3), 81: (3, 3, 3, 3), 243: (3, 3, 3, 3, 3)}[a.T]
    spec = NetSpec(filter_widths=widths)
    sp, st = synth.make_state_dicts(spec)
    lf = Lifter(spec, sp, st)
    uv, cam = synth.make_inputs(spec, a.B, seed=1)
    uvc, camc = torch.from_numpy(uv).cuda(), torch.from_numpy(cam).cuda()
    for _ in range(5):
        lf.forward_uv(uvc, camc)
    torch.cuda.synchronize()
    names = [o["name"] for o in lf.plan.describe()["ops"]]
    _capi.check(_capi.lib().r3d_debug_tc_trace(a.op, None, 0))
    lf.forward_uv(uvc, camc)
    buf = (C.c_int64 * (4 * 64 * 8))()
    _capi.check(_capi.lib().r3d_debug_tc_trace(0, buf, len(buf)))
    tr = np.array(buf, dtype=np.int64).reshape(4, 64, 8)
    t0 = tr[tr > 0].min()
    rel = np.where(tr > 0, tr - t0, -1)
    print("op", a.op, names[a.op], "(launch order: main-stream ops interleave with the GlobalInfo side chain)")
    print("producer : [tile start, loads issued]")
    print("mma      : [tile start, accumulator free, first operands landed, MMAs issued, (fused: 2nd GEMM issued)]")
    print("epilogue : [tile start, bias staged, accumulator full, chunk0..3 done, tile done]")
    print("store thr: per chunk round [staged tile ready, store engine has read it] x 4")
    for ti in range(min(a.tiles, 64)):
        if rel[2, ti, 0] < 0:
            break
        print(f"tile {ti:2d}  P {rel[0, ti, :2].tolist()}  M {rel[1, ti, :5].tolist()}  E {rel[2, ti, :8].tolist()}  S {rel[3, ti, :8].tolist()}")
    e = rel[2]
    n = int((e[:, 0] >= 0).sum())
    if n > 3:
        d = lambda i, j: float(np.median(e[2:n, j] - e[2:n, i]))
        print(json.dumps(dict(tiles=n, period=float(np.median(np.diff(e[1:n, 0]))), bias=d(0, 1), wait_full=d(1, 2), chunk0=d(2, 3), chunk1=d(3, 4),
                              chunk2=d(4, 5), chunk3=d(5, 6), tail=d(6, 7))))


if __name__ == "__main__":
    main()

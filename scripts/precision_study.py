"""Operand-rounding study: which tensor-core operand format meets the 1e-4 fp32 parity bar?

Emulates (on CPU) GEMMs whose operands are rounded/split the way the tcgen05 path would do it and
pushes them through the whole network (oracle graph), reporting normwise error vs the reference
fp64 golden output.  Test infrastructure (uses oracle/); results are recorded in DESIGN.md.
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.nn.functional as F
from oracle import ray3d_oracle as O
from ray3d_b200 import synth
from ray3d_b200.spec import NetSpec
import json

def bf16(x): return x.to(torch.bfloat16).to(torch.float32)
def tf32(x):
    i = x.contiguous().view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF     # round-to-nearest (ties away) to 10 mantissa bits
    return i.view(torch.float32)

def split(x, rnd, n):
    parts, r = [], x
    for _ in range(n):
        p = rnd(r); parts.append(p); r = r - p
    return parts

def make_mm(mode):
    def mm(a, w):   # a (M,K) fp32, w (N,K) fp32 -> fp32
        if mode == "fp32": return a @ w.t()
        rnd, n, prods = {"bf16": (bf16, 1, [(0, 0)]), "tf32": (tf32, 1, [(0, 0)]),
                         "bf16x3": (bf16, 2, [(0, 0), (0, 1), (1, 0)]),
                         "bf16x6": (bf16, 3, [(0, 0), (0, 1), (1, 0), (0, 2), (1, 1), (2, 0)]),
                         "tf32x3": (tf32, 2, [(0, 0), (0, 1), (1, 0)])}[mode]
        A, W = split(a, rnd, n), split(w, rnd, n)
        acc = torch.zeros(a.shape[0], w.shape[0], dtype=torch.float64)
        for i, j in prods: acc += A[i].double() @ W[j].double().t()
        return acc.float()
    return mm

def run(name, mode, meta):
    kw = dict(meta[name]["spec"]); kw["filter_widths"] = tuple(kw["filter_widths"]); spec = NetSpec(**kw)
    g = dict(np.load(f"tests/golden/{name}.npz"))
    sp, st = (O.to_torch_state(s) for s in synth.make_state_dicts(spec))
    mm = make_mm(mode)
    o_conv, o_lin = F.conv1d, F.linear
    def conv1d(x, w, b=None, stride=1):
        k = w.shape[2]; B, C, T = x.shape; To = T // stride if k == stride else T
        a = x.permute(0, 2, 1).reshape(B * To, k * C) if k == stride else x.permute(0, 2, 1).reshape(B * T, C)
        wm = w.permute(0, 2, 1).reshape(w.shape[0], k * C)
        y = mm(a, wm)
        if b is not None: y = y + b
        return y.reshape(B, To, -1).permute(0, 2, 1)
    def linear(x, w, b=None):
        y = mm(x, w)
        return y + b if b is not None else y
    F.conv1d, F.linear = conv1d, linear
    try:
        pos, trj, both = O.lift(sp, st, spec, torch.from_numpy(g["x"]), torch.from_numpy(g["param"]))
    finally:
        F.conv1d, F.linear = o_conv, o_lin
    r = lambda a, b: float(np.linalg.norm(a.double().numpy() - b) / np.linalg.norm(b))
    ref = g["pos64"] + g["trj64"]
    mx = float(np.abs(both.double().numpy() - ref).max() / np.abs(ref).max())
    return r(pos, g["pos64"]), r(trj, g["trj64"]), r(both, ref), mx

meta = json.load(open("tests/golden/meta.json"))
for name in ("h36m_s1_t243", "3dhp_s3_t243", "h36m_s1_t27"):
    for mode in ("fp32", "tf32", "bf16", "bf16x3", "tf32x3", "bf16x6"):
        print(f"{name:14s} {mode:7s} pos %.2e trj %.2e sum %.2e maxabs/max %.2e" % run(name, mode, meta))

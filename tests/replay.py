"""TEST INFRASTRUCTURE: numpy replay of the native plan's launch graph.

Takes the JSON graph (`Plan.describe()`), the BN-folded packed weights (`Plan.packed_layer`) and
replays prologue -> grouped GEMMs -> assemble with numpy, so wiring, packing and the input-stage
gather tables can be checked against the oracle/goldens on a box without a GPU.  Never imported
by the product package.
"""
import numpy as np


def first_layer_operand(xs, g):
    """(B, T, J*Cin) window -> (B, L0, k_pad) shared first-layer operand, columns per the plan's a0_map."""
    B, T, JC = xs.shape
    tc, w0, L0 = g["tc"], g["w0"], g["L0"]
    src = np.concatenate([xs.reshape(B, L0, w0 * JC), np.repeat(xs[:, tc][:, None, :], L0, axis=1)], axis=2)
    amap = np.asarray(g["a0_map"])
    out = src[:, :, np.maximum(amap, 0)]
    out[:, :, amap < 0] = 0
    return out


def replay(plan, x, param, dtype=np.float32):
    g = plan.describe()
    B, T, J, Cin = x.shape
    assert T == g["T"] and J == g["J"] and Cin == g["Cin"]
    JC = J * Cin
    mats = [np.zeros((B * r, ld), dtype=dtype) for r, ld, _ in g["mats"]]
    xs = x.reshape(B, T, JC).astype(dtype)
    tc, w0, L0 = g["tc"], g["w0"], g["L0"]
    # ---- input stage: one operand shared by all first-layer problems; row (b, tq) draws its columns from
    # [w0 frames | x[tc]] through the plan's column map (-1 = zero; the x - root / x - x[tc] differences live in the
    # folded expand_conv weights, the column order groups each joint group's columns into few K steps)
    A0 = mats[g["a0"][0]].reshape(B, L0, -1)
    A0[...] = first_layer_operand(xs, g)
    mats[g["inc"]][:, :JC] = xs[:, tc]
    lrelu = lambda v, s: np.where(v > 0, v, v * dtype(s))
    for e in g["embed"]:
        net = 1 << e["net"]
        w1, b1 = plan.packed_layer(net, "embedder.w1")
        w2, b2 = plan.packed_layer(net, "embedder.w2")
        h = lrelu(param.astype(dtype) @ w1.T.astype(dtype) + b1.astype(dtype), 0.01)
        o = lrelu(h @ w2.T.astype(dtype) + b2.astype(dtype), 0.01)
        for mid, col in e["dst"]:
            mats[mid][:, col:col + o.shape[1]] = o
    # ---- grouped GEMMs
    for op in g["ops"]:
        M = B * op["rows_per_seq"]
        for pr in op["prob"]:
            net, layer = pr["layer"].split(":", 1)
            w, b = plan.packed_layer(1 << int(net), layer)
            npad, K = w.shape
            A = mats[pr["a"]].reshape(-1)[: M * pr["a_ld"]].reshape(M, pr["a_ld"])[:, :K]
            out = A @ w.T.astype(dtype) + b.astype(dtype)
            if op["slope"] != 1:
                out = lrelu(out, op["slope"])
            if "layer2" in pr:        # fused conv pair: second (1x1) GEMM on the activated intermediate
                net2, layer2 = pr["layer2"].split(":", 1)
                w2, b2 = plan.packed_layer(1 << int(net2), layer2)
                out = lrelu(out[:, :w2.shape[1]] @ w2.T.astype(dtype) + b2.astype(dtype), op["slope"])
                npad = w2.shape[0]
            if pr["res"] >= 0:
                R = mats[pr["res"]].reshape(-1)[: M * pr["res_ld"]].reshape(M, pr["res_ld"])
                out = out + R[:, pr["res_col"]:pr["res_col"] + npad]
            for mid, col in pr["dst"]:
                n = min(npad, mats[mid].shape[1] - col)
                mats[mid][:M, col:col + n] = out[:, :n]
    # ---- assemble
    pos = trj = None
    if g["has_trj"]:
        trj = mats[g["heads"][5]][:, :3].reshape(B, 1, 1, 3).copy()
    if g["has_pos"]:
        pos = np.stack([mats[g["heads"][grp]][:, 3 * k:3 * k + 3] for grp, k in g["slots"]], axis=1).reshape(B, 1, J, 3)
    return pos, trj

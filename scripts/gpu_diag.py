"""GPU diagnostics run under gpurun: tensor-core GEMM self-tests (each in its own process so a trap or
hang cannot take the rest down) and quick forward timings.  Writes gpurun_out/diag.log."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SHAPES = [(128, 16, 64, 1), (128, 256, 64, 1), (128, 256, 128, 1), (256, 128, 256, 1), (300, 256, 768, 2), (1024, 1024, 1024, 1),
          (648, 256, 192, 6), (37, 16, 1024, 3), (5, 1024, 576, 2), (27648, 256, 768, 6), (82944, 256, 128, 6), (1024, 1024, 1024, 6)]


def child_selftest(args):
    from ray3d_b200 import _capi
    m, n, k, p, prec = int(args[0]), int(args[1]), int(args[2]), int(args[3]), args[4]
    err, t_tc, t_ff = _capi.selftest_gemm(m, n, k, p, prec)
    fl = 2.0 * m * n * k * p
    print(json.dumps(dict(m=m, n=n, k=k, p=p, prec=prec, err=err, ms_tc=t_tc, ms_ffma=t_ff, tflops_tc=fl / t_tc / 1e9,
                          tflops_ffma=fl / t_ff / 1e9)))


def child_forward(args):
    import numpy as np
    import torch
    from ray3d_b200 import Lifter, NetSpec, synth
    prec, T, B, stage = args[0], int(args[1]), int(args[2]), int(args[3])
    widths = {9: (3, 3), 27: (3, 3, 3), 81: (3, 3, 3, 3), 243: (3, 3, 3, 3, 3)}[T]
    spec = NetSpec(filter_widths=widths, stage=stage)
    sp, st = synth.make_state_dicts(spec)
    lf = Lifter(spec, sp, st, precision=prec)
    uv, cam = synth.make_inputs(spec, B, seed=1)
    uvc, camc = torch.from_numpy(uv).cuda(), torch.from_numpy(cam).cuda()
    for _ in range(3):
        out = lf.forward_uv(uvc, camc)[2]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 100
    for _ in range(n):
        out = lf.forward_uv(uvc, camc)[2]
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(json.dumps(dict(prec=prec, T=T, B=B, stage=stage, ms=ms, seq_per_s=B / ms * 1e3, finite=bool(torch.isfinite(out).all()),
                          ws_mb=lf.plan.workspace_bytes / 1e6)))


def run(cmd, timeout):
    t = time.time()
    try:
        r = subprocess.run([sys.executable, __file__] + cmd, capture_output=True, text=True, timeout=timeout)
        out = (r.stdout.strip().splitlines() or [""])[-1]
        tail = r.stderr.strip().splitlines()[-3:] if r.returncode else []
        return f"rc={r.returncode} {time.time() - t:5.1f}s {out} {' | '.join(tail)}"
    except subprocess.TimeoutExpired:
        return f"TIMEOUT after {timeout}s: {cmd}"


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "selftest":
        child_selftest(sys.argv[2:])
    elif len(sys.argv) > 1 and sys.argv[1] == "forward":
        child_forward(sys.argv[2:])
    else:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "diag.log"), "w") as f:
            def emit(s):
                print(s, flush=True)
                f.write(s + "\n")
                f.flush()
            for prec in ("bf16x3", "bf16"):
                for (m, n, k, p) in (SHAPES if prec == "bf16x3" else SHAPES[-4:]):
                    emit(run(["selftest", str(m), str(n), str(k), str(p), prec], 120))
            for prec, T, B, stage in (("fp32", 243, 1024, 1), ("bf16x3", 243, 1024, 1), ("bf16", 81, 4096, 1), ("bf16x3", 243, 512, 3),
                                      ("bf16x3", 27, 1, 1)):
                emit(run(["forward", prec, str(T), str(B), str(stage)], 300))

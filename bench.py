#!/usr/bin/env python
"""Benchmark of the Ray3D lifting hot path (BASELINE.json metric: sequences/sec, T=243, 17 joints).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # reference algorithm on the host CPU cores

A "step" is one pass of the hot path (camera-ray encode -> pose net + trajectory net -> pos + trj) over one
batch of synthetic sequences: BASELINE.json configs[1] (batch 1024, T=243, 17 joints, fp32 parity bar) per
GPU.  Under torchrun every rank lifts its own 1024-sequence shard (weak scaling) and the step ends with the
single all-gather of the outputs.  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

WORKLOADS = {
    # BASELINE.json configs; cfg1 (T=27, B=1) is the reference's CPU-runnable latency case -> parity tests only
    "cfg2": dict(desc="batch=1024/GPU, T=243, 17 joints, stage 1 (cfg_ray3d_h36m_stage1 arch '3,3,3,3,3')",
                 widths=(3, 3, 3, 3, 3), stage=1, batch=1024, res=1000),
    "cfg3": dict(desc="batch=4096, T=81, 17 joints, stage 1, bf16", widths=(3, 3, 3, 3), stage=1, batch=4096, res=1000),
    "cfg5": dict(desc="cfg_ray3d_3dhp_stage3 arch, T=243, batch=512/GPU", widths=(3, 3, 3, 3, 3), stage=3, batch=512, res=2048),
}


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks/throttle sampling during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        for ts, ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                clk, cmax = float(f[1]), float(f[2])
            except ValueError:
                continue
            mx.append(cmax)
            if t0 - 0.05 <= ts <= t1 + 0.15:
                sm.append(clk)
                try:
                    power.append(float(f[3]))
                except ValueError:
                    pass
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        if not sm:   # region shorter than one sample: fall back to every sample taken
            sm = [float(ln.split(",")[1]) for _, ln in self.lines if len(ln.split(",")) >= 9 and ln.split(",")[1].strip().replace(".", "").isdigit()]
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=sorted(reasons),
                    samples=len(sm), power_w_max=max(power) if power else None)


def cpu_reference_throughput(wl, spec, sample: int, steps: int, warmup: int):
    """Times the oracle port of the reference's eval step (ray encode in numpy float64 + torch CPU modules'
    functional ops) on all host threads.  Returns (seq/s, cores, ms/step)."""
    from oracle import ray3d_oracle as O
    from ray3d_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sp, st = synth.make_state_dicts(spec)
    sp, st = O.to_torch_state(sp), O.to_torch_state(st)
    uv, cam = synth.make_inputs(spec, sample, seed=1234 + 1, res=wl["res"])
    for _ in range(warmup):
        O.lift_uv(sp, st, spec, uv, cam)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.lift_uv(sp, st, spec, uv, cam)
    dt = (time.perf_counter() - t0) / steps
    return sample / dt, cores, dt * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)     # ~0.3 s of timed work: long enough for clock sampling
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default=None, choices=["fp32", "bf16x3", "bf16"])
    ap.add_argument("--cpu-sample", type=int, default=256, help="sequences per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    from ray3d_b200 import NetSpec, synth
    from ray3d_b200.spec import flops_per_sequence

    wl = WORKLOADS[args.workload]
    spec = NetSpec(num_joints=17, in_features=3, filter_widths=wl["widths"], stage=wl["stage"])
    precision = args.precision or ("bf16" if args.workload == "cfg3" else "bf16x3")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    B = wl["batch"]
    config = dict(workload=f"{args.workload}: {wl['desc']}", batch_per_gpu=B, frames=spec.receptive_field, joints=17,
                  stage=wl["stage"], parallelism=f"dp{world} (sequence shards, one all-gather of outputs)" if world > 1 else "single GPU",
                  pipeline="steps submitted to the plan's two lanes (r3d_submit_uv / r3d_join), two batches in flight",
                  l2="4 rotating input sets (133 MB) + 129 MB weights + >1 GB activations per step, all larger than the 126 MB L2")

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        v, cores, ms = cpu_reference_throughput(wl, spec, args.cpu_sample, args.steps, args.warmup)
        print(json.dumps({
            "impl": "reference", "metric": "sequences/sec (T=%d, 17 joints)" % spec.receptive_field, "value": v, "unit": "sequences/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": v, "unit": "sequences/s", "cores": cores, "kind": "port",
                             "sample": f"{args.cpu_sample} sequences of the same workload per step (oracle port of the reference "
                                       f"modules' torch CPU ops, numpy float64 ray encode included), {cores} threads"},
            "e2e": {"value": v, "unit": "sequences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ------------------------------------------------------------------ our arm (CUDA)
    import torch.distributed as dist
    from ray3d_b200 import Lifter
    from ray3d_b200 import dist as rdist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (ray3d_b200 has no CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # stdout carries exactly one JSON line: anything libraries print there (NCCL's version banner at communicator
    # creation) is sent to stderr until the line is written
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    sp, st = synth.make_state_dicts(spec)
    lifter = Lifter(spec, sp, st, precision=precision, device=local_rank)
    NSETS = 4
    sets = []
    for i in range(NSETS):
        uv, cam = synth.make_inputs(spec, B, seed=1234 + 1 + 17 * i + 1000 * rank, res=wl["res"])
        sets.append((torch.from_numpy(uv), torch.from_numpy(cam)))
    dsets = [(u.to(dev), c.to(dev)) for u, c in sets]
    hsets = [(u.pin_memory(), c.pin_memory()) for u, c in sets]
    total_B = B * world
    # N > 1: the (B, J+1, 3) results of every step are all-gathered; the collective of step i runs on NCCL's stream
    # under the kernels of step i+1 (two buffer pairs), and the timed region ends only after the last one has landed
    gatherer = rdist.OverlappedGather(B, spec.num_joints, dev, depth=2) if world > 1 else None

    # Steps are independent batches: each is submitted to one of the plan's two lanes (r3d_submit_uv, alternating) and
    # joined one step later, so two batches are in flight and the under-filled tail launches of one (upper tree levels,
    # FC heads) run beside the large launches of the next.  Every step is complete inside the timed region (drain()).
    inflight = []
    last = {}

    def finish(pend):
        pos, trj, both = lifter.join(pend)
        if world > 1:
            gatherer.submit(both, trj)
        last["out"] = both

    def step(i):
        uv, cam = dsets[i % NSETS]
        inflight.append(lifter.submit_uv(uv, cam, want_pos=False))
        if len(inflight) > 1:
            finish(inflight.pop(0))

    def drain():
        while inflight:
            finish(inflight.pop(0))

    def step_serial(i):          # one stream, one batch at a time (per-launch profiling pass)
        uv, cam = dsets[i % NSETS]
        return lifter.forward_uv(uv, cam, want_pos=False)[2]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    drain()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.time()
    e0.record()
    for i in range(args.steps):
        step(i)
    drain()                       # the current stream waits for both lanes
    out = last["out"]
    if world > 1:
        gatherer.drain()          # ... and for the collectives still in flight
    e1.record()
    barrier()
    t1 = time.time()
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    ms_step = ms_total / args.steps
    # Per-launch durations for the roofline: the SAME steps once more with CUDA events recorded around every
    # kernel on its launch stream.  Kept out of the region that produces `value` because an event between two
    # launches defeats their programmatic-dependent-launch overlap (this pass is therefore slightly slower).
    lifter.plan.set_profiling(True)
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prof_steps = min(args.steps, 64)                     # the C ABI keeps the last 64 profiled forwards
    pe0.record()
    for i in range(prof_steps):
        step_serial(i)
    pe1.record()
    barrier()
    ms_step_profiled = pe0.elapsed_time(pe1) / prof_steps
    launch_times, runs = lifter.plan.launch_times()
    lifter.plan.set_profiling(False)
    value = total_B / ms_step * 1e3

    # -------- end to end through the host-buffer C-ABI calls (H2D + kernels + D2H of every step inside the timed region)
    # (1) synchronous call: returns with the step's results in host memory (chunks of the batch overlap inside the call);
    # (2) streaming submit/wait, depth 2: step i+1's H2D copy overlaps step i's kernels (what a loader thread does).
    outs = [torch.empty((B, 1, spec.num_joints, 3), dtype=torch.float32).pin_memory() for _ in range(3)]
    for i in range(3):
        lifter.forward_uv_host(*hsets[i % NSETS], out=outs[i % 3])
    barrier()
    e2e_steps = min(args.steps, 100)

    def reduce_max_ms(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    w0 = time.perf_counter()
    for i in range(e2e_steps):
        lifter.forward_uv_host(*hsets[i % NSETS], out=outs[i % 3])      # returns after results are in host memory
    if world > 1:
        dist.barrier()
    e2e_sync_ms = reduce_max_ms((time.perf_counter() - w0) * 1e3 / e2e_steps)

    DEPTH = 3
    pending = []
    checksum = 0.0
    barrier()
    w0 = time.perf_counter()
    for i in range(e2e_steps):
        if len(pending) == DEPTH:
            tk, ob = pending.pop(0)
            lifter.wait(tk)
            checksum += float(ob[0, 0, 0, 2])                           # the step's result, read from host memory
        ob = outs[i % 3]
        pending.append((lifter.submit_uv_host(*hsets[i % NSETS], out=ob), ob))
    for tk, ob in pending:
        lifter.wait(tk)
        checksum += float(ob[0, 0, 0, 2])
    if world > 1:
        dist.barrier()
    e2e_ms = reduce_max_ms((time.perf_counter() - w0) * 1e3 / e2e_steps)
    assert checksum == checksum, "NaN in the streamed results"
    h2d = B * (spec.receptive_field * 17 * 2 + 6) * 4
    d2h = B * 17 * 3 * 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # -------- roofline of the dominant kernel (grouped GEMM launches of the step)
    peaks = read_peaks()
    graph = lifter.plan.describe()
    flops = [0.0] + [sum(2.0 * B * o["rows_per_seq"] * (q["n"] * q.get("alg_k", q["k"]) + q.get("n2", 0) * q.get("k2", 0)) for q in o["prob"]) for o in graph["ops"]] + [0.0]
    issue_mult = 3.0 if precision == "bf16x3" else 1.0
    per = [dict(name=n, ms=ms, gflop=f / 1e9) for (n, ms), f in zip(launch_times, flops)]
    gemms = [p for p in per if p["gflop"] > 0]
    top = max(gemms, key=lambda p: p["ms"])
    gemm_ms, gemm_fl = sum(p["ms"] for p in gemms), sum(p["gflop"] for p in gemms)
    peak = peaks["bf16_tflops_sustained"] if precision != "fp32" else 74.4
    # The dominant kernel is the grouped GEMM (gemm_tc_kernel / gemm_ffma_kernel): ~91 % of the step, launched once per
    # layer.  achieved = algorithmic flops of all its launches in a step / their summed CUDA-event durations, i.e.
    # flops per launch / average launch duration; the largest single launch is listed beside it.
    achieved = gemm_fl / gemm_ms                      # TFLOP/s (GFLOP / ms)
    traffic = None
    top_traffic = None
    tpath = os.path.join(ROOT, "profiles", "r1_dominant_kernel_traffic.json")
    if os.path.exists(tpath) and args.workload == "cfg2" and precision == "bf16x3":
        with open(tpath) as f:
            tj = json.load(f)
        top_traffic = tj.get(top["name"], {}).get("dram_bytes")
        # per-launch average over the captured GEMM launches (the three largest, 55 % of the GEMM time)
        cap = [v["dram_bytes"] for k, v in tj.items() if k != "input_stage"]
        traffic = sum(cap) / len(cap) if cap else None
    alg_bytes = (spec.receptive_field * 17 * 2 * 4 + 24 + 216) + lifter.plan.weight_bytes / B
    roofline = {
        "bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05 grouped GEMM, all %d launches of a step)" % len(gemms) if precision != "fp32" else "gemm_ffma_kernel (all launches of a step)",
        "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
        "traffic_note": "mean DRAM read+write bytes per launch over the ncu --set full captures of the 3 largest GEMM launches (profiles/r1_dominant_kernel_traffic.json)",
        "peak_source": peaks["source"] + (", sustained bf16 cuBLAS figure (kernel timed inside a long step)" if precision != "fp32" else "; fp32 FFMA nominal 148 SM x 128 x 2 x 1.965 GHz"),
        "note": ("achieved counts ALGORITHMIC fp32 flops (2*M*N*K of the reference's conv/linear shapes); the bf16x3 path issues 3 bf16 MMAs "
                 "per algorithmic MAC, so the tensor-pipe issue fraction is 3x frac" if precision == "bf16x3" else "algorithmic flops"),
        "tensor_issue_frac": achieved * issue_mult / peak if precision != "fp32" else None,
        "gemm_launches_per_step": len(gemms), "gemm_gflop_per_step": gemm_fl, "gemm_ms_per_step": gemm_ms,
        "gemm_share_of_step": gemm_ms / ms_step_profiled,
        "top_launch": {"name": top["name"], "ms": top["ms"], "gflop": top["gflop"], "achieved": top["gflop"] / top["ms"],
                       "frac": top["gflop"] / top["ms"] / peak, "share_of_step": top["ms"] / ms_step_profiled, "traffic": top_traffic},
        "ms_per_step_with_launch_events": ms_step_profiled,
        "hbm": {"algorithmic_bytes_per_seq": alg_bytes, "achieved_gbs": value / world * alg_bytes / 1e9, "peak_gbs": peaks["hbm_gbs"],
                "frac": value / world * alg_bytes / 1e9 / peaks["hbm_gbs"],
                "note": "arithmetic intensity ~1350 flop/B: the path is tensor-bound, the HBM fraction is reported as asked"},
        "launches": per, "profiled_steps": runs,
    }

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        v, cores, ms = cpu_reference_throughput(wl, spec, args.cpu_sample, 3, 1)
        cpu_baseline = {"value": v, "unit": "sequences/s", "cores": cores, "kind": "port",
                        "sample": f"3 timed passes over {args.cpu_sample} sequences of the same workload (oracle port, {cores} torch threads, "
                                  f"numpy float64 ray encode included), {ms:.0f} ms per pass"}

    # parity spot check on the benchmark inputs (not timed): first 4 sequences vs the oracle in float64
    from oracle import ray3d_oracle as O
    uv0, cam0 = sets[(args.steps - 1) % NSETS]
    ref = O.lift_uv(O.to_torch_state(sp, torch.float64), O.to_torch_state(st, torch.float64), spec, uv0[:4].numpy(), cam0[:4].numpy())[2].numpy()
    got = out[:4].cpu().numpy()
    relerr = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))

    line = {
        "metric": "sequences/sec (T=%d, 17 joints)" % spec.receptive_field, "value": value, "unit": "sequences/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32", "bf16x3": "f32 operands as bf16 hi+lo (3 tensor-core products), f32 accumulate", "bf16": "bf16, f32 accumulate"}[precision],
        "data": "synthetic (seeded uv + intrinsics, seeded reference-shaped weights)", "config": config,
        "clocks": clocks, "e2e": {"value": total_B / e2e_ms * 1e3, "unit": "sequences/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                                  "ms_per_step": e2e_ms, "steps": e2e_steps,
                                  "api": "Lifter.submit_uv_host / wait -> r3d_submit_uv_host / r3d_wait (pinned host buffers, 3 submissions in flight on the "
                                         "plan's two lanes: each step's H2D copy and tail launches overlap the neighbouring steps' kernels)",
                                  "sync_value": total_B / e2e_sync_ms * 1e3, "sync_ms_per_step": e2e_sync_ms,
                                  "sync_api": "Lifter.forward_uv_host -> r3d_forward_uv_host (one blocking call per step; inside it two 512-window chunks "
                                              "alternate between the lanes, the second chunk's copy under the first chunk's kernels)"},
        "gpu_launches": lifter.plan.kernel_launches * args.steps, "launches_per_step": lifter.plan.kernel_launches,
        "roofline": roofline, "cpu_baseline": cpu_baseline, "parity_relerr_vs_oracle_f64": relerr,
        "flops_per_sequence": flops_per_sequence(spec), "achieved_tflops_step": flops_per_sequence(spec) * value / 1e12,
        "precision": precision,
    }
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    if world > 1:
        os.dup2(2, 1)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

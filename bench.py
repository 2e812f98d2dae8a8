#!/usr/bin/env python
"""Benchmark of the Ray3D lifting hot path (BASELINE.json metric: sequences/sec, T=243, 17 joints).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # reference algorithm on the host CPU cores

A "step" is one pass of the hot path (camera-ray encode -> pose net + trajectory net -> pos + trj) over one
batch of synthetic sequences: BASELINE.json configs[1] (batch 1024, T=243, 17 joints, fp32 parity bar) per
GPU.  Under torchrun every rank lifts its own 1024-sequence shard (weak scaling) and the step ends with the
single all-gather of the outputs.  One JSON line is printed by rank 0; besides the contract's keys it carries
`extra` (the other BASELINE.json configurations measured in the same run: strict fp32, cfg3, cfg5), `e2e.video`
(whole-video evaluation from pixels through host buffers), `e2e.module_api` (the unmodified caller's
Model(...).get_pos_model()(x, param) loop) and, for N > 1, `gather_check` (the all-gathered tensor verified on
every rank).
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
                 widths=(3, 3, 3, 3, 3), stage=1, batch=1024, res=1000, precision="bf16x3"),
    "cfg3": dict(desc="batch=4096, T=81, 17 joints, stage 1, bf16", widths=(3, 3, 3, 3), stage=1, batch=4096, res=1000, precision="bf16"),
    "cfg5": dict(desc="cfg_ray3d_3dhp_stage3 arch, T=243, batch=512/GPU", widths=(3, 3, 3, 3, 3), stage=3, batch=512, res=2048,
                 precision="bf16x3"),
}
DTYPE = {"fp32": "f32", "bf16x3": "f32 operands as bf16 hi+lo (3 tensor-core products), f32 accumulate", "bf16": "bf16, f32 accumulate"}
TOL = {"fp32": 2e-6, "bf16x3": 1e-4, "bf16": 2e-2}          # normwise relative error bound vs the float64 oracle
FFMA_PEAK_TFLOPS = 74.4                                      # 148 SM x 128 lanes x 2 x 1.965 GHz (nominal)


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


def bind_to_gpu_numa_node(local_rank: int):
    """Pin this process to the CPU cores of the NUMA node its GPU hangs off (before any pinned buffer is allocated, so
    first-touch places the staging memory next to the GPU's PCIe root).  Best effort; returns what was done."""
    try:
        bdf = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local_rank)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if bdf.startswith("0000"):
            bdf = bdf[4:]                                      # nvidia-smi prints an 8-digit domain, sysfs a 4-digit one
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return {"node": None, "note": "single NUMA node / not reported"}
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"node": node, "cpus": len(cpus)}
    except Exception as e:                                     # containers without sysfs topology, no permission, ...
        return {"node": None, "note": f"not bound ({type(e).__name__})"}


def cpu_reference_throughput(wl, spec, sample: int, steps: int, warmup: int):
    """Times the oracle port of the reference's eval step (ray encode in numpy float64 + the same torch CPU ops the
    reference's modules dispatch, in-place activations included) on all host threads.  Returns (seq/s, cores, ms/step)."""
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


def oracle_relerr(spec, sp, st, uv, cam, got) -> float:
    """normwise relative error of `got` (pos+trj) vs the float64 oracle on the same windows"""
    from oracle import ray3d_oracle as O
    ref = O.lift_uv(O.to_torch_state(sp, torch.float64), O.to_torch_state(st, torch.float64), spec, np.asarray(uv), np.asarray(cam))[2].numpy()
    return float(np.linalg.norm(np.asarray(got, dtype=np.float64) - ref) / np.linalg.norm(ref))


class Run:
    """One workload on this rank: plan, rotating device input sets, the two-lane step loop."""
    NSETS = 4

    def __init__(self, wl, precision, rank, world, local_rank):
        from ray3d_b200 import Lifter, NetSpec, synth
        from ray3d_b200 import dist as rdist
        self.wl, self.precision, self.rank, self.world = wl, precision, rank, world
        self.spec = NetSpec(num_joints=17, in_features=3, filter_widths=wl["widths"], stage=wl["stage"])
        self.dev = torch.device("cuda", local_rank)
        self.B = wl["batch"]
        self.sp, self.st = synth.make_state_dicts(self.spec)
        # plan build options for A/B runs, e.g. R3D_BENCH_OPTIONS="tail_fusion=0" (default: the library's defaults)
        opts = {k: int(v) for k, v in (kv.split("=") for kv in os.environ.get("R3D_BENCH_OPTIONS", "").split(",") if kv)}
        self.lifter = Lifter(self.spec, self.sp, self.st, precision=precision, device=local_rank, options=opts)
        self.sets = []
        for i in range(self.NSETS):
            uv, cam = synth.make_inputs(self.spec, self.B, seed=self.seed(i, rank), res=wl["res"])
            self.sets.append((torch.from_numpy(uv), torch.from_numpy(cam)))
        self.dsets = [(u.to(self.dev), c.to(self.dev)) for u, c in self.sets]
        # N > 1: the (B, J+1, 3) results of every step are all-gathered; the collective of step i runs on NCCL's stream
        # under the kernels of step i+1 (two buffer pairs), and the timed region ends only after the last one has landed
        self.gatherer = rdist.OverlappedGather(self.B, self.spec.num_joints, self.dev, depth=2) if world > 1 else None
        self.inflight, self.last = [], {}

    @staticmethod
    def seed(i, rank):
        return 1234 + 1 + 17 * i + 1000 * rank

    # Steps are independent batches: each is submitted to one of the plan's two lanes (r3d_submit, alternating) and
    # joined one step later, so two batches are in flight and the under-filled tail launches of one (upper tree levels,
    # FC heads) run beside the large launches of the next.  Every step is complete inside the timed region (drain()).
    def _finish(self, pend, idx):
        pos, trj, both = self.lifter.join(pend)
        if self.world > 1:
            self.last["ticket"] = self.gatherer.submit(both, trj)
        self.last["out"], self.last["trj"], self.last["set"] = both, trj, idx

    def step(self, i):
        uv, cam = self.dsets[i % self.NSETS]
        self.inflight.append((self.lifter.submit_uv(uv, cam, want_pos=False), i % self.NSETS))
        if len(self.inflight) > 1:
            self._finish(*self.inflight.pop(0))

    def drain(self):
        while self.inflight:
            self._finish(*self.inflight.pop(0))

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, ms: float) -> float:
        if self.world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    def timed(self, steps, warmup, sampler=None):
        """W untimed steps, then exactly K steps between barrier + synchronize pairs, CUDA events, max over ranks."""
        for i in range(warmup):
            self.step(i)
        self.drain()
        self.barrier()
        if sampler is not None:
            sampler.start()
            time.sleep(0.25)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        t0 = time.time()
        e0.record()
        for i in range(steps):
            self.step(i)
        self.drain()                       # the current stream waits for both lanes
        if self.world > 1:
            self.gatherer.drain()          # ... and for the collectives still in flight
        e1.record()
        self.barrier()
        t1 = time.time()
        ms_total = self.max_over_ranks(e0.elapsed_time(e1))
        return ms_total / steps, (t0, t1)

    def parity(self, n=4) -> float:
        """first n sequences of the last step's local output vs the float64 oracle (not timed)"""
        uv0, cam0 = self.sets[self.last["set"]]
        return oracle_relerr(self.spec, self.sp, self.st, uv0[:n].numpy(), cam0[:n].numpy(), self.last["out"][:n].cpu().numpy())

    def gather_check(self):
        """N > 1: the all-gathered tensor of the last step, on EVERY rank: this rank's own rows are bit-identical to its
        local output, and a FOREIGN rank's rows match the float64 oracle on that rank's (regenerated) inputs."""
        from ray3d_b200 import synth
        from ray3d_b200 import dist as rdist
        import torch.distributed as dist
        full = self.gatherer.result(self.last["ticket"])
        both, trj = rdist.unpack_outputs(full)
        lo = self.rank * self.B
        own_ok = bool(torch.equal(both[lo:lo + self.B], self.last["out"]) and torch.equal(trj[lo:lo + self.B], self.last["trj"]))
        other = (self.rank + 1) % self.world
        uv_o, cam_o = synth.make_inputs(self.spec, self.B, seed=self.seed(self.last["set"], other), res=self.wl["res"])
        rel = oracle_relerr(self.spec, self.sp, self.st, uv_o[:2], cam_o[:2], both[other * self.B: other * self.B + 2].cpu().numpy())
        ok = own_ok and rel < TOL[self.precision]
        flags = torch.tensor([1.0 if ok else 0.0, rel], dtype=torch.float64, device=self.dev)
        worst = flags.clone()
        dist.all_reduce(flags[0:1], op=dist.ReduceOp.MIN)
        dist.all_reduce(worst[1:2], op=dist.ReduceOp.MAX)
        return dict(all_ranks_ok=bool(flags[0].item() == 1.0), own_rows_bit_identical=own_ok, foreign_rank=other,
                    foreign_relerr_vs_oracle_f64=rel, worst_foreign_relerr_over_ranks=float(worst[1].item()), tolerance=TOL[self.precision],
                    rows=int(full.shape[0]))


def gemm_roofline(run: Run, ms_step: float, timed_seconds: float, peaks, profile_steps: int):
    """Per-launch durations of the SAME steps with CUDA events recorded around every kernel on its launch stream.  Kept
    out of the region that produces `value` because an event between two launches defeats their programmatic-dependent
    -launch overlap and the side stream / second lane are off (this pass is therefore slower than the timed region)."""
    lf, B, precision = run.lifter, run.B, run.precision
    lf.plan.set_profiling(True)
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = min(profile_steps, 64)                            # the C ABI keeps the last 64 profiled forwards
    pe0.record()
    for i in range(n):
        uv, cam = run.dsets[i % run.NSETS]
        lf.forward_uv(uv, cam, want_pos=False)
    pe1.record()
    run.barrier()
    ms_prof = pe0.elapsed_time(pe1) / n
    launch_times, runs = lf.plan.launch_times()
    lf.plan.set_profiling(False)
    graph = lf.plan.describe()
    op_flops = [sum(2.0 * B * o["rows_per_seq"] * (q["n"] * q.get("alg_k", q["k"]) + q.get("n2", 0) * q.get("k2", 0)) for q in o["prob"]) for o in graph["ops"]]
    # one entry per kernel launch: the chained tail launch covers many ops (graph["launches"][k]["ops"])
    flops = [0.0] + [sum(op_flops[i] for i in L["ops"]) for L in graph["launches"]] + [0.0]
    assert len(flops) == len(launch_times), (len(flops), len(launch_times))
    per = [dict(name=nm, ms=ms, gflop=f / 1e9) for (nm, ms), f in zip(launch_times, flops)]
    for p, L in zip(per[1:-1], graph["launches"]):
        p["chained"] = bool(L["tail"])
    # The dominant kernel is gemm_tc_kernel (one launch per layer / fused layer pair).  A chained launch (tail_tc_kernel: the
    # GlobalInfo layers on a few CTA pairs of the side stream) is built to run BESIDE those launches on a sliver of the
    # GPU; serialised for this pass it occupies its 13 CTA pairs for the whole chain, so its duration is listed but
    # neither its time nor its flops enter `achieved`.  achieved_timed_region counts every flop against the whole step.
    all_gemms = [p for p in per if p["gflop"] > 0]
    gemms = [p for p in all_gemms if not p.get("chained")]
    chained = [p for p in all_gemms if p.get("chained")]
    top = max(gemms, key=lambda p: p["ms"])
    gemm_ms, gemm_fl = sum(p["ms"] for p in gemms), sum(p["gflop"] for p in gemms)
    all_fl = sum(p["gflop"] for p in all_gemms)
    tc = precision != "fp32"
    burst, sustained = (peaks["bf16_tflops"], peaks["bf16_tflops_sustained"]) if tc else (FFMA_PEAK_TFLOPS, FFMA_PEAK_TFLOPS)
    # burst figure for a short timed region (the board never reaches its power cap), sustained for a seconds-long one
    peak, which = (burst, "burst") if timed_seconds < 1.0 else (sustained, "sustained")
    achieved = gemm_fl / gemm_ms                          # TFLOP/s: algorithmic flops of all GEMM launches / their summed durations
    achieved_timed = all_fl / ms_step                    # ... / the timed region's step (lanes + side stream overlapped, every other kernel included)
    issue = 3.0 if precision == "bf16x3" else 1.0
    traffic = top_traffic = None
    for name in ("r2_dominant_kernel_traffic.json", "r1_dominant_kernel_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tpath) and run.wl is WORKLOADS["cfg2"] and precision == "bf16x3":
            with open(tpath) as f:
                tj = json.load(f)
            top_traffic = tj.get(top["name"], {}).get("dram_bytes")
            cap = [tj[p["name"]]["dram_bytes"] for p in gemms if p["name"] in tj]
            traffic = sum(cap) / len(cap) if cap else None
            traffic_src = name
            break
    else:
        traffic_src = None
    alg_bytes = (run.spec.receptive_field * 17 * 2 * 4 + 24 + 216) + lf.plan.weight_bytes / B
    return {
        "bound": "tensor",
        "kernel": ("gemm_tc_kernel (tcgen05 grouped GEMM / fused conv pair; all %d launches of a step)" % len(gemms)) if tc
                  else "gemm_ffma_kernel (all launches of a step)",
        "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "peak_kind": which,
        "frac_of_burst_peak": achieved / burst, "frac_of_sustained_peak": achieved / sustained,
        "achieved_timed_region": achieved_timed, "frac_timed_region": achieved_timed / peak,
        "traffic": traffic,
        "traffic_note": f"mean DRAM read+write bytes per launch over the ncu --set full captures of the same gemm_tc_kernel launches (profiles/{traffic_src}); "
                        "not re-measured inside this run" if traffic_src else None,
        "peak_source": peaks["source"] + (f", {which} bf16 cuBLAS figure (timed region {timed_seconds:.2f} s)" if tc
                                          else "; fp32 FFMA nominal 148 SM x 128 x 2 x 1.965 GHz"),
        "note": ("achieved counts ALGORITHMIC fp32 flops (2*M*N*K of the reference's conv/linear shapes) over the summed CUDA-event durations of "
                 "the GEMM launches in a serialised profiling pass; achieved_timed_region divides the same flops by the timed step. "
                 + ("The bf16x3 path issues 3 bf16 MMAs per algorithmic MAC: tensor_issue_frac = 3 x frac." if precision == "bf16x3" else "")),
        "tensor_issue_frac": achieved * issue / peak if tc else None,
        "tensor_issue_frac_timed_region": achieved_timed * issue / peak if tc else None,
        "gemm_launches_per_step": len(gemms), "gemm_gflop_per_step": gemm_fl, "gemm_ms_per_step": gemm_ms,
        "gemm_share_of_step": gemm_ms / ms_prof, "all_gemm_gflop_per_step": all_fl,
        "chained_launches": [dict(name=p["name"], ms=p["ms"], gflop=p["gflop"], kernel="tail_tc_kernel",
                                  note="runs on a few CTA pairs beside the per-layer launches; excluded from achieved/frac, "
                                       "included in achieved_timed_region") for p in chained],
        "top_launch": {"name": top["name"], "ms": top["ms"], "gflop": top["gflop"], "achieved": top["gflop"] / top["ms"],
                       "frac": top["gflop"] / top["ms"] / peak, "tensor_issue_frac": top["gflop"] / top["ms"] * issue / peak if tc else None,
                       "share_of_step": top["ms"] / ms_prof, "traffic": top_traffic},
        "ms_per_step_with_launch_events": ms_prof,
        "hbm": {"algorithmic_bytes_per_seq": alg_bytes, "achieved_gbs": B * 1e3 / ms_step * alg_bytes / 1e9, "peak_gbs": peaks["hbm_gbs"],
                "frac": B * 1e3 / ms_step * alg_bytes / 1e9 / peaks["hbm_gbs"],
                "note": "arithmetic intensity ~1350 flop/B: the path is tensor-bound, the HBM fraction is reported as asked"},
        "launches": per, "profiled_steps": runs,
    }


def e2e_host(run: Run, steps: int):
    """End to end through the host-buffer C-ABI calls: H2D + kernels + D2H of every step inside the timed region.
    (1) synchronous call: returns with the step's results in host memory (chunks of the batch overlap inside the call);
    (2) streaming submit/wait, depth 3: step i+1's H2D copy overlaps step i's kernels (what a loader thread does)."""
    lf, B, spec = run.lifter, run.B, run.spec
    hsets = [(u.pin_memory(), c.pin_memory()) for u, c in run.sets]
    outs = [torch.empty((B, 1, spec.num_joints, 3), dtype=torch.float32).pin_memory() for _ in range(3)]
    # the host-side ceiling of this path: plain pinned H2D copies of the same buffers, every rank at once (no kernels)
    dst = torch.empty_like(hsets[0][0], device=run.dev)
    for i in range(8):
        dst.copy_(hsets[i % run.NSETS][0], non_blocking=True)
    run.barrier()
    w0 = time.perf_counter()
    for i in range(40):
        dst.copy_(hsets[i % run.NSETS][0], non_blocking=True)
    torch.cuda.synchronize()
    copy_gbs = 40 * hsets[0][0].numel() * 4 / (time.perf_counter() - w0) / 1e9
    run.barrier()
    copy_gbs = -run.max_over_ranks(-copy_gbs)                            # slowest rank
    del dst
    for i in range(3):
        lf.forward_uv_host(*hsets[i % run.NSETS], out=outs[i % 3])
    run.barrier()
    w0 = time.perf_counter()
    for i in range(steps):
        lf.forward_uv_host(*hsets[i % run.NSETS], out=outs[i % 3])      # returns after results are in host memory
    run.barrier()
    sync_ms = run.max_over_ranks((time.perf_counter() - w0) * 1e3 / steps)

    DEPTH, pending, checksum = 3, [], 0.0
    for i in range(DEPTH):                                              # warm the streaming path as well (staging slots of its chunk size)
        lf.wait(lf.submit_uv_host(*hsets[i % run.NSETS], out=outs[i % 3]))
    run.barrier()
    dbg = os.environ.get("R3D_BENCH_DEBUG") == "1"
    stamps = []
    w0 = time.perf_counter()
    for i in range(steps):
        if len(pending) == DEPTH:
            tk, ob = pending.pop(0)
            lf.wait(tk)
            checksum += float(ob[0, 0, 0, 2])                           # the step's result, read from host memory
        ob = outs[i % 3]
        pending.append((lf.submit_uv_host(*hsets[i % run.NSETS], out=ob), ob))
        if dbg:
            stamps.append(time.perf_counter() - w0)
    for tk, ob in pending:
        lf.wait(tk)
        checksum += float(ob[0, 0, 0, 2])
    own_ms = (time.perf_counter() - w0) * 1e3 / steps
    if dbg:
        print("e2e streaming per-step stamps (ms):", " ".join(f"{t * 1e3:.2f}" for t in stamps), file=sys.stderr)
    run.barrier()
    ms = run.max_over_ranks(own_ms)
    assert checksum == checksum, "NaN in the streamed results"
    h2d = B * (spec.receptive_field * 17 * 2 + 6) * 4
    return dict(ms=ms, sync_ms=sync_ms, h2d=h2d, d2h=B * 17 * 3 * 4, own_h2d_gbs=h2d / own_ms / 1e6, copy_gbs=copy_gbs)


def e2e_video(run: Run, steps: int):
    """Whole-video evaluation from pixels through host buffers (r3d_submit_video_uv_host): what Trainer.evaluate_core does
    per video (trainer.py:297-356) -- eval_data_prepare, np.tile, .cuda(), forward, .cpu() -- with the F+RF-1 frames
    crossing PCIe once (136 B per frame) instead of RF-fold inflated windows.  One step = one video of F = batch frames
    -> batch sequences; same weights, same windows-per-step as the headline workload."""
    from ray3d_b200 import RayCamera
    lf, B, spec, T = run.lifter, run.B, run.spec, run.spec.receptive_field
    rng = np.random.default_rng(77 + run.rank)
    K = np.array([[1145.0494384765625 + 0.01234, 0.0, 512.54150390625], [0.0, 1143.7811279296875 + 0.0077, 515.4514770507812], [0.0, 0.0, 1.0]])
    c, s = np.cos(0.2), np.sin(0.2)
    cam = RayCamera(K, np.array([[1.0, 0.0, 0.0], [0.0, -s, -c], [0.0, c, -s]]), np.array([0.1, 1.6 * c, 1.6 * s]), res_w=1000, res_h=1002)
    row = torch.from_numpy(cam.table_row64()).pin_memory()
    vids = []
    for i in range(3):
        walk = np.cumsum(rng.normal(0, 2.0, size=(B + T - 1, 1, 2)), axis=0) + rng.uniform(300, 700, size=(1, 1, 2))
        uv = (walk + rng.normal(0, 60.0, size=(1, 17, 2)) + rng.normal(0, 1.5, size=(B + T - 1, 17, 2))).astype(np.float32)
        vids.append(torch.from_numpy(uv).pin_memory())
    outs = [torch.empty((B, 1, 17, 3), dtype=torch.float32).pin_memory() for _ in range(3)]
    for i in range(3):
        lf.wait(lf.submit_video_uv_host(vids[i], row, outs[i]))
    # parity of the video path on this input: first/last window vs the float64 oracle on materialised windows
    win = np.stack([vids[2][f:f + T].numpy() for f in (0, B - 1)])
    cam6 = np.tile(np.array([K[0, 0], K[1, 1], K[0, 2], K[1, 2], cam.cam_pitch_rad, cam.height]), (2, 1))
    rel = oracle_relerr(spec, run.sp, run.st, win.astype(np.float64), cam6, outs[2][[0, B - 1]].numpy())
    run.barrier()
    DEPTH, pending = 3, []
    w0 = time.perf_counter()
    for i in range(steps):
        if len(pending) == DEPTH:
            lf.wait(pending.pop(0))
        pending.append(lf.submit_video_uv_host(vids[i % 3], row, outs[i % 3]))
    for tk in pending:
        lf.wait(tk)
    run.barrier()
    ms = run.max_over_ranks((time.perf_counter() - w0) * 1e3 / steps)
    return dict(value=B * run.world / ms * 1e3, unit="sequences/s", ms_per_step=ms, steps=steps, frames_per_video=B,
                h2d_bytes_per_step=(B + T - 1) * 17 * 2 * 4 + 128, d2h_bytes_per_step=B * 17 * 3 * 4, parity_relerr_vs_oracle_f64=rel,
                api="Lifter.submit_video_uv_host / wait -> r3d_submit_video_uv_host / r3d_wait (pinned host buffers, 3 videos in flight); "
                    "float64 camera row; every frame ray-encoded once on the device, windows indexed in place")


def e2e_module_api(run: Run, steps: int):
    """The unmodified caller's loop (trainer.py:329-356) on the drop-in modules: Model(cfg).get_pos_model() /
    get_trj_model(), inputs_2d.cuda(), pos(x, param), trj(x, param), pos += trj, .cpu() -- ray-encoded float32 windows in
    pinned host memory (the reference encodes its dataset once, lib/dataset/__init__.py:191-203)."""
    import ray3d_b200
    from oracle import ray3d_oracle as O
    spec, B, wl = run.spec, run.B, run.wl
    cfg = {'MODEL': 'RIE', 'ARCHITECTURE': ",".join(str(w) for w in wl["widths"]), 'DROPOUT': 0.2, 'CAUSAL': False, 'CHANNELS': 256,
           'DENSE': False, 'NUM_KPTS': 17, 'INPUT_DIM': 3, 'CAMERA_EMBDDING': True, 'EXTRINSIC_DIM': 2, 'EMBEDD_DIM': 64,
           'LATENT_FEATURES_DIM': 256, 'DISABLE_OPTIMIZATIONS': False, 'STAGE': wl["stage"], 'TRAJECTORY_MODEL': True}
    os.environ["RAY3D_B200_PRECISION"] = run.precision
    m = ray3d_b200.Model(cfg, None, is_train=False)
    pos_m, trj_m = m.get_pos_model(), m.get_trj_model()
    pos_m.load_state_dict({"module." + k: torch.from_numpy(np.asarray(v)) for k, v in run.sp.items()}, strict=True)
    trj_m.load_state_dict({"module." + k: torch.from_numpy(np.asarray(v)) for k, v in run.st.items()}, strict=True)
    pos_m.eval(); trj_m.eval()
    hx, hp = [], []
    for u, c in run.sets[:2]:
        hx.append(torch.from_numpy(O.ray_encode_batch(u.numpy(), c.numpy())).pin_memory())
        hp.append(torch.from_numpy(np.ascontiguousarray(c.numpy()[:, [5, 4]])).pin_memory())

    def one(i):
        with torch.no_grad():
            x, prm = hx[i % 2].cuda(), hp[i % 2].cuda()              # trainer.py:329-334
            pos = pos_m(x, prm)                                        # :337
            trj = trj_m(x, prm)                                        # :346
            pos += trj                                                 # :353
            return pos.cpu()                                           # :355

    for i in range(3):
        got = one(i)
    rel = oracle_relerr(spec, run.sp, run.st, run.sets[0][0][:4].numpy(), run.sets[0][1][:4].numpy(), got[:4].numpy())
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    for i in range(steps):
        one(i)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - w0) * 1e3 / steps
    del m, pos_m, trj_m
    return dict(value=B / ms * 1e3, unit="sequences/s", ms_per_step=ms, steps=steps, h2d_bytes_per_step=B * (spec.receptive_field * 17 * 3 + 2) * 4,
                d2h_bytes_per_step=B * 17 * 3 * 4, parity_relerr_vs_oracle_f64=rel,
                api="ray3d_b200.Model(cfg).get_pos_model()(x, param); get_trj_model()(x, param); pos += trj with .cuda()/.cpu() around it "
                    "(blocking, one batch at a time; the two modules share one native plan)")


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
    ap.add_argument("--no-extra", action="store_true", help="skip the fp32 / cfg3 / cfg5 / video / module-API sub-measurements")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    from ray3d_b200 import NetSpec
    from ray3d_b200.spec import flops_per_sequence

    wl = WORKLOADS[args.workload]
    spec = NetSpec(num_joints=17, in_features=3, filter_widths=wl["widths"], stage=wl["stage"])
    precision = args.precision or wl["precision"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    B = wl["batch"]
    metric = "sequences/sec (T=%d, 17 joints)" % spec.receptive_field

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        v, cores, ms = cpu_reference_throughput(wl, spec, args.cpu_sample, args.steps, args.warmup)
        sample = (f"{args.cpu_sample} sequences of the same workload per step (oracle port: the torch CPU ops the reference's modules "
                  f"dispatch, in-place activations like theirs, numpy float64 ray encode included), {cores} torch threads")
        print(json.dumps({
            "impl": "reference", "metric": metric, "value": v, "unit": "sequences/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic (seeded uv + intrinsics, seeded reference-shaped weights)",
            "config": {"workload": f"{args.workload}: {wl['desc']}", "frames": spec.receptive_field, "joints": 17, "stage": wl["stage"],
                       "batch_per_step": args.cpu_sample,
                       "pipeline": f"torch CPU eval forward, {cores} threads, one {args.cpu_sample}-sequence pass per step (a bounded sample of "
                                   f"the {B}-sequence batch; the reference feeds whatever batch its caller built, trainer.py:323-337)",
                       "note": "the reference itself is Python and cannot travel to the GPU box: `kind: port` = oracle/ray3d_oracle.py, "
                               "bit-identical to the reference's modules on the golden fixtures"},
            "cpu_baseline": {"value": v, "unit": "sequences/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "sequences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ------------------------------------------------------------------ our arm (CUDA)
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (ray3d_b200 has no CPU path)")
    numa = bind_to_gpu_numa_node(local_rank)          # before any pinned allocation
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # stdout carries exactly one JSON line: anything libraries print there (NCCL's version banner at communicator
    # creation) is sent to stderr until the line is written
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    config = dict(workload=f"{args.workload}: {wl['desc']}", batch_per_gpu=B, frames=spec.receptive_field, joints=17,
                  stage=wl["stage"], parallelism=f"dp{world} (sequence shards, one all-gather of outputs)" if world > 1 else "single GPU",
                  pipeline="steps submitted to the plan's two lanes (r3d_submit / r3d_join), two batches in flight",
                  l2="4 rotating input sets (133 MB) + 129 MB weights + >0.5 GB activations per step, all larger than the 126 MB L2")

    run = Run(wl, precision, rank, world, local_rank)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_step, (t0, t1) = run.timed(args.steps, args.warmup, sampler)
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    value = B * world / ms_step * 1e3
    relerr = run.parity()
    gather = run.gather_check() if world > 1 else None
    if gather is not None and not gather["all_ranks_ok"]:
        raise SystemExit(f"rank {rank}: all-gathered output failed verification: {gather}")
    peaks = read_peaks()
    roofline = gemm_roofline(run, ms_step, ms_step * args.steps / 1e3, peaks, args.steps)

    e2e_steps = min(args.steps, 100)
    e = e2e_host(run, e2e_steps)
    e2e = {"value": B * world / e["ms"] * 1e3, "unit": "sequences/s", "h2d_bytes_per_step": e["h2d"], "d2h_bytes_per_step": e["d2h"],
           "ms_per_step": e["ms"], "steps": e2e_steps,
           "api": "Lifter.submit_uv_host / wait -> r3d_submit_host / r3d_wait (pinned host buffers, 3 submissions in flight on the "
                  "plan's two lanes: each step's H2D copy and tail launches overlap the neighbouring steps' kernels)",
           "sync_value": B * world / e["sync_ms"] * 1e3, "sync_ms_per_step": e["sync_ms"],
           "sync_api": "Lifter.forward_uv_host -> r3d_forward_host (one blocking call per step; inside it two 512-window chunks "
                       "alternate between the lanes, the second chunk's copy under the first chunk's kernels)",
           "rank0_h2d_gbs": e["own_h2d_gbs"], "numa_binding": numa,
           "h2d_ceiling": {"pinned_copy_gbs_per_rank_all_ranks_copying": e["copy_gbs"],
                           "bound_sequences_per_s": world * e["copy_gbs"] * 1e9 / (e["h2d"] / B),
                           "note": "plain cudaMemcpyAsync of the same pinned input buffers on every rank at once, slowest rank: what the host's memory "
                                   "system / PCIe root can feed; the window path ships RF-fold inflated input (33 KB per sequence), e2e.video ships 136 B"}}
    extra = {}
    if not args.no_extra:
        e2e["video"] = e2e_video(run, e2e_steps)
        if world == 1:
            e2e["module_api"] = e2e_module_api(run, min(args.steps, 30))

    launches_per_step = run.lifter.plan.kernel_launches
    weight_bytes = run.lifter.plan.weight_bytes
    del run
    torch.cuda.empty_cache()

    # -------- the other BASELINE.json configurations, measured in the same run (same loop, fewer steps)
    if not args.no_extra and args.workload == "cfg2":
        x_steps, x_warm = min(args.steps, 20), 3
        for name, xwl, xprec in (("fp32", WORKLOADS["cfg2"], "fp32"), ("cfg3", WORKLOADS["cfg3"], "bf16"), ("cfg5", WORKLOADS["cfg5"], "bf16x3")):
            if precision == xprec and xwl is wl:
                continue
            try:
                r = Run(xwl, xprec, rank, world, local_rank)
                steps_x = max(3, x_steps // 4) if xprec == "fp32" else x_steps       # fp32 FFMA steps are ~7 ms each
                ms_x, _ = r.timed(steps_x, x_warm)
                rl = gemm_roofline(r, ms_x, ms_x * steps_x / 1e3, peaks, min(steps_x, 8))
                extra[name] = dict(workload=xwl["desc"], precision=xprec, dtype=DTYPE[xprec], value=r.B * world / ms_x * 1e3, unit="sequences/s",
                                   ms_per_step=ms_x, steps=steps_x, warmup=x_warm, batch_per_gpu=r.B, n_gpus=world,
                                   parity_relerr_vs_oracle_f64=r.parity(), tolerance=TOL[xprec], flops_per_sequence=flops_per_sequence(r.spec),
                                   roofline={k: rl[k] for k in ("achieved", "peak", "peak_kind", "unit", "frac", "frac_timed_region", "top_launch")},
                                   gpu_launches=r.lifter.plan.kernel_launches * steps_x)
                del r
                torch.cuda.empty_cache()
            except Exception as ex:                                     # a sub-measurement must never take the headline line down
                extra[name] = {"error": f"{type(ex).__name__}: {ex}"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        v, cores, ms = cpu_reference_throughput(wl, spec, args.cpu_sample, 3, 1)
        cpu_baseline = {"value": v, "unit": "sequences/s", "cores": cores, "kind": "port",
                        "sample": f"3 timed passes over {args.cpu_sample} sequences of the same workload (oracle port, {cores} torch threads, "
                                  f"numpy float64 ray encode included), {ms:.0f} ms per pass"}

    line = {
        "metric": metric, "value": value, "unit": "sequences/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": DTYPE[precision],
        "data": "synthetic (seeded uv + intrinsics, seeded reference-shaped weights)", "config": config,
        "clocks": clocks, "e2e": e2e,
        "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step,
        "roofline": roofline, "cpu_baseline": cpu_baseline, "parity_relerr_vs_oracle_f64": relerr, "gather_check": gather,
        "flops_per_sequence": flops_per_sequence(spec), "achieved_tflops_step": flops_per_sequence(spec) * value / 1e12,
        "precision": precision, "weight_bytes": weight_bytes, "extra": extra,
    }
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    if world > 1:
        os.dup2(2, 1)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

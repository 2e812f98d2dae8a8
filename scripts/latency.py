"""Small-batch latency of the lifting path (BASELINE.json config 0: one clip, T=27, batch 1), with the launch sequence
replayed from a CUDA graph (default) and launched kernel by kernel (option graph_max_batch=0).  Per case: blocking
latency (call + synchronize, median of 300) and back-to-back rate (300 calls, one synchronize).  Run under gpurun."""
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from ray3d_b200 import Lifter, NetSpec, synth
    for T, widths in ((27, (3, 3, 3)), (243, (3, 3, 3, 3, 3))):
        spec = NetSpec(filter_widths=widths, stage=1)
        sp, st = synth.make_state_dicts(spec)
        for graphs in (1, 0):
            lf = Lifter(spec, sp, st, precision="bf16x3")
            if not graphs:
                lf.plan.set_option("graph_max_batch", 0)
            for B in (1, 16, 64):
                uv, cam = synth.make_inputs(spec, B, seed=3)
                uvc, camc = torch.from_numpy(uv).cuda(), torch.from_numpy(cam).cuda()
                for _ in range(20):
                    lf.forward_uv(uvc, camc)
                torch.cuda.synchronize()
                lat = []
                for _ in range(300):
                    t0 = time.perf_counter()
                    lf.forward_uv(uvc, camc)
                    torch.cuda.synchronize()
                    lat.append((time.perf_counter() - t0) * 1e6)
                t0 = time.perf_counter()
                for _ in range(300):
                    lf.forward_uv(uvc, camc)
                torch.cuda.synchronize()
                rate_us = (time.perf_counter() - t0) * 1e6 / 300
                print(json.dumps(dict(T=T, batch=B, cuda_graph=bool(graphs), latency_us_median=round(statistics.median(lat), 1),
                                      latency_us_p10=round(sorted(lat)[30], 1), back_to_back_us=round(rate_us, 1),
                                      graph_launches=lf.plan.graph_launches)), flush=True)
            del lf


if __name__ == "__main__":
    main()

"""Kernel timeline of ONE training step (B=128, bf16) from torch.profiler (CUPTI): start offset, duration, gap to the
previous kernel and name of every kernel in launch order -> gpurun_out/step_trace.txt.  The durations are the in-step
ones (warm L2, power-capped clocks), unlike the serialised cold-cache launch list ncu gives."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import bench


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/step_trace.txt"
    dev = torch.device("cuda", 0)
    from editor_b200.train import Trainer
    model, sd, x, label, cam, _ = bench.build_case(dev, 128, seed=1)
    model.train()
    tr = Trainer(model)
    xg = {k: v.to(dev) for k, v in x.items()}
    lg, cg = label.to(dev), cam.to(dev)
    for _ in range(4):
        tr.step(xg, lg, cg)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        tr.step(xg, lg, cg)
        tr.step(xg, lg, cg)
        torch.cuda.synchronize()
    tmp = out + ".json"
    prof.export_chrome_trace(tmp)
    ev = json.load(open(tmp))["traceEvents"]
    ks = sorted((e for e in ev if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy")), key=lambda e: e["ts"])
    # second step only: starts at the second gradient-arena memset / first kernel after the largest idle gap is unreliable, so
    # split at the midpoint of the kernel list (both steps launch the same kernels)
    half = len(ks) // 2
    ks = ks[half:]
    t0 = ks[0]["ts"]
    prev_end = t0
    with open(out, "w") as f:
        f.write("# start_us dur_us gap_us name   (one training step, %d kernels)\n" % len(ks))
        for e in ks:
            f.write("%9.1f %8.1f %7.1f %s\n" % (e["ts"] - t0, e["dur"], e["ts"] - prev_end, e["name"][:90]))
            prev_end = max(prev_end, e["ts"] + e["dur"])
        f.write("# total %.1f us, busy %.1f us\n" % (prev_end - t0, sum(e["dur"] for e in ks)))
    # CPU side: time the Python thread spends per step (launch-bound phases show up as GPU gaps)
    cpu = [e for e in ev if e.get("cat") == "cuda_runtime" and e.get("name", "").startswith("cudaLaunchKernel")]
    with open(out, "a") as f:
        f.write("# cudaLaunchKernel* calls in the two profiled steps: %d\n" % len(cpu))
    os.remove(tmp)


if __name__ == "__main__":
    main()

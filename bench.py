#!/usr/bin/env python
"""Benchmark of the EDITOR hot path on B200 (contract: see the task statement; one JSON line on rank 0).

    python bench.py --gpus N --steps K --warmup W            # own arm: CUDA path, bf16 train step, B=128 per GPU
    python bench.py --impl reference --gpus N ...            # reference arm: the CPU oracle on the host cores

A step is one training step (forward + loss + backward + gradient allreduce + SGD) of configs[1] of BASELINE.json:
RGBNT201 EDITOR.yml, ViT-B/16, batch 128 per GPU, bf16, synthetic RGB/NIR/TIR.
"""
import argparse
import gc
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = "images/sec (3-modal, B=128, 256x128)"
WORKLOAD = "RGBNT201 EDITOR.yml ViT-B/16 train step, batch 128 per GPU, bf16, synthetic RGB/NIR/TIR"
# algorithmic FLOPs per 3-modal image of one training step: 3 x forward GEMM FLOPs,
# backbone 68.03 GFLOP + HMA at the measured kept-token count (computed per run below)
BACKBONE_FWD_GFLOP = 68.032770048


def hma_fwd_gflop(n_sel):
    t = 1 + n_sel

    def blk(t):
        return 2 * t * 768 * 2304 + 4 * 12 * t * t * 64 + 2 * t * 768 * 768 + 4 * t * 768 * 3072
    return (3 * blk(t) + blk(3 * t)) / 1e9


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def wait_first_sample(self, timeout=5.0):
        """nvidia-smi takes ~0.1 s to initialise NVML (and holds driver locks while it does): it is started BEFORE the
        warm-up and the timed region only begins once it is polling steadily, so that its start-up never lands inside."""
        t0 = time.time()
        while self.proc is not None and not self.rows and time.time() - t0 < timeout:
            time.sleep(0.05)

    def stop(self, t_begin=None, t_end=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for t, r in self.rows if (t_begin is None or t >= t_begin) and (t_end is None or t <= t_end + 0.2)]
        for r in rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_case(device, batch, seed):
    import contextlib
    import io
    from editor_b200 import synth
    from editor_b200.config import cfg
    from editor_b200.modeling import make_model
    c = cfg.clone()
    c.merge_from_file(os.path.join(ROOT, "configs", "RGBNT201", "EDITOR.yml"))
    c.MODEL.PRETRAIN_CHOICE = "none"          # no checkpoint offline: random-init weights of the named architecture
    with contextlib.redirect_stdout(io.StringIO()):
        model = make_model(c, 171, 4)
    sd = synth.synthetic_state_dict(seed=1111, num_class=171, camera_num=4, al=True)
    model.load_state_dict(sd, strict=True)
    x, label, cam = synth.synthetic_batch(batch, 256, 128, seed=seed, num_cams=4, instances=16 if batch % 16 == 0 else 2)
    return model.to(device), sd, x, label, cam


def cpu_oracle_rate(batch, steps, warmup, sd=None):
    """images/s of the oracle's training step (forward + loss + backward + SGD) on the host cores."""
    from editor_b200 import synth
    from oracle import editor_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    if sd is None:
        sd = synth.synthetic_state_dict(seed=1111, num_class=171, camera_num=4, al=True)
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "centers" not in k and "running" not in k
              and not k.startswith("FREQ_INDEX") else v.clone()) for k, v in sd.items()}
    params = [v for v in sd.values() if v.requires_grad]
    opt = torch.optim.SGD(params, lr=0.001, momentum=0.9, weight_decay=1e-4)
    x, label, cam = synth.synthetic_batch(batch, 256, 128, seed=1, num_cams=4, instances=2)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        state = {}
        outs = orc.editor_forward(sd, x, cam, label=label, training=True, al=True, state_out=state)
        loss = orc.reference_loss(outs, label)
        loss.backward()
        opt.step()
        with torch.no_grad():
            for k, v in state.items():
                sd[k] = v
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return batch * len(times) / sum(times), sum(times) / len(times)


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    _, t_cal = cpu_oracle_rate(2, 1, 0)                     # calibration step (also the warm-up of the thread pool)
    budget = 150.0 / max(args.steps + args.warmup, 1)
    batch = 2
    for b in (4, 8, 16, 32):
        if t_cal * b / 2 <= budget:
            batch = b
    rate, t_step = cpu_oracle_rate(batch, args.steps, args.warmup)
    line = {"metric": METRIC, "value": rate, "unit": "images/sec", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD, "sample": "train step on %d 3-modal images per step, fp32, torch CPU" % batch},
            "cpu_baseline": {"value": rate, "unit": "images/sec", "cores": cores, "kind": "port",
                             "sample": "oracle/editor_oracle.py train step (fwd+loss+bwd+SGD), %d images/step x %d steps, "
                                       "%d threads" % (batch, args.steps, cores)},
            "e2e": {"value": rate, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own")
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--preheat", type=int, default=12, help="untimed conditioning steps before the warm-up")
    ap.add_argument("--gemm-mode", type=int, default=0, help="0 = CTA-pair tcgen05 tiles (default), 1 = single-CTA tiles")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL writes its banner ("NCCL version ...", and everything NCCL_DEBUG asks for) to stdout when the communicator
        # is created: send that to stderr so that stdout carries the one JSON line only
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=device)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    from editor_b200 import lib
    from editor_b200.train import Trainer
    lib.gemm_set_mode(args.gemm_mode)
    B = args.batch
    model, sd, x, label, cam = build_case(device, B, seed=1 + rank)
    model.train()
    trainer = Trainer(model)
    xg = {k: v.to(device) for k, v in x.items()}
    lg, cg = label.to(device), cam.to(device)
    xh = {k: v.pin_memory() for k, v in x.items()}
    lh, ch = label.pin_memory(), cam.pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value")
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # untimed conditioning before the W warm-up steps: the first second of load after an idle GPU runs into the power
    # limiter harder than the steady state does (first-region steps measured 15-30 % slower than later ones on some boxes)
    for _ in range(args.preheat):
        trainer.step(xg, lg, cg)
    for _ in range(args.warmup):
        trainer.step(xg, lg, cg)
    # long-lived objects (model, arena views, workspace) leave the cyclic collector's young generations: a full collection
    # landing inside the timed region cost one step ~18 ms (step_ms_each showed 57.9 ms once in eight)
    gc.collect()
    gc.freeze()
    if rank == 0:
        sampler.wait_first_sample()
    barrier()
    n0 = lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.time()
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    e0.record()
    for i in range(args.steps):
        loss, _ = trainer.step(xg, lg, cg)
        marks[i].record()
    e1.record()
    barrier()
    t_end = time.time()
    ms = e0.elapsed_time(e1)
    step_each = [round(a.elapsed_time(b), 2) for a, b in zip([e0] + marks[:-1], marks)]
    launches = lib.launch_count - n0
    n_sel = float(model.engine().last["num"].float().mean().item())
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    # ---- end to end ("e2e"): pinned host inputs -> H2D every step, loss read back every step
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(args.steps):
        xs = {k: v.to(device, non_blocking=True) for k, v in xh.items()}
        ls, cs = lh.to(device, non_blocking=True), ch.to(device, non_blocking=True)
        loss, _ = trainer.step(xs, ls, cs)
        loss_host = loss.item()
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    h2d = sum(v.numel() * v.element_size() for v in xh.values()) + lh.numel() * 8 + ch.numel() * 8
    # ---- dominant kernel (tcgen05 GEMM): per-launch CUDA-event timing over one extra step, outside the timed region
    roof = None
    # (every rank runs the two extra, untimed steps below: the step contains the gradient allreduce)
    lib.gemm_timing = []
    trainer.step(xg, lg, cg)
    torch.cuda.synchronize()
    tm = lib.gemm_timing
    lib.gemm_timing = None
    if rank == 0:
        fl = sum(f for f, _, _, _ in tm)
        tt = sum(a.elapsed_time(b) for _, a, b, _ in tm) * 1e-3
        shapes = {}
        for f, a, b, key in tm:
            r = shapes.setdefault(key, [0, 0.0, 0.0])
            r[0] += 1
            r[1] += a.elapsed_time(b)
            r[2] += f
        gemm_table = [{"shape": k, "launches": v[0], "ms": round(v[1], 3), "tflops": round(v[2] / v[1] / 1e9, 1)}
                      for k, v in sorted(shapes.items(), key=lambda kv: -kv[1][1])]
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = peaks.get("bf16_tflops_sustained", 1400.0)
        ach = fl / tt / 1e12 if tt > 0 else 0.0
        traffic = None
        try:        # dram bytes per launch from the committed ncu --set full capture of this kernel (profiles/)
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_gemm_traffic.json")))["mean_dram_bytes_per_launch"]
        except (OSError, KeyError):
            pass
        roof = {"bound": "tensor", "kernel": "gemm_bf16_kernel (tcgen05)", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": ach / peak, "traffic": traffic, "algorithmic_bytes_per_launch_mean": None,
                "launches_per_step": len(tm), "gemm_ms_per_step": tt * 1e3,
                "tile": "CTA pair 256x256 (cta_group::2)" if args.gemm_mode == 0 else "single CTA 128x256",
                "by_shape": gemm_table[:16],
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1400 (B200_PROFILING.md)"}
    eng = model.engine()
    eng.stats["events"] = []
    t0 = torch.cuda.Event(enable_timing=True)
    t0.record()
    trainer.step(xg, lg, cg)
    t1 = torch.cuda.Event(enable_timing=True)
    t1.record()
    torch.cuda.synchronize()
    ev = [("step_start", t0)] + eng.stats["events"] + [("step_end", t1)]
    eng.stats["events"] = None
    breakdown = {"%s->%s" % (a[0], b[0]): round(a[1].elapsed_time(b[1]), 3) for a, b in zip(ev[:-1], ev[1:])}
    barrier()
    # ---- secondary number: eval forward (fp32-faithful mode is what engine/processor.py:176-186 runs; bf16 also shown)
    eval_rates = {}
    model.eval()
    for mode in ("bf16", "fp32"):
        model.precision = mode
        nrep = 3 if mode == "bf16" else 1
        model(xg, cam_label=cg)
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(nrep):
            model(xg, cam_label=cg)
        a1.record()
        barrier()
        eval_rates[mode] = world * B * nrep / (a0.elapsed_time(a1) * 1e-3)
    model.precision = "auto"
    model.train()
    # ---- secondary number: retrieval evaluation of one epoch -- RGBNT100-sized query / gallery sets of
    # 2304-wide features: normalise + distance matrix + CMC / mAP on the GPU; the numpy oracle on a bounded query sample
    eval_metrics = None
    if rank == 0:
        from editor_b200 import metrics as M
        gq = torch.Generator(device="cpu").manual_seed(11)
        nq, ng, nid = 1715, 8575, 50
        feats = torch.randn(nq + ng, 2304, generator=gq)
        pids = torch.randint(0, nid, (nq + ng,), generator=gq).numpy()
        cams = torch.randint(0, 8, (nq + ng,), generator=gq).numpy()
        fd = feats.to(device)
        pq, pg = torch.from_numpy(pids[:nq]).to(device), torch.from_numpy(pids[nq:]).to(device)
        cq, cg = torch.from_numpy(cams[:nq]).to(device), torch.from_numpy(cams[nq:]).to(device)

        def run():
            f = M.normalize_(fd.clone())
            return M._rank(M.distmat_device(f[:nq], f[nq:]), pq, pg, cq, cg, 50)
        run()
        torch.cuda.synchronize()
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m0.record()
        cmc_g, map_g = run()
        m1.record()
        torch.cuda.synchronize()
        eval_metrics = {"queries": nq, "gallery": ng, "feature_dim": 2304, "gpu_ms": round(m0.elapsed_time(m1), 3),
                        "mAP": map_g, "rank1": float(cmc_g[0])}
        if world == 1 and not args.no_cpu_baseline:
            from oracle import eval_oracle as eo
            ns = 128
            t0 = time.perf_counter()
            fn = eo.l2_normalize(feats.numpy())
            d = eo.euclidean_distance(fn[:ns], fn[nq:])
            eo.eval_func(d, pids[:ns], pids[nq:], cams[:ns], cams[nq:])
            eval_metrics["cpu_oracle_ms_per_query"] = round((time.perf_counter() - t0) * 1e3 / ns, 3)
            eval_metrics["cpu_sample"] = "oracle/eval_oracle.py on %d of the %d queries" % (ns, nq)
        del fd
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ms_step = ms / args.steps
    value = world * B * args.steps / (ms * 1e-3)
    e2e_v = world * B * args.steps / (ms_e2e * 1e-3)
    step_gflop_img = 3 * (BACKBONE_FWD_GFLOP + hma_fwd_gflop(n_sel))
    peak_s = roof["peak"] if roof else 1400.0
    line = {"metric": METRIC, "value": value, "unit": "images/sec", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": world * B, "parallelism": "dp%d" % world,
                       "l2_policy": "inputs+activations per step (>18 GB) far exceed the 126 MB L2",
                       "kept_tokens_mean": n_sel, "drop_path": 0.1, "preheat_steps": args.preheat,
                       "step": "forward + CE/triplet loss + backward + grad allreduce + fused SGD"},
            "clocks": clocks,
            "e2e": {"value": e2e_v, "unit": "images/sec", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "step_ms_each": step_each,
            "step_tflops_of_peak": {"algorithmic_gflop_per_image": step_gflop_img,
                                    "achieved_tflops_per_gpu": step_gflop_img * B / ms_step,
                                    "frac_of_peak": step_gflop_img * B / ms_step / peak_s},
            "roofline": roof, "phase_ms": breakdown, "eval_forward_images_per_sec": eval_rates, "eval_metrics": eval_metrics,
            "loss": float(loss_host)}
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        rate, t_step = cpu_oracle_rate(4, 2, 1, sd)
        line["cpu_baseline"] = {"value": rate, "unit": "images/sec", "cores": cores, "kind": "port",
                                "sample": "oracle/editor_oracle.py train step (fwd+loss+bwd+SGD) fp32, 4 images/step x 2 "
                                          "steps after 1 warm-up, %d threads" % cores}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

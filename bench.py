#!/usr/bin/env python
"""Benchmark of the EDITOR hot path on B200 (contract: see the task statement; one JSON line on rank 0).

    python bench.py --gpus N --steps K --warmup W            # own arm: CUDA path, bf16 train step, B=128 per GPU
    python bench.py --config RGBNT100 ...                    # BASELINE.json configs[2] (128x256, AL=0, 9 outputs)
    python bench.py --config MSVR310 --precision fp32 ...    # BASELINE.json configs[4] (fp32-faithful training mode)
    python bench.py --impl reference --gpus N ...            # reference arm: the CPU oracle on the host cores

A step is one training step (forward + loss + backward + gradient allreduce + SGD) of the named yml at batch 128 per GPU
on synthetic RGB/NIR/TIR; the default is BASELINE.json configs[1]: RGBNT201 EDITOR.yml, ViT-B/16, bf16.

Besides the contract's keys the own arm at N=1 also reports, from the same box and outside the timed region:
  torch_eager_gpu   the UNMODIFIED reference (baseline/_ref) trained by its own engine/processor.py::do_train on this GPU,
                    under fp16 autocast + GradScaler (what the reference does) and under bf16 autocast;
  do_train_dropin   THIS repo's model driven by that same unmodified do_train;
  sfts_isolation    BASELINE.json configs[3]: the SFTS kernels alone at B=512, achieved HBM GB/s per kernel.
"""
import argparse
import gc
import importlib.util
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
# yml -> (num_class, cameras, H, W): SURVEY.md section 8 preamble (MSVR310's class count is data-dependent; 155 synthetic)
CASES = {"RGBNT201": (171, 4, 256, 128), "RGBNT100": (50, 8, 128, 256), "MSVR310": (155, 8, 128, 256)}
# algorithmic FLOPs per 3-modal image of one training step: 3 x forward GEMM FLOPs,
# backbone 68.03 GFLOP + HMA at the measured kept-token count (computed per run below)
BACKBONE_FWD_GFLOP = 68.032770048


def workload_name(config, precision, batch):
    return "%s EDITOR.yml ViT-B/16 train step, batch %d per GPU, %s, synthetic RGB/NIR/TIR" % (config, batch, precision)


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


def build_case(device, batch, seed, config="RGBNT201", drop_path=None):
    import contextlib
    import io
    from editor_b200 import synth
    from editor_b200.config import cfg
    from editor_b200.modeling import make_model
    C, cams, H, W = CASES[config]
    c = cfg.clone()
    c.merge_from_file(os.path.join(ROOT, "configs", config, "EDITOR.yml"))
    c.MODEL.PRETRAIN_CHOICE = "none"          # no checkpoint offline: random-init weights of the named architecture
    if drop_path is not None:
        c.MODEL.DROP_PATH = drop_path
    al = bool(c.MODEL.AL)
    with contextlib.redirect_stdout(io.StringIO()):
        model = make_model(c, C, cams)
    sd = synth.synthetic_state_dict(seed=1111, num_class=C, camera_num=cams, al=al)
    model.load_state_dict(sd, strict=True)
    x, label, cam = synth.synthetic_batch(batch, H, W, seed=seed, num_cams=cams, instances=16 if batch % 16 == 0 else 2)
    return model.to(device), sd, x, label, cam, float(c.MODEL.DROP_PATH)


def cpu_oracle_rate(batch, steps, warmup, sd=None, config="RGBNT201"):
    """images/s of the oracle's training step (forward + loss + backward + SGD) on the host cores."""
    from editor_b200 import synth
    from oracle import editor_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    C, cams, H, W = CASES[config]
    al = config == "RGBNT201"
    if sd is None:
        sd = synth.synthetic_state_dict(seed=1111, num_class=C, camera_num=cams, al=al)
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "centers" not in k and "running" not in k
              and not k.startswith("FREQ_INDEX") else v.clone()) for k, v in sd.items()}
    params = [v for v in sd.values() if v.requires_grad]
    opt = torch.optim.SGD(params, lr=0.001, momentum=0.9, weight_decay=1e-4)
    x, label, cam = synth.synthetic_batch(batch, H, W, seed=1, num_cams=cams, instances=2)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        state = {}
        outs = orc.editor_forward(sd, x, cam, label=label, training=True, al=al, state_out=state)
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
    """Reference arm (tier contract): the reference's CPU path -- its restatement oracle/editor_oracle.py, pinned against
    the unmodified reference by tests/golden -- on all host cores, on a bounded sample of the workload."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    _, t_cal = cpu_oracle_rate(2, 1, 0, config=args.config)     # calibration step (also warms up the thread pool)
    budget = 150.0 / max(args.steps + args.warmup, 1)
    batch = 2
    for b in (4, 8, 16, 32):
        if t_cal * b / 2 <= budget:
            batch = b
    rate, t_step = cpu_oracle_rate(batch, args.steps, args.warmup, config=args.config)
    line = {"metric": METRIC, "value": rate, "unit": "images/sec", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_name(args.config, args.precision, args.batch),
                       "sample": "train step on %d 3-modal images per step (NOT 128), fp32, torch CPU, %d threads"
                                 % (batch, cores), "same_config": False},
            "cpu_baseline": {"value": rate, "unit": "images/sec", "cores": cores, "kind": "port",
                             "sample": "oracle/editor_oracle.py train step (fwd+loss+bwd+SGD), %d images/step x %d steps, "
                                       "%d threads" % (batch, args.steps, cores)},
            "e2e": {"value": rate, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


class PrefetchedBatches:
    """End-to-end input path: every step's batch is copied from pinned host memory to one of two device buffers on a copy
    stream while the previous step computes (what a DataLoader(pin_memory=True) + prefetcher does); the compute stream
    waits on the copy's event, the copy stream waits until the step that last read the buffer has finished."""

    def __init__(self, host_batches, device):
        self.host = host_batches
        self.stream = torch.cuda.Stream(device=device)
        x, l, c = host_batches[0]
        self.dev = [({k: torch.empty_like(v, device=device) for k, v in x.items()}, torch.empty_like(l, device=device),
                     torch.empty_like(c, device=device)) for _ in range(2)]
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.free = [torch.cuda.Event() for _ in range(2)]
        self.bytes = sum(v.numel() * v.element_size() for v in x.values()) + l.numel() * 8 + c.numel() * 8
        for e in self.free:
            e.record()

    def issue(self, i):
        slot = i % 2
        hx, hl, hc = self.host[i % len(self.host)]
        dx, dl, dc = self.dev[slot]
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(self.free[slot])
            for k in hx:
                dx[k].copy_(hx[k], non_blocking=True)
            dl.copy_(hl, non_blocking=True)
            dc.copy_(hc, non_blocking=True)
            self.ready[slot].record(self.stream)

    def get(self, i):
        slot = i % 2
        torch.cuda.current_stream().wait_event(self.ready[slot])
        return self.dev[slot]

    def release(self, i):
        self.free[i % 2].record()


def gemm_roofline(tm, counts, peaks, gemm_mode):
    """FLOPs / bytes from the ACTUAL row counts: HMA GEMMs are launched with a static bound and trim at a device-side
    count (M_dev / K_dev), so the bound would over-count."""
    fl_total = t_total = by_total = 0.0
    shapes = {}
    for r in tm:
        M, N, K = r["M"], r["N"], r["K"]
        if r["M_dev"] is not None:
            M = min(M, counts.get(r["M_dev"], M))
        if r["K_dev"] is not None:
            K = min(K, counts.get(r["K_dev"], K))
        ms = r["e0"].elapsed_time(r["e1"])
        fl = 2.0 * M * N * K
        by = 2.0 * (M * K + N * K) + float(r["out_bytes_per_elem"]) * M * N
        fl_total += fl
        t_total += ms
        by_total += by
        s = shapes.setdefault(r["key"], [0, 0.0, 0.0, M, N, K])
        s[0] += 1
        s[1] += ms
        s[2] += fl
    table = [{"shape": k, "rows_actual": [v[3], v[4], v[5]], "launches": v[0], "ms": round(v[1], 3),
              "tflops": round(v[2] / v[1] / 1e9, 1)} for k, v in sorted(shapes.items(), key=lambda kv: -kv[1][1])]
    peak = peaks.get("bf16_tflops_sustained", 1400.0)
    ach = fl_total / (t_total * 1e-3) / 1e12 if t_total > 0 else 0.0
    traffic, traffic_src = None, None
    for name in ("r02_gemm_traffic.json", "r01_gemm_traffic.json"):
        try:        # dram bytes per launch from the committed ncu --set full capture of this kernel (profiles/)
            traffic = json.load(open(os.path.join(ROOT, "profiles", name)))["mean_dram_bytes_per_launch"]
            traffic_src = "profiles/" + name + " (ncu --set full capture, not this run)"
            break
        except (OSError, KeyError):
            pass
    return {"bound": "tensor", "kernel": "gemm_bf16_kernel (tcgen05)", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
            "frac": ach / peak, "traffic": traffic, "traffic_source": traffic_src,
            "algorithmic_bytes_per_launch_mean": by_total / max(len(tm), 1),
            "algorithmic_flop_per_launch_mean": fl_total / max(len(tm), 1),
            "flop_counting": "2*M*N*K with the device-side row counts read back after the step",
            "launches_per_step": len(tm), "gemm_ms_per_step": t_total,
            "tile": "CTA pair 256x256 (cta_group::2)" if gemm_mode == 0 else "single CTA 128x256",
            "by_shape": table[:16],
            "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1400 (B200_PROFILING.md)"}


def reference_gpu_legs(args):
    """The measured denominator: baseline/run_ref.py in fresh processes on this GPU (see its docstring)."""
    script = os.path.join(ROOT, "baseline", "run_ref.py")
    if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "engine")):
        why = {"unavailable": "baseline/_ref absent (the git-ignored copy of the reference made by baseline/install_ref.py)"}
        return why, why
    common = ["--config", args.config, "--batch", str(args.batch), "--steps", str(args.ref_steps), "--warmup", "3"]

    def leg(*extra):
        try:
            out = subprocess.run([sys.executable, script, *common, *extra], capture_output=True, text=True, timeout=600)
            if out.returncode != 0:
                return {"failed": out.stderr.strip().splitlines()[-1][:300] if out.stderr.strip() else "rc %d" % out.returncode}
            r = json.loads(out.stdout.strip().splitlines()[-1])
            return {k: r[k] for k in ("ms_per_step", "images_per_sec", "peak_mem_gb", "params_finite", "inputs", "losses",
                                      "ms_each", "loop", "amp", "drop_path") if k in r}
        except (subprocess.TimeoutExpired, ValueError, IndexError) as e:
            return {"failed": repr(e)[:300]}
    eager = {"what": "UNMODIFIED reference model + make_loss + make_optimizer trained by engine/processor.py::do_train on "
                     "this GPU, B=%d, synthetic P x K batches from pinned host memory" % args.batch,
             "fp16_scaler": leg("--model", "reference", "--amp", "fp16"),
             "bf16": leg("--model", "reference", "--amp", "bf16")}
    dropin = {"what": "this repo's make_model driven by the same unmodified do_train (GradScaler + per-tensor torch SGD)",
              "fp16_scaler": leg("--model", "ours", "--amp", "fp16")}
    return eager, dropin


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own")
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--config", default="RGBNT201", choices=sorted(CASES))
    ap.add_argument("--precision", default="bf16", choices=("bf16", "fp32"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip the torch_eager_gpu / do_train_dropin legs")
    ap.add_argument("--no-sfts", action="store_true", help="skip the SFTS isolation sweep")
    ap.add_argument("--ref-steps", type=int, default=8)
    ap.add_argument("--preheat", type=int, default=12, help="untimed conditioning steps before the warm-up")
    ap.add_argument("--gemm-mode", type=int, default=0, help="0 = CTA-pair tcgen05 tiles (default), 1 = single-CTA tiles")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel of the step eagerly (no CUDA graph)")
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
    fp32 = args.precision == "fp32"
    model, sd, x, label, cam, drop_path = build_case(device, B, seed=1 + rank, config=args.config,
                                                     drop_path=0.0 if fp32 else None)
    if fp32:
        model.precision = "fp32"            # fp32-faithful training (3-piece bf16 split GEMMs, fp32 attention): configs[4]
        args.preheat = min(args.preheat, 2)
    model.train()
    trainer = Trainer(model)
    xg = {k: v.to(device) for k, v in x.items()}
    lg, cg = label.to(device), cam.to(device)
    # two distinct host batches, pinned: the e2e loop alternates between them
    from editor_b200 import synth
    C, cams, H, W = CASES[args.config]
    x2, l2, c2 = synth.synthetic_batch(B, H, W, seed=101 + rank, num_cams=cams, instances=16 if B % 16 == 0 else 2)
    host = [({k: v.pin_memory() for k, v in xx.items()}, ll.pin_memory(), cc.pin_memory())
            for xx, ll, cc in ((x, label, cam), (x2, l2, c2))]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value")
    sampler = ClockSampler(local)
    # the whole step (forward + loss + backward + allreduce + SGD) as one CUDA graph: Trainer.capture; eager otherwise
    graphed = (not args.no_graph) and (not fp32) and trainer.capture(xg, lg, cg)
    if world > 1:       # every rank replays, or none does
        flag = torch.tensor([1 if graphed else 0], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        graphed = bool(flag.item()) and graphed
    if not graphed and not args.no_graph and not fp32 and world == 1:
        # a failed capture leaves torch's CUDA generator in capture mode (the next torch.rand raises): start over eagerly
        sys.stderr.write("bench.py: CUDA-graph capture failed (%s); restarting with --no-graph\n"
                         % getattr(trainer, "capture_error", None))
        sys.stderr.flush()
        os.execv(sys.executable, [sys.executable] + sys.argv + ["--no-graph"])
    if rank == 0:
        sampler.start()                      # (after the capture: a restart above must not orphan the nvidia-smi child)
    if graphed:
        xg, lg, cg = trainer.static_in       # device-resident inputs = the graph's static buffers (no per-step D2D copy)
    step_fn = trainer.step_graphed if graphed else trainer.step
    # untimed conditioning before the W warm-up steps: the first second of load after an idle GPU runs into the power
    # limiter harder than the steady state does (first-region steps measured 15-30 % slower than later ones on some boxes)
    for _ in range(args.preheat):
        step_fn(xg, lg, cg)
    for _ in range(args.warmup):
        step_fn(xg, lg, cg)
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
        loss, _ = step_fn(xg, lg, cg)
        marks[i].record()
    e1.record()
    barrier()
    t_end = time.time()
    ms = e0.elapsed_time(e1)
    step_each = [round(a.elapsed_time(b), 2) for a, b in zip([e0] + marks[:-1], marks)]
    launches = lib.launch_count - n0
    n_sel = float(model.engine().last["num"].float().mean().item())
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    # ---- end to end ("e2e"): every step's inputs come from pinned host memory (H2D inside the timed region, prefetched one
    # step ahead on a copy stream) and every step's loss is read back on the host (one step late, so that the read-back
    # does not drain the GPU): K copies and K read-backs in the region, the first copy and the last read-back exposed
    pf = PrefetchedBatches(host, device)
    loss_pinned = torch.zeros(args.steps, dtype=torch.float32).pin_memory()
    loss_ev = [torch.cuda.Event() for _ in range(args.steps)]
    losses_host = []
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    pf.issue(0)
    for i in range(args.steps):
        xs, ls, cs = pf.get(i)
        if i + 1 < args.steps:
            pf.issue(i + 1)
        loss, _ = step_fn(xs, ls, cs)       # graphed: one D2D copy into the graph's static inputs, then the replay
        pf.release(i)
        loss_pinned[i:i + 1].copy_(loss.detach().reshape(1), non_blocking=True)
        loss_ev[i].record()
        if i > 0:
            loss_ev[i - 1].synchronize()
            losses_host.append(float(loss_pinned[i - 1]))
    loss_ev[-1].synchronize()
    losses_host.append(float(loss_pinned[args.steps - 1]))
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    loss_host = losses_host[-1]
    # ---- the same end-to-end loop fed as the real pipeline would be (row f-4): decoded uint8 HWC images in pinned host
    # memory (3 bytes per pixel instead of 12), H2D one step ahead, the reference's training transform (resize / flip / pad +
    # crop / normalise / pixel-mode RandomErasing, data/datasets/make_dataloader.py:245-253) as one CUDA kernel per batch
    e2e_u8 = None
    try:
        from editor_b200 import data as edata
        from editor_b200.config import cfg as _cfg
        ac = _cfg.clone()
        ac.merge_from_file(os.path.join(ROOT, "configs", args.config, "EDITOR.yml"))
        aug = edata.GpuAugment(ac, device, seed=1 + rank)
        gq = torch.Generator().manual_seed(7 + rank)
        host_u8 = [({k: torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, generator=gq).pin_memory() for k in ("RGB", "NI", "TI")},
                    ll, cc) for _, ll, cc in host]
        pf8 = PrefetchedBatches(host_u8, device)
        # the augmentation writes straight into the step's input buffers (the graph's static inputs when graphed)
        xbuf = trainer.static_in[0] if graphed else {k: torch.empty(B, 3, H, W, dtype=torch.float32, device=device)
                                                      for k in ("RGB", "NI", "TI")}
        for i in range(2):                                   # untimed: tables, pinned parameter buffers
            pf8.issue(i)
            xs, ls, cs = pf8.get(i)
            step_fn(aug(xs, out=xbuf), ls, cs)
            pf8.release(i)
        barrier()
        e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e4.record()
        pf8.issue(0)
        for i in range(args.steps):
            xs, ls, cs = pf8.get(i)
            if i + 1 < args.steps:
                pf8.issue(i + 1)
            xf = aug(xs, out=xbuf)
            pf8.release(i)
            loss8, _ = step_fn(xf, ls, cs)
            loss_pinned[i:i + 1].copy_(loss8.detach().reshape(1), non_blocking=True)
            loss_ev[i].record()
            if i > 0:
                loss_ev[i - 1].synchronize()
        loss_ev[-1].synchronize()
        e5.record()
        barrier()
        t8 = torch.tensor([e4.elapsed_time(e5)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t8, op=dist.ReduceOp.MAX)
        e2e_u8 = {"value": world * B * args.steps / (float(t8[0]) * 1e-3), "unit": "images/sec",
                  "ms_per_step": float(t8[0]) / args.steps, "h2d_bytes_per_step": pf8.bytes, "d2h_bytes_per_step": 4,
                  "how": "uint8 HWC images from pinned host memory, H2D one step ahead, edb_augment_u8 (resize/flip/pad-crop/"
                         "normalise/RandomErasing) on the device, then the training step; loss read back every step"}
    except Exception as e:      # noqa: BLE001 - secondary number
        e2e_u8 = {"failed": repr(e)[:300]}
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    h2d = pf.bytes
    # ---- dominant kernel (tcgen05 GEMM): per-launch CUDA-event timing over one extra step, outside the timed region
    roof = None
    # (every rank runs the two extra, untimed steps below: the step contains the gradient allreduce)
    lib.gemm_timing = []
    trainer.step(xg, lg, cg)
    torch.cuda.synchronize()
    tm = lib.gemm_timing
    lib.gemm_timing = None
    eng = model.engine()
    if rank == 0:
        sel = eng.sel
        counts = {}
        if sel.get("T_dev") is not None:
            counts[sel["T_dev"]] = int(sel["seq_off"][-1].item())
            counts[sel["T3_dev"]] = int(sel["seq_off3"][-1].item())
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        roof = gemm_roofline(tm, counts, peaks, args.gemm_mode)
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
    model.precision = "fp32" if fp32 else "auto"
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
        cams_np = torch.randint(0, 8, (nq + ng,), generator=gq).numpy()
        fd = feats.to(device)
        pq, pg = torch.from_numpy(pids[:nq]).to(device), torch.from_numpy(pids[nq:]).to(device)
        cq, cgal = torch.from_numpy(cams_np[:nq]).to(device), torch.from_numpy(cams_np[nq:]).to(device)

        def run():
            f = M.normalize_(fd.clone())
            return M._rank(M.distmat_device(f[:nq], f[nq:]), pq, pg, cq, cgal, 50)
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
            tq = time.perf_counter()
            fn = eo.l2_normalize(feats.numpy())
            d = eo.euclidean_distance(fn[:ns], fn[nq:])
            eo.eval_func(d, pids[:ns], pids[nq:], cams_np[:ns], cams_np[nq:])
            eval_metrics["cpu_oracle_ms_per_query"] = round((time.perf_counter() - tq) * 1e3 / ns, 3)
            eval_metrics["cpu_sample"] = "oracle/eval_oracle.py on %d of the %d queries" % (ns, nq)
        del fd
    if rank != 0:
        finish(world)
        return
    ms_step = ms / args.steps
    value = world * B * args.steps / (ms * 1e-3)
    e2e_v = world * B * args.steps / (ms_e2e * 1e-3)
    step_gflop_img = 3 * (BACKBONE_FWD_GFLOP + hma_fwd_gflop(n_sel))
    peak_s = roof["peak"] if roof else 1400.0
    line = {"metric": METRIC, "value": value, "unit": "images/sec", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if not fp32 else "f32 (3-piece bf16 split on tcgen05)", "data": "synthetic",
            "config": {"workload": workload_name(args.config, args.precision, B), "global_batch": world * B,
                       "parallelism": "dp%d" % world,
                       "l2_policy": "inputs+activations per step (>18 GB) far exceed the 126 MB L2",
                       "kept_tokens_mean": n_sel, "drop_path": drop_path, "preheat_steps": args.preheat,
                       "step": "forward + CE/triplet loss + backward + grad allreduce + fused SGD",
                       "cuda_graph": bool(graphed), "cuda_graph_error": getattr(trainer, "capture_error", None),
                       "untimed_steps_before_warmup": {"graph_capture_warmup": 3 if graphed else 0,
                                                       "preheat": args.preheat}},
            "clocks": clocks,
            "e2e": {"value": e2e_v, "unit": "images/sec", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps,
                    "how": "pinned host batches, H2D on a copy stream one step ahead (double-buffered), loss read back on the "
                           "host every step, one step late; all K copies and K read-backs inside the timed region"},
            "e2e_uint8_pipeline": e2e_u8,
            "gpu_launches": launches, "step_ms_each": step_each,
            "step_tflops_of_peak": {"algorithmic_gflop_per_image": step_gflop_img,
                                    "achieved_tflops_per_gpu": step_gflop_img * B / ms_step,
                                    "frac_of_peak": step_gflop_img * B / ms_step / peak_s},
            "roofline": roof, "phase_ms": breakdown, "eval_forward_images_per_sec": eval_rates, "eval_metrics": eval_metrics,
            "loss": float(loss_host)}
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        rate, t_step = cpu_oracle_rate(4, 2, 1, sd, config=args.config)
        line["cpu_baseline"] = {"value": rate, "unit": "images/sec", "cores": cores, "kind": "port",
                                "sample": "oracle/editor_oracle.py train step (fwd+loss+bwd+SGD) fp32, 4 images/step (NOT 128) "
                                          "x 2 steps after 1 warm-up, %d threads" % cores}
    if world == 1 and not args.no_sfts and not fp32:
        # BASELINE.json configs[3]: SFTS kernels in isolation, B = 512 (free the training workspace first)
        del trainer, eng
        model._engine = None
        gc.unfreeze()
        gc.collect()
        torch.cuda.empty_cache()
        spec = importlib.util.spec_from_file_location("sfts_bench", os.path.join(ROOT, "tools", "sfts_bench.py"))
        sb = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(sb)
        try:
            line["sfts_isolation"] = sb.run(512, "bf16", sweep=((10, 2), (16, 1), (32, 2), (64, 4), (96, 8)))
        except Exception as e:      # noqa: BLE001 - the main line must still be printed
            line["sfts_isolation"] = {"failed": repr(e)[:300]}
        torch.cuda.empty_cache()
    if world == 1 and not args.no_ref_gpu and not fp32:
        torch.cuda.empty_cache()
        eager, dropin = reference_gpu_legs(args)
        line["torch_eager_gpu"], line["do_train_dropin"] = eager, dropin
        try:
            ref_bf16 = eager["bf16"]["images_per_sec"]
            line["vs_torch_eager_gpu"] = {
                "e2e_over_reference_bf16": e2e_v / ref_bf16,
                "e2e_over_reference_fp16_scaler": e2e_v / eager["fp16_scaler"]["images_per_sec"],
                "value_over_reference_bf16": value / ref_bf16,
                "dropin_do_train_over_reference_fp16_scaler": dropin["fp16_scaler"]["images_per_sec"]
                / eager["fp16_scaler"]["images_per_sec"],
                "note": "same GPU, same B and yml; the reference legs read their batches from pinned host memory inside the "
                        "iteration (measured in round 2: 1.5 % of its 185 ms; device-resident inputs 697 vs 688 img/s)"}
        except (KeyError, TypeError, ZeroDivisionError):
            pass
    print(json.dumps(line), flush=True)
    finish(world)


def finish(world):
    """Leave without tearing NCCL down.  At 8 ranks `dist.destroy_process_group()` never returned once the step had been
    captured into a CUDA graph (the communicator is still referenced by the graph's captured collectives; seen on an
    8 x B200 box: the JSON line was out after 60 s, the processes sat in the teardown until the launcher's timeout).  Every
    rank waits for the others (so no rank exits under a peer's collective), flushes, and exits the process directly:
    the driver and the OS reclaim the communicator, the graph and the device memory."""
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        # belt and braces: the line is already out; if even the closing barrier should stall, leave after a minute
        t = threading.Timer(60.0, lambda: os._exit(0))
        t.daemon = True
        t.start()
        try:
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
        except Exception:      # noqa: BLE001 - nothing left to do but exit
            pass
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()

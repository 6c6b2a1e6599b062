"""Bisects what breaks CUDA-graph capture of the training step: captures growing prefixes of the step in fresh processes.
Usage (GPU box): python tools/graph_probe.py            # runs every stage in a subprocess and prints one line each"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
STAGES = ["fwd_loss_bwd", "full_step"]


def stage(name):
    import torch
    import __graft_entry__ as ge
    from editor_b200.train import Trainer, editor_loss
    from editor_b200 import lib
    model, sd, x, label, cam, _ = ge._small_case(True, 4)
    model = model.cuda()
    x = {k: v.cuda() for k, v in x.items()}
    label, cam = label.cuda(), cam.cuda()
    tr = Trainer(model)

    def body():
        if name == "eval_fwd":
            model.eval()
            model.precision = "bf16"
            return model(x, cam_label=cam)
        model.train()
        if name.startswith("full_step"):
            return tr.step(x, label, cam)[0]
        with torch.autocast("cuda", dtype=torch.bfloat16):
            outs = model(x, label=label, cam_label=cam, writer=None, epoch=1)
            if name == "train_fwd":
                return outs[0]
            loss = editor_loss(outs, label)
        if name == "fwd_loss_bwd":
            loss.backward()
        return loss
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            body()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    mode = {"full_step_threadlocal": "thread_local", "full_step_relaxed": "relaxed"}.get(name, "global")
    g = torch.cuda.CUDAGraph()
    n0 = lib.launch_count
    lib.capture_window = True
    print("capture stream will be a side stream; default stream is %#x" % torch.cuda.default_stream().cuda_stream, flush=True)
    try:
        with torch.cuda.graph(g, capture_error_mode=mode):
            print("capturing on %#x" % torch.cuda.current_stream().cuda_stream, flush=True)
            cap = torch.cuda.current_stream().cuda_stream
            # a torch-level watcher between python statements: first line after which the capture is no longer active
            import sys as _sys
            state = {"bad": False}

            def tracer(frame, event, arg):
                if event == "line" and not state["bad"] and ("editor_b200" in frame.f_code.co_filename or "graph_probe" in frame.f_code.co_filename):
                    st = lib.capture_status(cap)
                    if st != 1:
                        state["bad"] = True
                        print("EDB_CAPTURE_DEBUG: status %d first seen BEFORE executing %s:%d (%s)" % (
                            st, frame.f_code.co_filename, frame.f_lineno, frame.f_code.co_name), flush=True)
                return tracer
            import threading as _th
            _th.settrace(tracer)
            _sys.settrace(tracer)
            try:
                out = body()
            finally:
                _sys.settrace(None)
                _th.settrace(None)
            print("EDB_CAPTURE_DEBUG: status at the end of the body: %d" % lib.capture_status(cap), flush=True)
    finally:
        lib.capture_window = False
    n1 = lib.launch_count
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    print("STAGE %s OK: %d C-ABI launches captured, result %s" % (name, n1 - n0, float(out.float().flatten()[0])))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        stage(sys.argv[1])
    else:
        for s in STAGES:
            r = subprocess.run([sys.executable, __file__, s], capture_output=True, text=True, timeout=600,
                               env=dict(os.environ, EDB_CAPTURE_DEBUG="1"))
            dbg = [l for l in r.stdout.splitlines() if "EDB_CAPTURE_DEBUG" in l or "captur" in l or "File" in l or l.startswith("    ")]
            for l in dbg[:40]:
                print("   ", l, flush=True)
            print("    (%d debug lines)" % len(dbg), flush=True)
            ok = [l for l in r.stdout.splitlines() if l.startswith("STAGE")]
            if not ok:
                print("---- full stderr of stage %s ----" % s, flush=True)
                print(r.stderr[-6000:], flush=True)
            print(ok[0] if ok else "STAGE %s FAILED: %s" % (s, " | ".join(r.stderr.strip().splitlines()[-6:])[:1500]), flush=True)

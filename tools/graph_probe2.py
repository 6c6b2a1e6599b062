"""Second bisect of the capture failure: which ingredient of the BACKWARD invalidates a CUDA-graph capture?"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
STAGES = ["torch_only", "custom_fn_ctypes", "ours_sum_loss", "ours_no_mt", "ours_tail_only", "ours_backbone_only", "ours_hma_only"]


def stage(name):
    import torch
    import __graft_entry__ as ge
    from editor_b200 import lib
    from editor_b200.train import editor_loss
    torch.cuda.init()
    body = None
    if name == "torch_only":
        net = torch.nn.Sequential(torch.nn.Linear(64, 64), torch.nn.GELU(), torch.nn.Linear(64, 8)).cuda()
        xin = torch.randn(16, 64, device="cuda")

        def body():
            net.zero_grad(set_to_none=False)
            l = net(xin).square().sum()
            l.backward()
            return l
    elif name == "custom_fn_ctypes":
        lib.load()

        class Scale(torch.autograd.Function):
            @staticmethod
            def forward(ctx, x, a):
                y = torch.empty_like(x)
                lib.call("edb_scale_by", x.data_ptr(), a.data_ptr(), y.data_ptr(), x.numel(), lib.stream_ptr())
                ctx.a = a
                return y

            @staticmethod
            def backward(ctx, g):
                d = torch.empty_like(g)
                lib.call("edb_scale_by", g.contiguous().data_ptr(), ctx.a.data_ptr(), d.data_ptr(), g.numel(), lib.stream_ptr())
                return d, None
        w = torch.randn(1024, device="cuda", requires_grad=True)
        a = torch.full((1,), 3.0, device="cuda")

        def body():
            w.grad = None
            l = Scale.apply(w * 2.0, a).sum()
            l.backward()
            return l
    else:
        model, sd, x, label, cam, _ = ge._small_case(True, 4)
        model = model.cuda().train()
        x = {k: v.cuda() for k, v in x.items()}
        label, cam = label.cuda(), cam.cuda()
        eng = model.engine()

        def body():
            model.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                outs = model(x, label=label, cam_label=cam, writer=None, epoch=1)
            if name in ("ours_sum_loss", "ours_no_mt"):
                l = sum(o.float().sum() for o in outs)
            elif name == "ours_tail_only":       # gradient reaches the tail linears / BN only through `ori` (AL head)
                l = outs[2].float().sum()
            elif name == "ours_backbone_only":
                l = outs[3].float().sum()        # ori = cat of the backbone cls tokens
            else:
                l = outs[1].float().sum() + outs[4].float()   # cls4t + aux: through REDUCE, HMA, backbone
            l.backward()
            return l
    if name == "ours_no_mt":
        torch.autograd.set_multithreading_enabled(False)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            body()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = body()
    g.replay()
    torch.cuda.synchronize()
    print("STAGE %s OK %s" % (name, float(out)))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        stage(sys.argv[1])
    else:
        for s in STAGES:
            r = subprocess.run([sys.executable, __file__, s], capture_output=True, text=True, timeout=600)
            ok = [l for l in r.stdout.splitlines() if l.startswith("STAGE")]
            print(ok[0] if ok else "STAGE %s FAILED: %s" % (s, " | ".join(r.stderr.strip().splitlines()[-4:])[:600]), flush=True)

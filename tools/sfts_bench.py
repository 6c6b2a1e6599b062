"""SFTS isolation (BASELINE.json configs[3], faithful variant of SURVEY.md 8(d)-4): B=512 synthetic attention maps,
images and tokens; sweeps FREQUENCY_KEEP and HEAD_KEEP; reports achieved HBM GB/s of each SFTS kernel against the
measured copy bandwidth.  Usage (GPU box): python tools/sfts_bench.py [--batch 512] > gpurun_out/sfts_bench.json"""
import argparse
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from editor_b200 import lib  # noqa: E402


def timed(fn, flush, n=5):
    fn()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / n


def run(batch=512, maps_dtype="bf16", sweep=((10, 2), (16, 1), (32, 2), (64, 4), (96, 8))):
    """Returns the isolation report as a dict (bench.py embeds it as `sfts_isolation`)."""
    args = argparse.Namespace(batch=batch, maps=maps_dtype)
    B, S, H, L = args.batch, 3 * args.batch, 12, 12
    dev = "cuda"
    peak = 6459.6
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                           "MEASURED_PEAKS.json")))["hbm_gbs"]
    except OSError:
        pass
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
    mdt = torch.bfloat16 if args.maps == "bf16" else torch.float32
    g = torch.Generator(device=dev).manual_seed(0)
    maps = []
    for l in range(L):
        m = torch.zeros(S * H, 129, 136, dtype=mdt, device=dev)
        for c0 in range(0, S * H, 1024):
            blk = torch.softmax(torch.randn(min(1024, S * H - c0), 129, 129, device=dev, generator=g) * 2, -1)
            m[c0:c0 + blk.shape[0], :, :129] = blk.to(mdt)
        maps.append(m)
    imgs = [torch.round(torch.clamp(0.25 * torch.randn(B, 3, 256, 128, device=dev, generator=g) + 0.5, 0, 1) * 255) / 127.5 - 1
            for _ in range(3)]
    tokens = torch.randn(3, B, 129, 768, device=dev, generator=g)
    counts = torch.empty(B, 128, dtype=torch.int32, device=dev)
    index = torch.empty(B, 4, dtype=torch.int32, device=dev)
    seq_off = torch.empty(B + 1, dtype=torch.int32, device=dev)
    seq_off3 = torch.empty(B + 1, dtype=torch.int32, device=dev)
    packed = torch.empty(3, B * 129, 768, device=dev)
    loss = torch.zeros(1, device=dev)
    arr = (ctypes.c_void_p * L)(*[m.data_ptr() for m in maps])
    st = lib.stream_ptr()
    esz = 2 if args.maps == "bf16" else 4
    out = {"batch": B, "maps_dtype": args.maps, "hbm_peak_gbs": peak, "runs": []}
    for fk, hk in sweep:
        def f_freq():
            lib.call("edb_freq_counts", imgs[0].data_ptr(), imgs[1].data_ptr(), imgs[2].data_ptr(), B, 256, 128,
                     counts.data_ptr(), st)
            lib.call("edb_topk_mask", counts.data_ptr(), 0, 128, B, 128, fk, index.data_ptr(), 0, st)

        def f_roll():
            lib.call("edb_rollout_topk", arr, L, int(args.maps != "bf16"), S, B, H, 129, 136, hk, index.data_ptr(), None,
                     None, st)

        def f_pack():
            lib.call("edb_index_finalize", index.data_ptr(), B, seq_off.data_ptr(), seq_off3.data_ptr(), st)
            lib.call("edb_sfts_pack_fwd", tokens.data_ptr(), index.data_ptr(), seq_off.data_ptr(), B, B * 129,
                     packed.data_ptr(), loss.data_ptr(), st)
        t_freq = timed(f_freq, flush)
        f_freq()
        t_roll = timed(f_roll, flush)
        t_pack = timed(f_pack, flush)
        f_freq(); f_roll(); f_pack()
        torch.cuda.synchronize()
        n_sel = float((seq_off[-1].item() - B) / B)
        b_freq = 3 * 3 * 256 * 128 * 4 * B + 128 * 4 * B
        b_roll = S * H * ((L - 1) * 129 * 129 + 129) * esz
        b_pack = 3 * B * 129 * 768 * 4 + 3 * B * (1 + n_sel) * 768 * 4
        tot_b, tot_t = b_freq + b_roll + b_pack, t_freq + t_roll + t_pack
        out["runs"].append({"frequency_keep": fk, "head_keep": hk, "kept_tokens_mean": n_sel,
                            "bytes": {"freq_mask": b_freq, "rollout_topk": b_roll, "pack_bcc": b_pack},
                            "freq_mask": {"ms": t_freq, "GBps": b_freq / t_freq / 1e6, "frac": b_freq / t_freq / 1e6 / peak},
                            "rollout_topk": {"ms": t_roll, "GBps": b_roll / t_roll / 1e6, "frac": b_roll / t_roll / 1e6 / peak},
                            "pack_bcc": {"ms": t_pack, "GBps": b_pack / t_pack / 1e6, "frac": b_pack / t_pack / 1e6 / peak},
                            "sfts_total": {"ms": tot_t, "bytes": tot_b, "GBps": tot_b / tot_t / 1e6,
                                           "frac": tot_b / tot_t / 1e6 / peak}})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--maps", default="bf16")
    args = ap.parse_args()
    print(json.dumps(run(args.batch, args.maps)))


if __name__ == "__main__":
    main()

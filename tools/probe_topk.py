"""GPU probe: what does torch.topk on CUDA do with ties?  (SURVEY.md D7 / hard part 1.)

Writes gpurun_out/topk_probe.json.  The candidate rule -- "every element strictly greater than the k-th value,
then elements equal to it in ascending index order" -- is checked on int32 / fp32 / bf16 rows with heavy ties.
"""
import json
import os

import torch


def rule_mask(x, k):
    """Candidate rule, vectorised on CPU: returns bool mask [R, N]."""
    xs = x.double()
    kth = torch.sort(xs, dim=1, descending=True).values[:, k - 1:k]
    gt = xs > kth
    eq = xs == kth
    need = k - gt.sum(1, keepdim=True)
    rank_eq = torch.cumsum(eq.long(), 1)
    return gt | (eq & (rank_eq <= need))


def rule_mask_desc(x, k):
    """Alternative: ties filled from the highest index down."""
    return torch.flip(rule_mask(torch.flip(x, [1]), k), [1])


def main():
    out = {"torch": torch.__version__, "device": torch.cuda.get_device_name(0), "cases": []}
    g = torch.Generator(device="cpu").manual_seed(0)
    cases = []
    for k in (1, 2, 10, 16, 64, 96):
        cases.append(("int32_ties", torch.randint(100, 150, (4096, 128), generator=g, dtype=torch.int32), k))
        cases.append(("int32_all_equal", torch.full((64, 128), 256, dtype=torch.int32), k))
        cases.append(("fp32_ties", torch.randint(0, 20, (4096, 128), generator=g).float() / 16.0, k))
        cases.append(("fp32_notie", torch.rand(4096, 128, generator=g), k))
        cases.append(("bf16_ties", torch.rand(4096, 128, generator=g).to(torch.bfloat16), k))
        cases.append(("fp16_ties", torch.rand(4096, 128, generator=g).to(torch.float16), k))
    for name, x, k in cases:
        idx = torch.topk(x.cuda(), k, dim=1).indices.cpu()
        m = torch.zeros(x.shape, dtype=torch.bool)
        m.scatter_(1, idx, True)
        asc = (m == rule_mask(x, k)).all(1)
        desc = (m == rule_mask_desc(x, k)).all(1)
        idx_cpu = torch.topk(x, k, dim=1).indices
        mc = torch.zeros(x.shape, dtype=torch.bool)
        mc.scatter_(1, idx_cpu, True)
        out["cases"].append({"name": name, "k": k, "rows": x.shape[0],
                             "cuda_matches_asc_rule": int(asc.sum()), "cuda_matches_desc_rule": int(desc.sum()),
                             "cpu_matches_cuda": int((m == mc).all(1).sum()),
                             "cpu_matches_asc_rule": int((mc == rule_mask(x, k)).all(1).sum())})
    # one small explicit example for the docs
    x = torch.tensor([[5, 7, 7, 3, 7, 7, 1, 7]], dtype=torch.int32)
    out["example_k3"] = {"x": x.tolist(), "cuda_idx": torch.topk(x.cuda(), 3, dim=1).indices.cpu().tolist(),
                         "cpu_idx": torch.topk(x, 3, dim=1).indices.tolist()}
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/topk_probe.json", "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out)[:3000])


if __name__ == "__main__":
    main()

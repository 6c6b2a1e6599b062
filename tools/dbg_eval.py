import numpy as np, torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from editor_b200 import lib, metrics as M
from oracle import eval_oracle as eo
d = np.array([[0.1, 0.2, 0.3], [0.3, 0.2, 0.1]], dtype=np.float32)
dev = torch.device("cuda")
dist = torch.from_numpy(d).to(dev)
qp, gp = torch.tensor([7, 9], device=dev), torch.tensor([7, 8, 7], device=dev)
qk, gk = torch.tensor([0, 0], device=dev), torch.tensor([1, 1, 1], device=dev)
ap = torch.full((2,), -5.0, dtype=torch.float64, device=dev); first = torch.full((2,), -7, dtype=torch.int32, device=dev)
over = torch.zeros(1, dtype=torch.int32, device=dev)
lib.call("edb_eval_rank", dist.data_ptr(), dist.stride(0), 2, 3, qp.data_ptr(), gp.data_ptr(), qk.data_ptr(), gk.data_ptr(),
         ap.data_ptr(), first.data_ptr(), over.data_ptr(), lib.stream_ptr())
torch.cuda.synchronize()
print("ap", ap.tolist(), "first", first.tolist(), "over", over.tolist())
print("M.eval_func", M.eval_func(d, np.array([7, 9]), np.array([7, 8, 7]), np.array([0, 0]), np.array([1, 1, 1])))
print("oracle    ", eo.eval_func(d, np.array([7, 9]), np.array([7, 8, 7]), np.array([0, 0]), np.array([1, 1, 1])))
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ref_eval.npz"))
nq = int(g["small_num_query"]); pids, cams = g["small_pids"], g["small_cams"]
cmc, m = M.eval_func(g["small_dist"], pids[:nq], pids[nq:], cams[:nq], cams[nq:])
print("gpu  ", m, cmc[:8], cmc.dtype, cmc.shape)
print("gold ", float(g["small_mAP"]), g["small_cmc"][:8], g["small_cmc"].dtype, g["small_cmc"].shape)
f = M.normalize_(torch.from_numpy(g["small_feats"]).cuda())
dd = M.distmat_device(f[:nq], f[nq:]).cpu().numpy()
print("dist err", np.abs(dd - g["small_dist"]).max(), dd[0, :4], g["small_dist"][0, :4])

"""GPU mirror of the reference's retrieval evaluation (utils/metrics.py) -- SURVEY.md section 8, row f-3.

Same names, arguments and return values as the reference so that engine/processor.py:52-55,144-150 keeps working:
``R1_mAP_eval(num_query, max_rank, feat_norm)`` / ``R1_mAP(...)`` with ``reset() / update() / compute()``,
``euclidean_distance(qf, gf)``, ``eval_func(...)``, ``eval_func_msrv(...)``.  The arithmetic runs in
``libeditor_b200.so`` (csrc/evalrank.cu): features stay on the device (the reference copies every batch to the host,
utils/metrics.py:257), the distance matrix is an fp32 CUDA kernel and CMC / AP come from a rank-counting kernel instead of
``np.argsort`` + a Python loop over the queries.  No CPU fallback: without a CUDA device these raise ``EdbError``.
k-reciprocal re-ranking (utils/reranking.py; TEST.RE_RANKING is 'no' in every shipped config) is out of scope.
"""
import numpy as np
import torch

from . import lib


def _dev():
    if not torch.cuda.is_available():
        raise lib.EdbError("editor_b200.metrics runs on a CUDA device (sm_100a) only; there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _f32(x, dev):
    t = torch.as_tensor(x)
    return t.to(device=dev, dtype=torch.float32).contiguous()


def _ids(x, dev):
    t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))
    return t.to(device=dev, dtype=torch.int64).contiguous()


def normalize_(feats):
    """In-place F.normalize(feats, dim=1, p=2) (utils/metrics.py:255-256)."""
    lib.call("edb_eval_normalize", feats.data_ptr(), feats.stride(0), feats.shape[0], feats.shape[1], 1e-12,
             lib.stream_ptr())
    return feats


def distmat_device(qf, gf):
    """Squared euclidean distances [Q, G] on the device (utils/metrics.py:12-18)."""
    dist = torch.empty(qf.shape[0], gf.shape[0], dtype=torch.float32, device=qf.device)
    lib.call("edb_eval_distmat", qf.data_ptr(), qf.stride(0), qf.shape[0], gf.data_ptr(), gf.stride(0), gf.shape[0],
             qf.shape[1], dist.data_ptr(), dist.stride(0), lib.stream_ptr())
    return dist


def euclidean_distance(qf, gf):
    """utils/metrics.py:12-18 -- returns a numpy array like the reference."""
    dev = _dev()
    return distmat_device(_f32(qf, dev), _f32(gf, dev)).cpu().numpy()


def _rank(dist, q_pids, g_pids, q_keys, g_keys, max_rank):
    dev = dist.device
    Q, G = dist.shape
    if G < max_rank:                                             # utils/metrics.py:141-143
        max_rank = G
        print("Note: number of gallery samples is quite small, got {}".format(G))
    ap = torch.empty(Q, dtype=torch.float64, device=dev)
    first = torch.empty(Q, dtype=torch.int32, device=dev)
    over = torch.zeros(1, dtype=torch.int32, device=dev)
    qp, gp, qk, gk = _ids(q_pids, dev), _ids(g_pids, dev), _ids(q_keys, dev), _ids(g_keys, dev)   # kept alive for the call
    lib.call("edb_eval_rank", dist.data_ptr(), dist.stride(0), Q, G, qp.data_ptr(), gp.data_ptr(), qk.data_ptr(),
             gk.data_ptr(), ap.data_ptr(), first.data_ptr(), over.data_ptr(), lib.stream_ptr())
    if int(over.item()) != 0:
        raise lib.EdbError("eval_rank: a query has more than 2048 correct gallery matches")
    valid = first > 0
    n_valid = int(valid.sum().item())
    assert n_valid > 0, "Error: all query identities do not appear in gallery"       # utils/metrics.py:186
    ranks = torch.arange(1, max_rank + 1, device=dev, dtype=torch.int32)
    hits = ((first[valid, None] <= ranks[None, :])).sum(0)
    # :188-189 -- float32 counts divided by the number of valid queries, on the host with numpy like the reference (torch
    # divides by a scalar as a multiplication by its reciprocal: 7/30 came out one ulp off)
    cmc = hits.cpu().numpy().astype(np.float32) / float(n_valid)
    m_ap = float(ap[valid].mean().item())                                            # :190 (fp64)
    return cmc, m_ap


def eval_func(distmat, q_pids, g_pids, q_camids, g_camids, max_rank=50):
    """utils/metrics.py:133-191 (market1501 protocol: same pid AND same camera as the query is removed)."""
    dev = _dev()
    return _rank(_f32(distmat, dev), q_pids, g_pids, q_camids, g_camids, max_rank)


def eval_func_msrv(distmat, q_pids, g_pids, q_camids, g_camids, q_sceneids, g_sceneids, max_rank=50):
    """utils/metrics.py:36-130 (MSVR310 protocol: same pid AND same scene as the query is removed)."""
    dev = _dev()
    return _rank(_f32(distmat, dev), q_pids, g_pids, q_sceneids, g_sceneids, max_rank)


class R1_mAP_eval:
    """utils/metrics.py:239-283."""

    def __init__(self, num_query, max_rank=20, feat_norm=True, reranking=False):
        if reranking:
            raise NotImplementedError("k-reciprocal re-ranking (utils/reranking.py) is outside the accelerated path")
        self.num_query, self.max_rank, self.feat_norm = num_query, max_rank, feat_norm
        self.reset()

    def reset(self):
        self.feats, self.pids, self.camids = [], [], []

    def update(self, output):                       # called once per batch
        feat, pid, camid = output
        self.feats.append(_f32(feat.detach(), _dev()))          # stays on the device (the reference: feat.cpu())
        self.pids.extend(np.asarray(pid))
        self.camids.extend(np.asarray(camid))

    def _normalizes(self):
        """utils/metrics.py:263 tests `if self.feat_norm:` -- ANY non-empty string (also TEST.FEAT_NORM 'no') normalises;
        R1_mAP (:217) compares with 'yes'.  Mirrored as written."""
        return bool(self.feat_norm)

    def _split(self):
        feats = torch.cat(self.feats, dim=0)
        if self._normalizes():
            print("The test feature is normalized")
            feats = normalize_(feats.clone())
        return feats[:self.num_query], feats[self.num_query:]

    def compute(self, vis=0):                        # called after each epoch
        qf, gf = self._split()
        pids, cams = np.asarray(self.pids), np.asarray(self.camids)
        nq = self.num_query
        print("=> Computing DistMat with euclidean_distance")
        dist = distmat_device(qf, gf)
        cmc, m_ap = _rank(dist, pids[:nq], pids[nq:], cams[:nq], cams[nq:], 50)      # eval_func default max_rank (:133)
        return cmc, m_ap, dist.cpu().numpy(), self.pids, self.camids, qf, gf


class R1_mAP(R1_mAP_eval):
    """utils/metrics.py:193-237 (MSVR310: scene ids, eval_func_msrv)."""

    def __init__(self, num_query, max_rank=50, feat_norm="yes"):
        super().__init__(num_query, max_rank, feat_norm)

    def _normalizes(self):
        return self.feat_norm == "yes"                 # utils/metrics.py:217

    def reset(self):
        super().reset()
        self.sceneids, self.img_path = [], []

    def update(self, output):
        feat, pid, camid, sceneid, img_path = output
        super().update((feat, pid, camid))
        self.sceneids.extend(np.asarray(torch.as_tensor(sceneid).cpu()))
        self.img_path.extend(img_path if img_path is not None else [])

    def compute(self, cfg=None):
        qf, gf = self._split()
        pids, cams, scenes = np.asarray(self.pids), np.asarray(self.camids), np.asarray(self.sceneids)
        nq = self.num_query
        dist = distmat_device(qf, gf)
        cmc, m_ap = _rank(dist, pids[:nq], pids[nq:], scenes[:nq], scenes[nq:], 50)
        return cmc, m_ap, dist.cpu().numpy(), self.pids, self.camids, qf, gf

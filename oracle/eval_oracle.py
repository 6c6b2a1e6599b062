"""TEST INFRASTRUCTURE ONLY -- CPU (numpy) restatement of the reference's retrieval evaluation, SURVEY.md section 8 row f-3.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file; nothing in editor_b200/ does.
Pinned: tests/test_eval_oracle_golden.py checks it against tests/golden/ref_eval.npz, which
tests/golden/make_golden_eval.py produced by calling the UNMODIFIED utils/metrics.py of the reference.

Follows:
  euclidean_distance   utils/metrics.py:12-18     dist = |q|^2 + |g|^2 - 2 q.g^T           (fp32, torch addmm)
  R1_mAP_eval.compute  utils/metrics.py:251-283   F.normalize(feats, dim=1, p=2) -> split query / gallery -> distmat
  eval_func            utils/metrics.py:133-191   market1501 protocol: gallery items with the query's pid AND camid are removed
  eval_func_msrv       utils/metrics.py:36-130    MSVR310 protocol: gallery items with the query's pid AND sceneid are removed

Tie order: the reference ranks with np.argsort(distmat) (default introsort; the order of exactly equal distances is an
implementation detail of numpy).  The restatement and the CUDA kernel rank equal distances by ascending gallery index
(np.argsort(kind="stable")); the golden data is tie-free, on which both orders coincide.
"""
import numpy as np


def l2_normalize(feats, eps=1e-12):
    """torch.nn.functional.normalize(feats, dim=1, p=2): x / max(|x|_2, eps)   (utils/metrics.py:255-256)."""
    f = np.asarray(feats, dtype=np.float32)
    n = np.sqrt((f.astype(np.float32) ** 2).sum(axis=1, keepdims=True, dtype=np.float32))
    return (f / np.maximum(n, np.float32(eps))).astype(np.float32)


def euclidean_distance(qf, gf):
    """utils/metrics.py:12-18 -- squared euclidean distances (no sqrt), fp32."""
    qf = np.asarray(qf, dtype=np.float32)
    gf = np.asarray(gf, dtype=np.float32)
    qq = (qf * qf).sum(axis=1, keepdims=True, dtype=np.float32)
    gg = (gf * gf).sum(axis=1, keepdims=True, dtype=np.float32)
    return (qq + gg.T - np.float32(2.0) * (qf @ gf.T)).astype(np.float32)


def _rank_metrics(distmat, q_pids, g_pids, remove_fn, max_rank):
    """Shared body of eval_func (:133-191) and eval_func_msrv (:36-130)."""
    num_q, num_g = distmat.shape
    max_rank = min(max_rank, num_g)                                   # :141-143
    indices = np.argsort(distmat, axis=1, kind="stable")              # :144
    matches = (g_pids[indices] == q_pids[:, None]).astype(np.int32)   # :147
    all_cmc, all_ap = [], []
    first_rank = np.full(num_q, -1, dtype=np.int64)
    ap_per_query = np.full(num_q, np.nan, dtype=np.float64)
    for q in range(num_q):
        order = indices[q]
        keep = np.invert(remove_fn(q, order))                         # :159-160
        orig = matches[q][keep]
        if not np.any(orig):                                          # :165-167 query identity absent from the gallery
            continue
        cmc = orig.cumsum()
        cmc[cmc > 1] = 1
        all_cmc.append(cmc[:max_rank])                                # :172
        first_rank[q] = int(np.argmax(orig)) + 1
        num_rel = orig.sum()
        tmp = orig.cumsum() / (np.arange(1, orig.shape[0] + 1) * 1.0)  # :180-182
        ap = (tmp * orig).sum() / num_rel                             # :183-184
        all_ap.append(ap)
        ap_per_query[q] = ap
    assert len(all_ap) > 0, "Error: all query identities do not appear in gallery"
    cmc = np.asarray(all_cmc).astype(np.float32).sum(0) / float(len(all_ap))      # :188-189
    return cmc, float(np.mean(all_ap)), ap_per_query, first_rank


def eval_func(distmat, q_pids, g_pids, q_camids, g_camids, max_rank=50, details=False):
    """utils/metrics.py:133-191."""
    q_pids, g_pids, q_camids, g_camids = map(np.asarray, (q_pids, g_pids, q_camids, g_camids))
    out = _rank_metrics(np.asarray(distmat), q_pids, g_pids,
                        lambda q, order: (g_pids[order] == q_pids[q]) & (g_camids[order] == q_camids[q]), max_rank)
    return out if details else out[:2]


def eval_func_msrv(distmat, q_pids, g_pids, q_camids, g_camids, q_sceneids, g_sceneids, max_rank=50, details=False):
    """utils/metrics.py:36-130 (the rank-list file it writes, :58-59,90-98, is a side effect, not a result)."""
    q_pids, g_pids, q_sceneids, g_sceneids = map(np.asarray, (q_pids, g_pids, q_sceneids, g_sceneids))
    out = _rank_metrics(np.asarray(distmat), q_pids, g_pids,
                        lambda q, order: (g_pids[order] == q_pids[q]) & (g_sceneids[order] == q_sceneids[q]), max_rank)
    return out if details else out[:2]


def r1_map_eval(feats, pids, camids, num_query, feat_norm=True, sceneids=None, max_rank=50):
    """R1_mAP_eval.compute (:251-283) / R1_mAP.compute (:209-237) on a feature matrix [N, F]."""
    feats = np.asarray(feats, dtype=np.float32)
    if feat_norm:
        feats = l2_normalize(feats)
    pids, camids = np.asarray(pids), np.asarray(camids)
    qf, gf = feats[:num_query], feats[num_query:]
    dist = euclidean_distance(qf, gf)
    if sceneids is None:
        cmc, m_ap = eval_func(dist, pids[:num_query], pids[num_query:], camids[:num_query], camids[num_query:], max_rank)
    else:
        sceneids = np.asarray(sceneids)
        cmc, m_ap = eval_func_msrv(dist, pids[:num_query], pids[num_query:], camids[:num_query], camids[num_query:],
                                   sceneids[:num_query], sceneids[num_query:], max_rank)
    return cmc, m_ap, dist

"""Retrieval evaluation on the GPU (editor_b200/metrics.py -> csrc/evalrank.cu) against the reference's
own results (tests/golden/ref_eval.npz) and the numpy oracle (oracle/eval_oracle.py)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_eval.npz"))


def _case(name):
    return {k[len(name) + 1:]: GOLD[k] for k in GOLD.files if k.startswith(name + "_")}


@pytest.mark.parametrize("case", ["small", "wide"])
def test_eval_matches_reference_golden(case):
    from editor_b200 import metrics as M
    g = _case(case)
    nq = int(g["num_query"])
    pids, cams, scenes = g["pids"], g["cams"], g["scenes"]
    # ranking on the reference's own distance matrix: integer logic -> identical CMC, AP in fp64
    cmc, m_ap = M.eval_func(g["dist"], pids[:nq], pids[nq:], cams[:nq], cams[nq:])
    assert np.array_equal(cmc, g["cmc"]) and abs(m_ap - float(g["mAP"])) < 1e-12
    cmc, m_ap = M.eval_func_msrv(g["dist"], pids[:nq], pids[nq:], cams[:nq], cams[nq:], scenes[:nq], scenes[nq:])
    assert np.array_equal(cmc, g["cmc_msrv"]) and abs(m_ap - float(g["mAP_msrv"])) < 1e-12
    # normalise + distance matrix: fp32, only the summation order differs from torch's CPU addmm
    f = M.normalize_(torch.from_numpy(g["feats"]).cuda())
    d = M.distmat_device(f[:nq], f[nq:]).cpu().numpy()
    assert np.abs(d - g["dist"]).max() < 2e-6
    # the evaluator classes, fed in batches like engine/processor.py:136-150 does
    ev = M.R1_mAP_eval(nq, max_rank=50, feat_norm="yes")
    ev.reset()
    for a in range(0, len(pids), 37):
        ev.update((torch.from_numpy(g["feats"][a:a + 37]).cuda(), pids[a:a + 37], cams[a:a + 37]))
    cmc, m_ap, dist, _, _, qf, gf = ev.compute()
    assert dist.shape == g["dist"].shape and qf.shape[0] == nq
    assert np.abs(cmc - g["cmc"]).max() <= 1.0 / nq + 1e-6 and abs(m_ap - float(g["mAP"])) < 2e-3
    ev = M.R1_mAP(nq, max_rank=50, feat_norm="yes")
    ev.reset()
    ev.update((torch.from_numpy(g["feats"]).cuda(), pids, cams, torch.from_numpy(scenes), None))
    cmc, m_ap = ev.compute(None)[:2]
    assert np.abs(cmc - g["cmc_msrv"]).max() <= 1.0 / nq + 1e-6 and abs(m_ap - float(g["mAP_msrv"])) < 2e-3


def test_eval_full_width_vs_oracle():
    """2304-wide features (cls4t, make_model.py:258), a few hundred queries against a few thousand gallery items, many
    correct matches per query (several 256-match passes of the kernel)."""
    from editor_b200 import metrics as M
    from oracle import eval_oracle as eo
    g = np.random.default_rng(5)
    n_ids, per, dim, nq = 6, 420, 2304, 200
    centers = g.normal(size=(n_ids, dim)).astype(np.float32) * 0.05
    pids = np.repeat(np.arange(n_ids), per)
    feats = (centers[pids] + g.normal(size=(len(pids), dim)).astype(np.float32)).astype(np.float32)
    cams = g.integers(0, 5, size=len(pids))
    perm = g.permutation(len(pids))
    pids, feats, cams = pids[perm], feats[perm], cams[perm]
    f = M.normalize_(torch.from_numpy(feats).cuda())
    d = M.distmat_device(f[:nq], f[nq:])
    ref_d = eo.euclidean_distance(eo.l2_normalize(feats)[:nq], eo.l2_normalize(feats)[nq:])
    assert np.abs(d.cpu().numpy() - ref_d).max() < 5e-6
    # same distance matrix on both sides -> exact agreement of the ranking
    dn = d.cpu().numpy()
    cmc, m_ap = M.eval_func(dn, pids[:nq], pids[nq:], cams[:nq], cams[nq:])
    rc, rm = eo.eval_func(dn, pids[:nq], pids[nq:], cams[:nq], cams[nq:])
    assert np.array_equal(cmc, rc) and abs(m_ap - rm) < 1e-12


def test_eval_edge_cases():
    from editor_b200 import lib, metrics as M
    from oracle import eval_oracle as eo
    d = np.array([[0.1, 0.2, 0.3], [0.3, 0.2, 0.1]], dtype=np.float32)
    cmc, m_ap = M.eval_func(d, np.array([7, 9]), np.array([7, 8, 7]), np.array([0, 0]), np.array([1, 1, 1]))   # 9 absent
    assert m_ap == pytest.approx((1.0 + 2.0 / 3.0) / 2.0, abs=1e-15) and cmc.tolist() == [1.0, 1.0, 1.0]
    with pytest.raises(AssertionError):
        M.eval_func(d, np.array([5, 6]), np.array([7, 8, 7]), np.array([0, 0]), np.array([1, 1, 1]))
    cmc, m_ap = M.eval_func(d[:1], np.array([7]), np.array([7, 8, 7]), np.array([1]), np.array([1, 0, 0]))
    assert m_ap == pytest.approx(0.5) and cmc.tolist()[:2] == [0.0, 1.0]      # (always min(max_rank, G) entries)
    # ties: ascending gallery index, like the oracle
    g = np.random.default_rng(9)
    dq = g.integers(0, 6, size=(40, 300)).astype(np.float32)
    gp, qp = g.integers(0, 8, size=300), g.integers(0, 8, size=40)
    gc, qc = g.integers(0, 3, size=300), g.integers(0, 3, size=40)
    cmc, m_ap = M.eval_func(dq, qp, gp, qc, gc)
    rc, rm = eo.eval_func(dq, qp, gp, qc, gc)
    assert np.array_equal(cmc, rc) and abs(m_ap - rm) < 1e-12
    # more correct matches than the kernel holds per query -> loud error, not a wrong number
    big = np.zeros((1, 2100), dtype=np.float32)
    with pytest.raises(lib.EdbError):
        M.eval_func(big, np.array([1]), np.full(2100, 1), np.array([0]), np.full(2100, 1))
    with pytest.raises(NotImplementedError):
        M.R1_mAP_eval(3, reranking=True)

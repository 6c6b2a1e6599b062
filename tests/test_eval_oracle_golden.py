"""The numpy restatement of the reference's retrieval evaluation (oracle/eval_oracle.py) against what
the UNMODIFIED utils/metrics.py returned (tests/golden/ref_eval.npz, written by tests/golden/make_golden_eval.py)."""
import os

import numpy as np
import pytest

from oracle import eval_oracle as eo

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_eval.npz"))


@pytest.mark.parametrize("case", ["small", "wide"])
def test_eval_oracle_matches_reference_metrics(case):
    g = {k[len(case) + 1:]: GOLD[k] for k in GOLD.files if k.startswith(case + "_")}
    nq = int(g["num_query"])
    feats = eo.l2_normalize(g["feats"])
    dist = eo.euclidean_distance(feats[:nq], feats[nq:])
    assert np.abs(dist - g["dist"]).max() < 2e-6                      # fp32 summation order of the matmul only
    pids, cams, scenes = g["pids"], g["cams"], g["scenes"]
    # on the reference's own distance matrix the ranking metrics are reproduced exactly (integer logic + float64 AP)
    cmc, m_ap = eo.eval_func(g["dist"], pids[:nq], pids[nq:], cams[:nq], cams[nq:])
    assert np.array_equal(cmc, g["cmc"]) and abs(m_ap - float(g["mAP"])) < 1e-12
    cmc, m_ap = eo.eval_func_msrv(g["dist"], pids[:nq], pids[nq:], cams[:nq], cams[nq:], scenes[:nq], scenes[nq:])
    assert np.array_equal(cmc, g["cmc_msrv"]) and abs(m_ap - float(g["mAP_msrv"])) < 1e-12
    # end to end from the raw features
    cmc, m_ap, _ = eo.r1_map_eval(g["feats"], pids, cams, nq)
    assert np.abs(cmc - g["cmc"]).max() < 1e-6 + 1.0 / nq and abs(m_ap - float(g["mAP"])) < 2e-3


def test_eval_oracle_edge_cases():
    # a query whose identity is absent from the gallery is skipped; all absent -> the reference's assertion
    d = np.array([[0.1, 0.2, 0.3], [0.3, 0.2, 0.1]], dtype=np.float32)
    cmc, m_ap = eo.eval_func(d, np.array([7, 9]), np.array([7, 8, 7]), np.array([0, 0]), np.array([1, 1, 1]))
    assert m_ap == pytest.approx((1.0 + 2.0 / 3.0) / 2.0) and cmc[0] == 1.0 and len(cmc) == 3
    with pytest.raises(AssertionError):
        eo.eval_func(d, np.array([5, 6]), np.array([7, 8, 7]), np.array([0, 0]), np.array([1, 1, 1]))
    # same pid AND same camera as the query is removed from its ranking (market1501 protocol)
    cmc, m_ap = eo.eval_func(d[:1], np.array([7]), np.array([7, 8, 7]), np.array([1]), np.array([1, 0, 0]))
    assert m_ap == pytest.approx(0.5) and cmc[0] == 0.0 and cmc[1] == 1.0
    # equal distances: ascending gallery index
    d2 = np.array([[0.5, 0.5, 0.5]], dtype=np.float32)
    cmc, m_ap = eo.eval_func(d2, np.array([3]), np.array([4, 3, 3]), np.array([0]), np.array([1, 1, 1]))
    assert cmc.tolist() == [0.0, 1.0, 1.0] and m_ap == pytest.approx((1 / 2 + 2 / 3) / 2)

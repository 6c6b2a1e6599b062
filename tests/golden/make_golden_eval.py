"""Generate tests/golden/ref_eval.npz by calling the UNMODIFIED utils/metrics.py of the reference (/root/reference).

Run in the build container only:   python tests/golden/make_golden_eval.py

Synthetic retrieval problems (seeded): clustered features, several cameras
and scenes so that both removal rules fire, a few query identities that are absent from the gallery (the `continue`
branch, utils/metrics.py:165-167).  Stored: features, ids, and what the reference returns -- distmat
(euclidean_distance, :12-18), cmc / mAP of eval_func (:133-191) and of eval_func_msrv (:36-130).

The reference's eval_func_msrv uses `np.str` (:47), removed in numpy 1.24: the script restores the alias (`np.str = str`)
for the call -- the value it feeds is never used in the result -- and runs in a scratch directory because the function
writes a rank list to ./re.txt.
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("EDITOR_REFERENCE_ROOT", "/root/reference")


def import_reference_metrics():
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, REF)
    import utils.metrics as rm                      # noqa: E402  (the reference's module, unmodified)
    return rm


def make_case(seed, n_ids, per_id, n_query, dim, n_cams, n_scenes, absent):
    g = np.random.default_rng(seed)
    centers = g.normal(size=(n_ids + absent, dim)).astype(np.float32) * 0.35       # overlapping identities
    pids = np.repeat(np.arange(n_ids), per_id)
    feats = centers[pids] + g.normal(size=(len(pids), dim)).astype(np.float32)
    cams = g.integers(0, n_cams, size=len(pids))
    scenes = g.integers(0, n_scenes, size=len(pids))
    perm = g.permutation(len(pids))
    pids, feats, cams, scenes = pids[perm], feats[perm], cams[perm], scenes[perm]
    # queries of identities that never appear in the gallery go first
    a_feats = centers[n_ids:] + g.normal(size=(absent, dim)).astype(np.float32)
    a_pids = np.arange(n_ids, n_ids + absent)
    feats = np.concatenate([a_feats, feats]).astype(np.float32)
    pids = np.concatenate([a_pids, pids]).astype(np.int64)
    cams = np.concatenate([g.integers(0, n_cams, size=absent), cams]).astype(np.int64)
    scenes = np.concatenate([g.integers(0, n_scenes, size=absent), scenes]).astype(np.int64)
    return feats, pids, cams, scenes, n_query + absent


def main():
    rm = import_reference_metrics()
    np.str = str                                  # see the module docstring
    out = {}
    cases = {"small": (1, 12, 9, 30, 32, 3, 2, 2), "wide": (2, 40, 14, 150, 96, 6, 4, 3)}
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            for name, spec in cases.items():
                feats, pids, cams, scenes, nq = make_case(*spec)
                f = torch.nn.functional.normalize(torch.from_numpy(feats), dim=1, p=2)      # metrics.py:255-256
                qf, gf = f[:nq], f[nq:]
                dist = rm.euclidean_distance(qf, gf)                                          # :12-18
                cmc, m_ap = rm.eval_func(dist, pids[:nq], pids[nq:], cams[:nq], cams[nq:])    # :133-191
                cmc_s, m_ap_s = rm.eval_func_msrv(dist, pids[:nq], pids[nq:], cams[:nq], cams[nq:], scenes[:nq], scenes[nq:])
                # fp32 distances of normalised features sit on a 2.4e-7 grid around 2: some rows hold equal values; the
                # reference orders them by numpy's introsort, the oracle / kernel by gallery index -- the stored results are
                # what the reference returned, and tests/test_eval_oracle_golden.py shows the oracle reproduces them
                tie_rows = int(((np.diff(np.sort(dist, axis=1), axis=1) == 0).any(axis=1)).sum())
                out["%s_tie_rows" % name] = np.int64(tie_rows)
                for k, v in dict(feats=feats, pids=pids, cams=cams, scenes=scenes, num_query=np.int64(nq), dist=dist,
                                 cmc=cmc, mAP=np.float64(m_ap), cmc_msrv=cmc_s, mAP_msrv=np.float64(m_ap_s)).items():
                    out["%s_%s" % (name, k)] = v
                print(name, "nq", nq, "ng", len(pids) - nq, "mAP %.6f R1 %.4f | msrv mAP %.6f R1 %.4f" %
                      (m_ap, cmc[0], m_ap_s, cmc_s[0]))
        finally:
            os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, "ref_eval.npz"), **out)


if __name__ == "__main__":
    main()

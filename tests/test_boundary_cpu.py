"""CPU-side checks of the drop-in boundary: C-ABI symbols, state_dict schema, config surface, failure without a GPU."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from editor_b200 import lib
    if not os.path.exists(lib.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    so = ctypes.CDLL(lib.LIB_PATH)
    header = open(os.path.join(ROOT, "include", "editor_b200.h")).read()
    declared = set(re.findall(r"\b(edb_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(so, name), name
    assert declared == set(lib.SIGNATURES), declared ^ set(lib.SIGNATURES)
    assert so.edb_version() >= 100


def test_state_dict_schema_matches_reference_appendix_c():
    import __graft_entry__ as ge
    from editor_b200 import synth
    for al, C, cams in ((True, 171, 4), (False, 50, 8)):
        model, sd, *_ = ge._small_case(al, 2)
        own = model.state_dict()
        ref = synth.state_dict_schema(C, cams, al=al)
        assert list(own.keys()) == list(ref.keys())
        assert len(own) == (222 if al else 216)
        for k, (shape, dtype, _) in ref.items():
            assert tuple(own[k].shape) == shape and own[k].dtype == dtype, k
    n = sum(p.numel() for p in model.parameters() if p.requires_grad)
    assert abs(n / 1e6 - 118.548) < 0.01           # Results/Parameter.png: 118.55 M (RGBNT100 head, AL=0)


def test_config_surface_loads_reference_yml_keys():
    from config import cfg
    from modeling import make_model, build_model
    assert build_model is make_model
    for ds in ("RGBNT201", "RGBNT100", "MSVR310", "Market1501-MM"):
        c = cfg.clone()
        c.merge_from_file(os.path.join(ROOT, "configs", ds, "EDITOR.yml"))
        c.merge_from_list(["MODEL.DROP_PATH", "0.0", "MODEL.HEAD_KEEP", 2])
        assert c.MODEL.FREQUENCY_KEEP == 10 and c.MODEL.HEAD_KEEP == 2 and c.MODEL.DROP_PATH == 0.0
    with pytest.raises(KeyError):
        c.merge_from_list(["MODEL.NOT_A_KEY", 1])


def test_product_path_fails_loudly_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import __graft_entry__ as ge
    from editor_b200 import lib
    model, sd, x, label, cam, _ = ge._small_case(True, 2)
    with pytest.raises(lib.EdbError):
        model.eval()(x, cam_label=cam)


def test_metrics_mirror_fails_loudly_without_cuda_and_keeps_the_reference_surface():
    import inspect
    import numpy as np
    from editor_b200 import lib, metrics as M
    # same names / arguments as utils/metrics.py:12,133,36,193,239 of the reference
    assert list(inspect.signature(M.eval_func).parameters) == ["distmat", "q_pids", "g_pids", "q_camids", "g_camids", "max_rank"]
    assert list(inspect.signature(M.eval_func_msrv).parameters)[:7] == ["distmat", "q_pids", "g_pids", "q_camids", "g_camids",
                                                                        "q_sceneids", "g_sceneids"]
    assert list(inspect.signature(M.R1_mAP_eval.__init__).parameters) == ["self", "num_query", "max_rank", "feat_norm", "reranking"]
    for cls in (M.R1_mAP_eval, M.R1_mAP):
        assert all(hasattr(cls, n) for n in ("reset", "update", "compute"))
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(lib.EdbError):
        M.eval_func(np.zeros((1, 2), np.float32), np.array([1]), np.array([1, 2]), np.array([0]), np.array([1, 1]))
    with pytest.raises(lib.EdbError):
        M.euclidean_distance(np.zeros((1, 4), np.float32), np.zeros((2, 4), np.float32))


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under editor_b200/, modeling/ or config/ may import or execute it."""
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|oracle\.", re.M)
    for top in ("editor_b200", "modeling", "config"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith(".py"):
                    src = open(os.path.join(dirpath, f)).read()
                    assert not pat.search(src), os.path.join(dirpath, f)

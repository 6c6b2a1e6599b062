"""Host-side pieces of bench.py that need no GPU: the GEMM roofline bookkeeping (FLOPs / bytes from the ACTUAL row counts
of the packed HMA GEMMs) and the workload naming."""
import bench


class _Ev:
    def __init__(self, t):
        self.t = t

    def elapsed_time(self, other):
        return other.t - self.t


def test_gemm_roofline_counts_device_side_rows():
    T_dev, T3_dev = 0xA000, 0xB000
    tm = [
        # backbone GEMM: static rows
        {"M": 49536, "N": 3072, "K": 768, "M_dev": None, "K_dev": None, "e0": _Ev(0.0), "e1": _Ev(0.25),
         "out_bytes_per_elem": 4, "key": "49536x3072x768 NN epi1"},
        # packed HMA GEMM: launched with the configuration bound 10624, 6447 real rows on the device
        {"M": 10624, "N": 3072, "K": 768, "M_dev": T_dev, "K_dev": None, "e0": _Ev(1.0), "e1": _Ev(1.05),
         "out_bytes_per_elem": 2, "key": "10624x3072x768 NN epi1"},
        # its wgrad: the reduction length is the device-side count
        {"M": 3072, "N": 768, "K": 10624, "M_dev": None, "K_dev": T_dev, "e0": _Ev(2.0), "e1": _Ev(2.03),
         "out_bytes_per_elem": 4, "key": "3072x768x10624 TT epi4"},
    ]
    r = bench.gemm_roofline(tm, {T_dev: 6447, T3_dev: 3 * 6447}, {"bf16_tflops_sustained": 1000.0}, 0)
    flops = 2.0 * (49536 * 3072 * 768 + 6447 * 3072 * 768 + 3072 * 768 * 6447)
    ms = 0.25 + 0.05 + 0.03
    assert abs(r["achieved"] - flops / (ms * 1e-3) / 1e12) < 1e-6 * r["achieved"]
    assert abs(r["frac"] - r["achieved"] / 1000.0) < 1e-12 and r["launches_per_step"] == 3
    assert abs(r["algorithmic_flop_per_launch_mean"] - flops / 3) < 1.0
    by = {x["shape"]: x for x in r["by_shape"]}
    assert by["10624x3072x768 NN epi1"]["rows_actual"] == [6447, 3072, 768]
    assert by["3072x768x10624 TT epi4"]["rows_actual"] == [3072, 768, 6447]
    bytes_ = (2.0 * (49536 * 768 + 3072 * 768) + 4.0 * 49536 * 3072) + (2.0 * (6447 * 768 + 3072 * 768) + 2.0 * 6447 * 3072) + \
             (2.0 * (3072 * 6447 + 768 * 6447) + 4.0 * 3072 * 768)
    assert abs(r["algorithmic_bytes_per_launch_mean"] - bytes_ / 3) < 1.0
    assert r["traffic"] is None or "not this run" in r["traffic_source"]


def test_workload_names_the_configuration():
    assert bench.workload_name("RGBNT100", "bf16", 128).startswith("RGBNT100 EDITOR.yml ViT-B/16 train step, batch 128")
    assert set(bench.CASES) == {"RGBNT201", "RGBNT100", "MSVR310"}
    assert abs(bench.hma_fwd_gflop(128) + 0 - 11.570024448) < 1e-6        # dense HMA forward

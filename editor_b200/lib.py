"""ctypes binding of libeditor_b200.so (C ABI in include/editor_b200.h).

The product path has no fallback: if the shared library is missing or a call fails this raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# EDB_LIB selects another build of the same library (e.g. the -DEDB_MBAR_TIMEOUT debug copy, csrc/Makefile)
LIB_PATH = os.environ.get("EDB_LIB") or os.path.join(_HERE, "lib", "libeditor_b200.so")

EPI_STORE, EPI_GELU, EPI_RESIDUAL, EPI_GELU_BWD, EPI_ATOMIC = 0, 1, 2, 3, 4
PREC_BF16, PREC_FP32 = 0, 1

c_int, c_ll, c_float, c_vp, c_sz = ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t


class EdbError(RuntimeError):
    pass


class GemmDesc(ctypes.Structure):
    _fields_ = [
        ("M", c_int), ("N", c_int), ("K", c_int),
        ("A", c_vp), ("lda", c_ll), ("a_mn_major", c_int),
        ("B", c_vp), ("ldb", c_ll), ("b_mn_major", c_int),
        ("D", c_vp), ("ldd", c_ll), ("out_f32", c_int),
        ("epilogue", c_int),
        ("bias", c_vp),
        ("aux", c_vp), ("ld_aux", c_ll), ("aux_f32", c_int),
        ("out2", c_vp), ("ld_out2", c_ll),
        ("alpha", c_float),
        ("split_k", c_int),
        ("row_scale", c_vp), ("scale_group", c_int),
        ("M_dev", c_vp), ("K_dev", c_vp),
        ("colsum", c_vp),
    ]


class AttnDesc(ctypes.Structure):
    _fields_ = [
        ("qkv", c_vp), ("ld_qkv", c_ll),
        ("out", c_vp), ("ld_out", c_ll),
        ("P", c_vp), ("p_rows", c_ll), ("ldp", c_ll),
        ("seq_off", c_vp), ("fixed_len", c_int), ("nseq", c_int), ("heads", c_int), ("max_len", c_int),
        ("scale", c_float),
        ("f32", c_int),
        ("impl", c_int),
        ("d_out", c_vp), ("ld_dout", c_ll),
        ("d_qkv", c_vp),
        ("total_rows", c_ll),
    ]


# name -> (restype, argtypes); every symbol include/editor_b200.h declares
SIGNATURES = {
    "edb_version": (c_int, []),
    "edb_last_error": (ctypes.c_char_p, []),
    "edb_gemm_bf16": (c_int, [ctypes.POINTER(GemmDesc), c_vp]),
    "edb_gemm_set_mode": (c_int, [c_int]),
    "edb_layernorm_fwd": (c_int, [c_vp, c_ll, c_vp, c_vp, c_float, c_vp, c_ll, c_int, c_vp, c_vp, c_int, c_int, c_vp,
                                  c_vp]),
    "edb_cast_rows_f32_bf16": (c_int, [c_vp, c_vp, c_int, c_int, c_vp, c_vp]),
    "edb_zero_rows": (c_int, [c_vp, c_ll, c_vp, c_int, c_vp]),
    "edb_layernorm_bwd_workspace_bytes": (c_sz, []),
    "edb_layernorm_bwd": (c_int, [c_vp, c_ll, c_int, c_vp, c_ll, c_vp, c_vp, c_vp, c_vp, c_vp, c_ll, c_vp, c_ll, c_vp,
                                  c_vp, c_vp, c_vp, c_sz, c_int, c_int, c_vp, c_int, c_vp, c_vp]),
    "edb_colsum": (c_int, [c_vp, c_ll, c_int, c_int, c_int, c_vp, c_vp]),
    "edb_cast_f32_bf16": (c_int, [c_vp, c_vp, c_sz, c_vp]),
    "edb_sgd_step": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_sz, c_float, c_float, c_float, c_float, c_float, c_float,
                             c_int, c_vp]),
    "edb_split_bf16x3": (c_int, [c_vp, c_ll, c_int, c_int, c_vp, c_int, c_vp]),
    "edb_patch_im2col": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_vp, c_ll, c_int, c_vp]),
    "edb_embed_assemble": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_float, c_int, c_int, c_int, c_vp, c_vp]),
    "edb_embed_assemble_bwd": (c_int, [c_vp, c_int, c_int, c_int, c_vp, c_float, c_vp, c_vp, c_vp, c_int, c_vp]),
    "edb_gelu_bwd_f32": (c_int, [c_vp, c_vp, c_vp, c_sz, c_vp]),
    "edb_attention_fwd": (c_int, [ctypes.POINTER(AttnDesc), c_vp]),
    "edb_attention_bwd": (c_int, [ctypes.POINTER(AttnDesc), c_vp]),
    "edb_freq_counts": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_vp, c_vp]),
    "edb_topk_mask": (c_int, [c_vp, c_int, c_ll, c_int, c_int, c_int, c_vp, c_int, c_vp]),
    "edb_rollout_topk": (c_int, [ctypes.POINTER(c_vp), c_int, c_int, c_int, c_int, c_int, c_ll, c_ll, c_int, c_vp, c_vp,
                                 c_vp, c_vp]),
    "edb_index_finalize": (c_int, [c_vp, c_int, c_vp, c_vp, c_vp]),
    "edb_sfts_pack_fwd": (c_int, [c_vp, c_vp, c_vp, c_int, c_ll, c_vp, c_vp, c_vp]),
    "edb_sfts_pack_bwd": (c_int, [c_vp, c_vp, c_vp, c_int, c_ll, c_vp, c_vp, c_vp, c_vp]),
    "edb_joint_gather": (c_int, [c_vp, c_ll, c_vp, c_vp, c_int, c_int, c_int, c_vp]),
    "edb_pool_fwd": (c_int, [c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_vp]),
    "edb_pool_bwd": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_vp]),
    "edb_cls_rows": (c_int, [c_vp, c_ll, c_vp, c_int, c_vp, c_int, c_vp]),
    "edb_bn1d_fwd": (c_int, [c_vp, c_ll, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_float, c_float, c_vp, c_ll, c_vp, c_vp,
                             c_vp]),
    "edb_bn1d_bwd": (c_int, [c_vp, c_ll, c_vp, c_ll, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_ll, c_vp, c_vp, c_vp]),
    "edb_ocfr_fwd": (c_int, [c_vp, c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_float, c_vp, c_vp, c_vp, c_vp]),
    "edb_ocfr_bwd": (c_int, [c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "edb_ce_smooth": (c_int, [c_vp, c_ll, c_vp, c_int, c_int, c_float, c_vp, c_vp, c_ll, c_vp]),
    "edb_triplet_workspace_bytes": (c_sz, [c_int]),
    "edb_triplet_fwd": (c_int, [c_vp, c_ll, c_vp, c_int, c_int, c_vp, c_vp, c_sz, c_vp]),
    "edb_triplet_bwd": (c_int, [c_vp, c_ll, c_int, c_int, c_vp, c_vp, c_vp, c_ll, c_int, c_vp]),
    "edb_scale_by": (c_int, [c_vp, c_vp, c_vp, c_sz, c_vp]),
    "edb_eval_normalize": (c_int, [c_vp, c_ll, c_int, c_int, c_float, c_vp]),
    "edb_eval_distmat": (c_int, [c_vp, c_ll, c_int, c_vp, c_ll, c_int, c_int, c_vp, c_ll, c_vp]),
    "edb_eval_rank": (c_int, [c_vp, c_ll, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "edb_augment_workspace_bytes": (c_sz, [c_int, c_int, c_int, c_int]),
    "edb_augment_u8": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_int, c_vp, c_vp, c_int,
                               c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
}

_lib = None


def load():
    """Load the shared library once; fail loudly when it has not been built (``__graft_entry__.build()``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EdbError("libeditor_b200.so not built at %s -- run __graft_entry__.build()" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if os.environ.get("EDB_GEMM_MODE"):           # A/B runs: 1 = single-CTA GEMM tiles (see edb_gemm_set_mode)
            lib.edb_gemm_set_mode(int(os.environ["EDB_GEMM_MODE"]))
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise EdbError("editor_b200 call failed (%d): %s" % (rc, load().edb_last_error().decode()))


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return t.data_ptr() if t is not None else None


def _f32(t):
    return int(t.dtype == torch.float32)


launch_count = 0   # kernels launched through the C ABI (bench.py reports it as gpu_launches)
gemm_timing = None  # bench.py sets this to a list to collect one record (shape, device-side row count pointers, CUDA
#                     events) per GEMM launch; FLOPs / bytes are computed from the ACTUAL row counts after the step


_capture_debug = os.environ.get("EDB_CAPTURE_DEBUG")       # tools/graph_probe.py: report the call that invalidates a capture
capture_window = False          # set by tools/graph_probe.py around `with torch.cuda.graph(...)`
_cudart = None


def capture_status(stream=None):
    """cudaStreamIsCapturing of the current stream: 0 = none, 1 = active, 2 = invalidated (debugging aid)."""
    global _cudart
    if _cudart is None:
        for name in ("libcudart.so.12", "libcudart.so"):
            try:
                _cudart = ctypes.CDLL(name)
                break
            except OSError:
                continue
    st = ctypes.c_int(-1)
    rc = _cudart.cudaStreamIsCapturing(ctypes.c_void_p(stream if stream is not None else stream_ptr()), ctypes.byref(st))
    return st.value if rc == 0 else -rc


def call(name, *args):
    global launch_count
    launch_count += 1
    check(getattr(load(), name)(*args))
    if _capture_debug and capture_window:
        st = capture_status()
        if st != 1:
            import threading
            import traceback
            print("EDB_CAPTURE_DEBUG: after %s (launch %d, thread %s) the capture status of stream %#x is %d "
                  "(0 none, 1 active, 2 invalidated)" % (name, launch_count, threading.current_thread().name,
                                                         torch.cuda.current_stream().cuda_stream, st), flush=True)
            if not getattr(call, "_reported", False):
                call._reported = True
                print("".join(traceback.format_stack(limit=10)), flush=True)


def gemm_set_mode(mode):
    """0 = automatic (CTA pairs, tcgen05.mma.cta_group::2), 1 = single-CTA tiles only."""
    check(load().edb_gemm_set_mode(mode))


def gemm(A, B, D, M, N, K, a_mn=False, b_mn=False, epilogue=EPI_STORE, bias=None, aux=None, out2=None,
         alpha=1.0, split_k=1, row_scale=None, scale_group=1, M_dev=None, K_dev=None, colsum=None):
    """D[M,N] = epi(sum_k A(m,k) B(n,k)); A/B bf16 CUDA tensors, D bf16 or fp32 (2-D, row pitch = stride(0))."""
    d = GemmDesc()
    d.M, d.N, d.K = M, N, K
    d.A, d.lda, d.a_mn_major = A.data_ptr(), A.stride(0), int(a_mn)
    d.B, d.ldb, d.b_mn_major = B.data_ptr(), B.stride(0), int(b_mn)
    d.D, d.ldd, d.out_f32 = D.data_ptr(), D.stride(0), _f32(D)
    d.epilogue = epilogue
    d.bias = bias.data_ptr() if bias is not None else None
    if aux is not None:
        d.aux, d.ld_aux, d.aux_f32 = aux.data_ptr(), aux.stride(0), _f32(aux)
    if out2 is not None:
        d.out2, d.ld_out2 = out2.data_ptr(), out2.stride(0)
    d.alpha = alpha
    d.split_k = split_k
    d.row_scale = row_scale.data_ptr() if row_scale is not None else None
    d.scale_group = scale_group
    d.M_dev, d.K_dev = M_dev, K_dev
    d.colsum = colsum.data_ptr() if colsum is not None else None
    if gemm_timing is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        call("edb_gemm_bf16", ctypes.byref(d), stream_ptr())
        e1.record()
        esz = lambda t: 0 if t is None else t.element_size()          # noqa: E731
        gemm_timing.append({"M": M, "N": N, "K": K, "M_dev": M_dev, "K_dev": K_dev, "e0": e0, "e1": e1,
                            "out_bytes_per_elem": esz(D) + esz(aux) + esz(out2),
                            "key": "%dx%dx%d %s%s epi%d%s" % (M, N, K, "T" if a_mn else "N", "T" if b_mn else "N", epilogue,
                                                              " sk%d" % split_k if split_k > 1 else "")})
        return D
    call("edb_gemm_bf16", ctypes.byref(d), stream_ptr())
    return D


def layernorm_fwd(x, gamma, beta, eps, y, mean=None, rstd=None, rows=None, rows_dev=None):
    rows = x.shape[0] if rows is None else rows
    call("edb_layernorm_fwd", x.data_ptr(), x.stride(0), gamma.data_ptr(), beta.data_ptr(), eps, y.data_ptr(),
         y.stride(0), _f32(y), ptr(mean), ptr(rstd), rows, x.shape[1], rows_dev, stream_ptr())
    return y


_ln_ws = {}


def layernorm_bwd(dy, x, mean, rstd, gamma, g_in, g_out, g_bf16, dgamma, dbeta, dcol, rows=None, row_scale=None,
                  scale_group=1, rows_dev=None):
    rows = x.shape[0] if rows is None else rows
    dev = x.device
    ws = _ln_ws.get(dev)
    if ws is None:
        ws = _ln_ws[dev] = torch.empty(load().edb_layernorm_bwd_workspace_bytes(), dtype=torch.uint8, device=dev)
    ldg = (g_out if g_out is not None else x).stride(0)
    call("edb_layernorm_bwd", dy.data_ptr(), dy.stride(0), _f32(dy), x.data_ptr(), x.stride(0), mean.data_ptr(),
         rstd.data_ptr(), gamma.data_ptr(), ptr(g_in), ptr(g_out), ldg, ptr(g_bf16),
         g_bf16.stride(0) if g_bf16 is not None else 0, ptr(dgamma), ptr(dbeta), ptr(dcol), ws.data_ptr(), ws.numel(),
         rows, x.shape[1], ptr(row_scale), scale_group, rows_dev, stream_ptr())


def colsum(src, out, rows=None, n=None):
    call("edb_colsum", src.data_ptr(), src.stride(0), _f32(src), src.shape[0] if rows is None else rows,
         src.shape[1] if n is None else n, out.data_ptr(), stream_ptr())


def cast_bf16(src, dst, n=None):
    call("edb_cast_f32_bf16", src.data_ptr(), dst.data_ptr(), src.numel() if n is None else n, stream_ptr())
    return dst


def split3(src, dst, role, rows=None):
    """src fp32 [rows,K] -> dst bf16 [rows, 6K] (role 0 = A side, 1 = B side) or [6*rows, K] (roles 2 / 3)."""
    call("edb_split_bf16x3", src.data_ptr(), src.stride(0), src.shape[0] if rows is None else rows, src.shape[1],
         dst.data_ptr(), role, stream_ptr())
    return dst


def attention(qkv, out, P, nseq, heads, max_len, scale, seq_off=None, fixed_len=0, p_rows=0, ldp=0, impl=0,
              d_out=None, d_qkv=None, backward=False, total_rows=0):
    d = AttnDesc()
    d.qkv, d.ld_qkv = qkv.data_ptr(), qkv.stride(0)
    if out is not None:
        d.out, d.ld_out = out.data_ptr(), out.stride(0)
    d.P, d.p_rows, d.ldp = ptr(P), p_rows, ldp
    d.seq_off, d.fixed_len, d.nseq, d.heads, d.max_len = ptr(seq_off), fixed_len, nseq, heads, max_len
    d.scale, d.f32, d.impl = scale, _f32(qkv), impl
    d.total_rows = total_rows
    if backward:
        d.d_out, d.ld_dout, d.d_qkv = d_out.data_ptr(), d_out.stride(0), d_qkv.data_ptr()
    call("edb_attention_bwd" if backward else "edb_attention_fwd", ctypes.byref(d), stream_ptr())

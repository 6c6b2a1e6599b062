// Tensor-core (tcgen05/TMEM/TMA) attention over PACKED variable-length sequences: HMA's AttentionMask
// (vit_pytorch.py:240-258) after packing -- per modality 1+n_sel <= 128 tokens, joint 3(1+n_sel) <= 256 tokens per sample
// (masked rows/keys are exact zeros in the reference, so only kept tokens are ever touched).
//
// forward : one CTA per (sequence, head, 128-query tile).  S = Q K^T over KP (=128 or 256) padded keys in TMEM, three-pass
//           softmax straight out of TMEM (max / sum / normalise), P (bf16, zero outside the sequence) to swizzled smem,
//           O = P V, P tile to HBM by TMA for the backward.
// backward: one CTA per (sequence, head), looping over the query tiles; dK / dV accumulate in TMEM across tiles, dS
//           overwrites P in shared memory and is consumed both K-major (dQ = dS K) and MN-major (dK = dS^T Q).
#include "ptx.cuh"
#include "abi_internal.h"

namespace edb {

constexpr int AV_HD = 64;
constexpr uint32_t AV_QTILE = 128 * 128;   // bytes of a [128 lines x 128 B] tile

struct AttnVarParams {
    const __nv_bfloat16* qkv; long long ld_qkv;
    __nv_bfloat16* out; long long ld_out;
    const __nv_bfloat16* d_out; long long ld_dout;
    __nv_bfloat16* d_qkv;
    const int* seq_off;
    int H, p_rows;                 // p_rows = rows of one (seq, head) block of P (multiple of 128)
    float scale, scale_log2e;
    uint32_t idesc_s, idesc_o, idesc_mnmn, idesc_kmn;
};

__device__ __forceinline__ void av_store_row64(__nv_bfloat16* dst, uint32_t taddr, bool valid) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + c * 32, r);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
            for (int t = 0; t < 32; t += 8) {
                uint4 u;
                uint32_t* w = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
                for (int z = 0; z < 4; ++z) {
                    __nv_bfloat162 hb = __floats2bfloat162_rn(__uint_as_float(r[t + 2 * z]), __uint_as_float(r[t + 2 * z + 1]));
                    w[z] = *reinterpret_cast<uint32_t*>(&hb);
                }
                *reinterpret_cast<uint4*>(dst + c * 32 + t) = u;
            }
        }
    }
}

template <int KP>
__global__ void __launch_bounds__(160)
attn_var_fwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
                    const __grid_constant__ CUtensorMap map_p, const AttnVarParams p) {
    constexpr uint32_t KV_BYTES = KP * 128;
    constexpr int NCHUNK = KP / 64;
    constexpr uint32_t TM_COLS = KP == 128 ? 256 : 512;
    constexpr uint32_t TM_O = KP;
    const int sh = blockIdx.x, qt = blockIdx.y;
    const int s = sh / p.H, h = sh % p.H;
    const int off = p.seq_off[s];
    const int L = p.seq_off[s + 1] - off;
    if (qt * 128 >= L) return;

    extern __shared__ __align__(1024) uint8_t smem_raw[];   // 128B-swizzled tiles need a 1024-byte aligned base
    uint8_t* smem = smem_raw;
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + AV_QTILE;
    uint8_t* sV = sK + KV_BYTES;
    uint8_t* sP = sV + KV_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + NCHUNK * AV_QTILE);
    uint64_t *bar_load = bars, *bar_s = bars + 1, *bar_p = bars + 2, *bar_o = bars + 3;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int HC = p.H * AV_HD;

    if (warp == 4) {
        if (lane == 0) {
            tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_kv); tma_prefetch_desc(&map_p);
            mbar_init(bar_load, 1); mbar_init(bar_s, 1); mbar_init(bar_p, 128); mbar_init(bar_o, 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<TM_COLS>(tmem_ptr);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;

    if (warp == 4) {
        if (lane == 0) {
            mbar_expect_tx(bar_load, AV_QTILE + 2 * KV_BYTES);
            tma_load_2d(sQ, &map_q, bar_load, h * AV_HD, off + qt * 128);
            tma_load_2d(sK, &map_kv, bar_load, HC + h * AV_HD, off);
            tma_load_2d(sV, &map_kv, bar_load, 2 * HC + h * AV_HD, off);
            mbar_wait(bar_load, 0);
            tc_fence_after();
            const uint32_t aq = smem_u32(sQ), ak = smem_u32(sK), av = smem_u32(sV), ap = smem_u32(sP);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                tc_mma_bf16(tmem, make_smem_desc(aq + k * 32, 0, 1024), make_smem_desc(ak + k * 32, 0, 1024), p.idesc_s, k > 0);
            tc_commit(bar_s);
            mbar_wait(bar_p, 0);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < KP / 16; ++k)
                tc_mma_bf16(tmem + TM_O, make_smem_desc(ap + (k >> 2) * AV_QTILE + (k & 3) * 32, 0, 1024),
                            make_smem_desc(av + k * 2048, KV_BYTES, 1024), p.idesc_o, k > 0);
            tc_commit(bar_o);
#pragma unroll
            for (int c = 0; c < NCHUNK; ++c) tma_store_2d(&map_p, sP + c * AV_QTILE, c * 64, sh * p.p_rows + qt * 128);
            tma_store_commit();
            tma_store_wait_read();
        }
    } else {
        const int i = threadIdx.x;
        const int row = qt * 128 + i;
        const bool valid = row < L;
        const uint32_t tS = tmem + (static_cast<uint32_t>(warp * 32) << 16);
        const int nch = (L + 31) >> 5;           // 32-column chunks that contain real keys
        mbar_wait(bar_s, 0);
        tc_fence_after();
        float mx = -INFINITY;
        for (int c = 0; c < nch; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(tS + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int t = 0; t < 32; ++t)
                if (c * 32 + t < L) mx = fmaxf(mx, __uint_as_float(r[t]));
        }
        const float mb = mx * p.scale_log2e;
        float sum = 0.f;
        for (int c = 0; c < nch; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(tS + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int t = 0; t < 32; ++t)
                if (c * 32 + t < L) sum += ex2_approx(__uint_as_float(r[t]) * p.scale_log2e - mb);
        }
        const float inv = valid ? 1.0f / sum : 0.f;
        for (int c = 0; c < KP / 32; ++c) {
            uint32_t r[32];
            if (c < nch) {
                tmem_ld_32x32(tS + c * 32, r);
                tmem_ld_wait();
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint4 u = make_uint4(0u, 0u, 0u, 0u);
                if (c < nch && valid) {
                    uint32_t* w = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const int j = c * 32 + q * 8 + 2 * t;
                        const float a = j < L ? ex2_approx(__uint_as_float(r[q * 8 + 2 * t]) * p.scale_log2e - mb) * inv : 0.f;
                        const float b = j + 1 < L ? ex2_approx(__uint_as_float(r[q * 8 + 2 * t + 1]) * p.scale_log2e - mb) * inv : 0.f;
                        __nv_bfloat162 hb = __floats2bfloat162_rn(a, b);
                        w[t] = *reinterpret_cast<uint32_t*>(&hb);
                    }
                }
                *reinterpret_cast<uint4*>(sP + (c >> 1) * AV_QTILE + sw128(i, (c & 1) * 4 + q)) = u;
            }
        }
        fence_proxy_async_smem();
        mbar_arrive(bar_p);
        mbar_wait(bar_o, 0);
        tc_fence_after();
        av_store_row64(p.out + (size_t)(off + row) * p.ld_out + h * AV_HD, tS + TM_O, valid);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc<TM_COLS>(tmem);
}

template <int KP>
__global__ void __launch_bounds__(160)
attn_var_bwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
                    const __grid_constant__ CUtensorMap map_do, const __grid_constant__ CUtensorMap map_p,
                    const AttnVarParams p) {
    constexpr uint32_t KV_BYTES = KP * 128;
    constexpr int NCHUNK = KP / 64;
    constexpr int NKT = KP / 128;                      // 128-key tiles of dK / dV
    constexpr uint32_t TM_COLS = KP == 128 ? 256 : 512;
    constexpr uint32_t TM_DP = 0, TM_DQ = 0, TM_DV = KP, TM_DK = KP + NKT * 64;
    const int sh = blockIdx.x;
    const int s = sh / p.H, h = sh % p.H;
    const int off = p.seq_off[s];
    const int L = p.seq_off[s + 1] - off;
    if (L <= 0) return;
    const int nqt = (L + 127) >> 7;

    extern __shared__ __align__(1024) uint8_t smem_raw[];   // 128B-swizzled tiles need a 1024-byte aligned base
    uint8_t* smem = smem_raw;
    uint8_t* sQ = smem;
    uint8_t* sdO = sQ + AV_QTILE;
    uint8_t* sK = sdO + AV_QTILE;
    uint8_t* sV = sK + KV_BYTES;
    uint8_t* sP = sV + KV_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + NCHUNK * AV_QTILE);
    uint64_t *bar_kv = bars, *bar_load = bars + 1, *bar_dp = bars + 2, *bar_dv = bars + 3, *bar_ds = bars + 4,
             *bar_dq = bars + 5, *bar_dk = bars + 6, *bar_dqread = bars + 7;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int HC = p.H * AV_HD;

    if (warp == 4) {
        if (lane == 0) {
            tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_kv); tma_prefetch_desc(&map_do); tma_prefetch_desc(&map_p);
            mbar_init(bar_kv, 1); mbar_init(bar_load, 1); mbar_init(bar_dp, 1); mbar_init(bar_dv, 1);
            mbar_init(bar_ds, 128); mbar_init(bar_dq, 1); mbar_init(bar_dk, 1); mbar_init(bar_dqread, 128);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<TM_COLS>(tmem_ptr);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;

    if (warp == 4) {
        if (lane == 0) {
            mbar_expect_tx(bar_kv, 2 * KV_BYTES);
            tma_load_2d(sK, &map_kv, bar_kv, HC + h * AV_HD, off);
            tma_load_2d(sV, &map_kv, bar_kv, 2 * HC + h * AV_HD, off);
            const uint32_t aq = smem_u32(sQ), ak = smem_u32(sK), av = smem_u32(sV), ado = smem_u32(sdO), ap = smem_u32(sP);
            for (int qt = 0; qt < nqt; ++qt) {
                const uint32_t ph = qt & 1;
                if (qt > 0) {
                    mbar_wait(bar_dk, ph ^ 1);        // previous tile's MMAs have finished reading sQ / sdO / sP
                    mbar_wait(bar_dqread, ph ^ 1);    // ... and its dQ has been read out of TMEM
                    tc_fence_after();
                }
                mbar_expect_tx(bar_load, 2 * AV_QTILE + NCHUNK * AV_QTILE);
                tma_load_2d(sQ, &map_q, bar_load, h * AV_HD, off + qt * 128);
                tma_load_2d(sdO, &map_do, bar_load, h * AV_HD, off + qt * 128);
#pragma unroll
                for (int c = 0; c < NCHUNK; ++c)
                    tma_load_2d(sP + c * AV_QTILE, &map_p, bar_load, c * 64, sh * p.p_rows + qt * 128);
                if (qt == 0) mbar_wait(bar_kv, 0);
                mbar_wait(bar_load, ph);
                tc_fence_after();
                // dP = dO V^T
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    tc_mma_bf16(tmem + TM_DP, make_smem_desc(ado + k * 32, 0, 1024), make_smem_desc(av + k * 32, 0, 1024),
                                p.idesc_s, k > 0);
                tc_commit(bar_dp);
                // dV[kt] += P^T dO
#pragma unroll
                for (int kt = 0; kt < NKT; ++kt)
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        tc_mma_bf16(tmem + TM_DV + kt * 64, make_smem_desc(ap + 2 * kt * AV_QTILE + k * 2048, AV_QTILE, 1024),
                                    make_smem_desc(ado + k * 2048, AV_QTILE, 1024), p.idesc_mnmn, (qt > 0 || k > 0) ? 1u : 0u);
                tc_commit(bar_dv);
                mbar_wait(bar_ds, ph);
                tc_fence_after();
                // dQ = dS K
#pragma unroll
                for (int k = 0; k < KP / 16; ++k)
                    tc_mma_bf16(tmem + TM_DQ, make_smem_desc(ap + (k >> 2) * AV_QTILE + (k & 3) * 32, 0, 1024),
                                make_smem_desc(ak + k * 2048, KV_BYTES, 1024), p.idesc_kmn, k > 0);
                tc_commit(bar_dq);
                // dK[kt] += dS^T Q
#pragma unroll
                for (int kt = 0; kt < NKT; ++kt)
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        tc_mma_bf16(tmem + TM_DK + kt * 64, make_smem_desc(ap + 2 * kt * AV_QTILE + k * 2048, AV_QTILE, 1024),
                                    make_smem_desc(aq + k * 2048, AV_QTILE, 1024), p.idesc_mnmn, (qt > 0 || k > 0) ? 1u : 0u);
                tc_commit(bar_dk);
            }
        }
    } else {
        const int i = threadIdx.x;
        const uint32_t tB = tmem + (static_cast<uint32_t>(warp * 32) << 16);
        const int nch = (L + 31) >> 5;
        for (int qt = 0; qt < nqt; ++qt) {
            const uint32_t ph = qt & 1;
            const int row = qt * 128 + i;
            const bool valid = row < L;
            mbar_wait(bar_load, ph);
            mbar_wait(bar_dp, ph);
            tc_fence_after();
            float delta = 0.f;
            for (int c = 0; c < nch; ++c) {
                uint32_t r[32];
                tmem_ld_32x32(tB + TM_DP + c * 32, r);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint4 u = *reinterpret_cast<const uint4*>(sP + (c >> 1) * AV_QTILE + sw128(i, (c & 1) * 4 + q));
                    const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const float2 f = __bfloat1622float2(hh[t]);
                        if (f.x != 0.f) delta += f.x * __uint_as_float(r[q * 8 + 2 * t]);
                        if (f.y != 0.f) delta += f.y * __uint_as_float(r[q * 8 + 2 * t + 1]);
                    }
                }
            }
            mbar_wait(bar_dv, ph);                 // the dV MMAs no longer read P: overwrite it with dS
            for (int c = 0; c < nch; ++c) {
                uint32_t r[32];
                tmem_ld_32x32(tB + TM_DP + c * 32, r);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint4* pa = reinterpret_cast<uint4*>(sP + (c >> 1) * AV_QTILE + sw128(i, (c & 1) * 4 + q));
                    uint4 u = *pa;
                    __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const float2 f = __bfloat1622float2(hh[t]);
                        const float a = f.x != 0.f ? f.x * (__uint_as_float(r[q * 8 + 2 * t]) - delta) * p.scale : 0.f;
                        const float b = f.y != 0.f ? f.y * (__uint_as_float(r[q * 8 + 2 * t + 1]) - delta) * p.scale : 0.f;
                        hh[t] = __floats2bfloat162_rn(a, b);
                    }
                    *pa = u;
                }
            }
            tc_fence_before();
            fence_proxy_async_smem();
            mbar_arrive(bar_ds);
            mbar_wait(bar_dq, ph);
            tc_fence_after();
            av_store_row64(p.d_qkv + (size_t)(off + row) * p.ld_qkv + h * AV_HD, tB + TM_DQ, valid);
            tc_fence_before();
            mbar_arrive(bar_dqread);
        }
        mbar_wait(bar_dk, (nqt - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int kt = 0; kt < NKT; ++kt) {
            const int key = kt * 128 + i;
            if (kt * 128 < L) {
                __nv_bfloat16* krow = p.d_qkv + (size_t)(off + key) * p.ld_qkv + h * AV_HD;
                av_store_row64(krow + 2 * HC, tB + TM_DV + kt * 64, key < L);
                av_store_row64(krow + HC, tB + TM_DK + kt * 64, key < L);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc<TM_COLS>(tmem);
}

template <int KP>
static int launch_var(const EdbAttnDesc& d, bool bwd, cudaStream_t st) {
    const long long W3 = 3LL * d.heads * AV_HD, W1 = (long long)d.heads * AV_HD;
    const int nqt = (d.max_len + 127) / 128;
    if (d.p_rows != nqt * 128 || d.ldp != KP) return edb_set_error(EDB_ERR_SHAPE, "attention(var): P must be [nqt*128][KP] per (seq, head)");
    CUtensorMap mq, mkv, mdo, mp;
    EDB_TRY(make_tmap_bf16(&mq, d.qkv, W3, d.total_rows, d.ld_qkv, 128));
    EDB_TRY(make_tmap_bf16(&mkv, d.qkv, W3, d.total_rows, d.ld_qkv, KP));
    EDB_TRY(make_tmap_bf16(&mp, d.P, KP, (long long)d.nseq * d.heads * d.p_rows, KP, 128));
    AttnVarParams p{};
    p.qkv = (const __nv_bfloat16*)d.qkv; p.ld_qkv = d.ld_qkv; p.out = (__nv_bfloat16*)d.out; p.ld_out = d.ld_out;
    p.d_out = (const __nv_bfloat16*)d.d_out; p.ld_dout = d.ld_dout; p.d_qkv = (__nv_bfloat16*)d.d_qkv;
    p.seq_off = d.seq_off; p.H = d.heads; p.p_rows = (int)d.p_rows;
    p.scale = d.scale; p.scale_log2e = d.scale * 1.4426950408889634f;
    p.idesc_s = make_idesc_bf16(128, KP, 0, 0);
    p.idesc_o = make_idesc_bf16(128, AV_HD, 0, 1);
    p.idesc_mnmn = make_idesc_bf16(128, AV_HD, 1, 1);
    p.idesc_kmn = make_idesc_bf16(128, AV_HD, 0, 1);
    if (!bwd) {
        const int smem = AV_QTILE + 2 * KP * 128 + (KP / 64) * AV_QTILE + 128 + 1024;
        static bool configured = false;
        if (!configured) {
            cudaError_t e = cudaFuncSetAttribute(attn_var_fwd_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return edb_set_error(EDB_ERR_CUDA, cudaGetErrorString(e));
            configured = true;
        }
        attn_var_fwd_kernel<KP><<<dim3(d.nseq * d.heads, nqt), 160, smem, st>>>(mq, mkv, mp, p);
    } else {
        EDB_TRY(make_tmap_bf16(&mdo, d.d_out, W1, d.total_rows, d.ld_dout, 128));
        const int smem = 2 * AV_QTILE + 2 * KP * 128 + (KP / 64) * AV_QTILE + 128 + 1024;
        static bool configured = false;
        if (!configured) {
            cudaError_t e = cudaFuncSetAttribute(attn_var_bwd_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return edb_set_error(EDB_ERR_CUDA, cudaGetErrorString(e));
            configured = true;
        }
        attn_var_bwd_kernel<KP><<<d.nseq * d.heads, 160, smem, st>>>(mq, mkv, mdo, mp, p);
    }
    EDB_CHECK_LAUNCH();
    return EDB_OK;
}

// impl 2: packed var-len sequences of at most 256 tokens, bf16, P stored as [nqt*128][KP] blocks (KP = 128 or 256)
int attention_var(const EdbAttnDesc& d, bool bwd, cudaStream_t st) {
    if (d.nseq <= 0) return EDB_OK;
    if (d.f32 || d.seq_off == nullptr || d.P == nullptr || d.total_rows <= 0 || d.max_len > 256)
        return edb_set_error(EDB_ERR_UNSUPPORTED, "attention(var): needs bf16, seq_off, P, total_rows, max_len <= 256");
    if (d.ldp == 128 && d.max_len <= 128) return launch_var<128>(d, bwd, st);
    if (d.ldp == 256) return launch_var<256>(d, bwd, st);
    return edb_set_error(EDB_ERR_SHAPE, "attention(var): ldp must be 128 (max_len <= 128) or 256");
}

}  // namespace edb

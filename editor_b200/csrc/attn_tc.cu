// Tensor-core attention for the backbone's 129-token sequences (Attention.forward, vit_pytorch.py:184-198) on sm_100a:
// tcgen05.mma with TMEM accumulators, operands staged by TMA, one CTA per (sequence, head), two CTAs per SM.
//
// "129 = 128 + 1": the 128 patch queries form one M=128 MMA tile; the cls query (token 0) is a single row and is
// computed by one warp on CUDA cores from the same shared-memory K/V tiles.  Keys are padded 129 -> 144 (next multiple of
// the MMA K step); padded score columns are masked, padded P columns are exact zeros.
//
// forward:   S[128x144] = Q K^T (4 MMAs, K=64)  ->  softmax in registers (one thread per query row, exp2)  ->
//            P (bf16) to swizzled smem + TMA store to HBM (SFTS reads it: SFTS.py:145-153; the backward reuses it)  ->
//            O[128x64] = P V (9 MMAs over 144 keys, V consumed MN-major straight from its TMA tile)  ->  bf16 rows.
#include "ptx.cuh"
#include "abi_internal.h"

namespace edb {

__device__ __forceinline__ float2 lds_bf162(const uint8_t* p) {
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
}
__device__ __forceinline__ float lds_bf16(const uint8_t* p) { return __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(p)); }

constexpr int AT_L = 129;       // tokens per sequence
constexpr int AT_KP = 144;      // keys padded to a multiple of 16
constexpr int AT_PLD = 136;     // pitch of P in HBM
constexpr int AT_HD = 64;
constexpr uint32_t AT_TILE_Q = 128 * 128;          // bytes: 128 rows x 128 B
constexpr uint32_t AT_TILE_KV = AT_KP * 128;       // 18432
constexpr uint32_t AT_CHUNK_P = 128 * 128;         // one 64-key chunk of P: 128 rows x 128 B
constexpr uint32_t AT_F_SQ = 0, AT_F_SK = AT_TILE_Q, AT_F_SV = AT_F_SK + AT_TILE_KV, AT_F_SP = AT_F_SV + AT_TILE_KV;
constexpr uint32_t AT_F_BAR = AT_F_SP + 3 * AT_CHUNK_P;
constexpr uint32_t AT_F_TOTAL = AT_F_BAR + 128 + 1024;
constexpr uint32_t AT_TM_O = 192;                  // TMEM column of the O accumulator (S occupies [0,144))

struct AttnTcParams {
    const __nv_bfloat16* qkv; long long ld_qkv;
    __nv_bfloat16* out; long long ld_out;
    __nv_bfloat16* P;
    int H;
    float scale_log2e;          // scale * log2(e)
    uint32_t idesc_s, idesc_o;
};

// PERSISTENT: a CTA loops over (sequence, head) items  blockIdx.x, blockIdx.x + gridDim.x, ...  with every shared-memory
// tile released as early as its last reader allows, so that the loads of item n+1 run under the softmax of item n and a
// CTA pays its set-up (barrier init, TMEM allocation, descriptor fetch) once instead of once per item:
//   Q, K  of item n+1 are requested when the S MMA of item n has retired and warp 5 has finished its cls-row scores;
//   V     of item n+1 when the PV MMA of item n has retired and warp 5 has finished its cls-row output;
//   S(n+1) = Q K^T is issued as soon as Q, K(n+1) have landed and P(n) has been written (every thread has read S(n));
//   the P tile is handed back to the softmax threads once its TMA store has read it.
// All barriers complete exactly once per item, so the parity of every wait is (local item index & 1).
constexpr uint32_t AT_QK_BYTES = AT_TILE_Q + AT_TILE_KV, AT_V_BYTES = AT_TILE_KV;

__global__ void __launch_bounds__(192, 2)
attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
                   const __grid_constant__ CUtensorMap map_p, const AttnTcParams p, const int n_items) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];   // 128B-swizzled tiles need a 1024-byte aligned base
    uint8_t* smem = smem_raw;
    uint8_t* sQ = smem + AT_F_SQ;
    uint8_t* sK = smem + AT_F_SK;
    uint8_t* sV = smem + AT_F_SV;
    uint8_t* sP = smem + AT_F_SP;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AT_F_BAR);
    uint64_t *bar_qk = bars, *bar_v = bars + 1, *bar_s = bars + 2, *bar_p = bars + 3, *bar_o = bars + 4,
             *bar_pfree = bars + 5, *bar_kdone = bars + 6, *bar_vdone = bars + 7;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int HC = p.H * AT_HD;
    const int stride = gridDim.x;

    if (warp == 4) {
        if (lane == 0) {
            tma_prefetch_desc(&map_q);
            tma_prefetch_desc(&map_kv);
            tma_prefetch_desc(&map_p);
            mbar_init(bar_qk, 1);
            mbar_init(bar_v, 1);
            mbar_init(bar_s, 1);
            mbar_init(bar_p, 128);
            mbar_init(bar_o, 1);
            mbar_init(bar_pfree, 1);
            mbar_init(bar_kdone, 1);
            mbar_init(bar_vdone, 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<256>(tmem_ptr);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;

    if (warp == 4) {
        if (lane == 0) {
            // ---------------- TMA producer + MMA issuer (one thread)
            auto load_qk = [&](int blk) {
                const int s = blk / p.H, h = blk - s * p.H;
                mbar_expect_tx(bar_qk, AT_QK_BYTES);
                tma_load_2d(sQ, &map_q, bar_qk, h * AT_HD, s * AT_L + 1);
                tma_load_2d(sK, &map_kv, bar_qk, HC + h * AT_HD, s * AT_L);
            };
            auto load_v = [&](int blk) {
                const int s = blk / p.H, h = blk - s * p.H;
                mbar_expect_tx(bar_v, AT_V_BYTES);
                tma_load_2d(sV, &map_kv, bar_v, 2 * HC + h * AT_HD, s * AT_L);
            };
            const uint32_t aq = smem_u32(sQ), ak = smem_u32(sK), ap = smem_u32(sP), av = smem_u32(sV);
            if ((int)blockIdx.x < n_items) {
                load_qk(blockIdx.x);
                load_v(blockIdx.x);
            }
            auto mma_s = [&]() {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    tc_mma_bf16(tmem, make_smem_desc(aq + k * 32, 0, 1024), make_smem_desc(ak + k * 32, 0, 1024), p.idesc_s,
                                k > 0);
                tc_commit(bar_s);
            };
            uint32_t ph = 0;
            if ((int)blockIdx.x < n_items) {
                mbar_wait(bar_qk, 0);
                tc_fence_after();
                mma_s();                          // S(0)
            }
            for (int blk = blockIdx.x; blk < n_items; blk += stride, ph ^= 1) {
                const int nxt = blk + stride;
                mbar_wait(bar_s, ph);             // the S MMA has read Q and K ...
                mbar_wait(bar_kdone, ph);         // ... and so has warp 5: both tiles are free
                if (nxt < n_items) load_qk(nxt);
                mbar_wait(bar_p, ph);             // P(n) is in shared memory; O(n-1) and S(n) have been read
                mbar_wait(bar_v, ph);             // V(n) has landed
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < AT_KP / 16; ++k)
                    tc_mma_bf16(tmem + AT_TM_O, make_smem_desc(ap + (k >> 2) * AT_CHUNK_P + (k & 3) * 32, 0, 1024),
                                make_smem_desc(av + k * 2048, AT_TILE_KV, 1024), p.idesc_o, k > 0);
                tc_commit(bar_o);
                // S(n+1) right behind PV(n): its TMEM columns are free (every thread arrived on bar_p(n) after reading
                // S(n)), so the next item's scores are ready by the time the softmax threads finish this item's epilogue
                if (nxt < n_items) {
                    mbar_wait(bar_qk, ph ^ 1);
                    tc_fence_after();
                    mma_s();
                }
                // P rows 1..128 of this (seq, head) to HBM; columns >= 136 are clipped by the tensor map
#pragma unroll
                for (int c = 0; c < 3; ++c) tma_store_2d(&map_p, sP + c * AT_CHUNK_P, c * 64, blk * AT_L + 1);
                tma_store_commit();
                mbar_wait(bar_o, ph);             // the PV MMA has read V (and P) ...
                mbar_wait(bar_vdone, ph);         // ... and so has warp 5
                if (nxt < n_items) load_v(nxt);
                tma_store_wait_read();            // the bulk store has read the P tile: the softmax threads may refill it
                mbar_arrive(bar_pfree);
            }
        }
    } else if (warp == 5) {
        // ---------------- cls query (token 0) on CUDA cores
        uint32_t ph = 0;
        for (int blk = blockIdx.x; blk < n_items; blk += stride, ph ^= 1) {
            const int s = blk / p.H, h = blk - s * p.H;
            const int row0 = s * AT_L;
            const __nv_bfloat16* q0 = p.qkv + (size_t)row0 * p.ld_qkv + h * AT_HD;
            float q[AT_HD];
#pragma unroll
            for (int d = 0; d < AT_HD; d += 2) {
                const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(q0 + d));
                q[d] = f.x; q[d + 1] = f.y;
            }
            mbar_wait(bar_qk, ph);
            float sc[5];
            float mx = -INFINITY;
#pragma unroll
            for (int jj = 0; jj < 5; ++jj) {
                const int j = jj * 32 + lane;
                float acc = -INFINITY;
                if (j < AT_L) {
                    acc = 0.f;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const uint4 u = *reinterpret_cast<const uint4*>(sK + sw128(j, c));
                        const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            const float2 f = __bfloat1622float2(hh[t]);
                            acc += q[c * 8 + 2 * t] * f.x + q[c * 8 + 2 * t + 1] * f.y;
                        }
                    }
                }
                sc[jj] = acc;
                mx = fmaxf(mx, acc);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_kdone);       // this warp no longer reads the K tile
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            float sum = 0.f;
#pragma unroll
            for (int jj = 0; jj < 5; ++jj) {
                sc[jj] = (jj * 32 + lane < AT_L) ? ex2_approx((sc[jj] - mx) * p.scale_log2e) : 0.f;
                sum += sc[jj];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const float inv = 1.0f / sum;
            __nv_bfloat16* prow = p.P + (size_t)blk * AT_L * AT_PLD;
#pragma unroll
            for (int jj = 0; jj < 5; ++jj) {
                const int j = jj * 32 + lane;
                const __nv_bfloat16 pb = __float2bfloat16(sc[jj] * inv);
                sc[jj] = __bfloat162float(pb);
                if (j < AT_PLD) prow[j] = pb;
            }
            mbar_wait(bar_v, ph);
            float o0 = 0.f, o1 = 0.f;
#pragma unroll
            for (int jj = 0; jj < 5; ++jj) {
                const int jn = (jj < 4) ? 32 : 1;
                _Pragma("unroll 8") for (int t = 0; t < jn; ++t) {
                    const int j = jj * 32 + t;
                    const float pj = __shfl_sync(0xffffffffu, sc[jj], t);
                    const float2 f = lds_bf162(sV + sw128(j, lane >> 2) + (lane & 3) * 4);
                    o0 += pj * f.x;
                    o1 += pj * f.y;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_vdone);       // this warp no longer reads the V tile
            *reinterpret_cast<__nv_bfloat162*>(p.out + (size_t)row0 * p.ld_out + h * AT_HD + 2 * lane) =
                __floats2bfloat162_rn(o0, o1);
        }
    } else {
        // ---------------- softmax + epilogue: thread i owns query token i+1 (TMEM lane i)
        const int i = threadIdx.x;
        const uint32_t tS = tmem + (static_cast<uint32_t>(warp * 32) << 16);
        uint32_t ph = 0;
        bool first = true;
        for (int blk = blockIdx.x; blk < n_items; blk += stride, ph ^= 1) {
            const int s = blk / p.H, h = blk - s * p.H;
            const int row0 = s * AT_L;
            mbar_wait(bar_s, ph);
            tc_fence_after();
            // all five TMEM loads of the score row are in flight before the one wait (each used to pay the full TMEM
            // round trip on its own: long-scoreboard was the top stall of the softmax warps)
            uint32_t r0[32], r1[32], r2[32], r3[32], r4[16];
            tmem_ld_32x32(tS, r0);
            tmem_ld_32x32(tS + 32, r1);
            tmem_ld_32x32(tS + 64, r2);
            tmem_ld_32x32(tS + 96, r3);
            tmem_ld_32x16(tS + 128, r4);
            tmem_ld_wait();
            float e[132];
            float mx = -INFINITY;
#pragma unroll
            for (int t = 0; t < 32; ++t) {
                e[t] = __uint_as_float(r0[t]);
                e[32 + t] = __uint_as_float(r1[t]);
                e[64 + t] = __uint_as_float(r2[t]);
                e[96 + t] = __uint_as_float(r3[t]);
                mx = fmaxf(fmaxf(mx, e[t]), fmaxf(fmaxf(e[32 + t], e[64 + t]), e[96 + t]));
            }
            e[128] = __uint_as_float(r4[0]);
            mx = fmaxf(mx, e[128]);
            const float mb = mx * p.scale_log2e;
            float sum = 0.f;
#pragma unroll
            for (int t = 0; t < AT_L; ++t) {
                e[t] = ex2_approx(e[t] * p.scale_log2e - mb);
                sum += e[t];
            }
            const float inv = 1.0f / sum;
            e[129] = e[130] = e[131] = 0.f;
            if (!first) mbar_wait(bar_pfree, ph ^ 1);      // the TMA store of the previous item's P has read the tile
            first = false;
#pragma unroll
            for (int q8 = 0; q8 < AT_KP / 8; ++q8) {       // 18 chunks of 8 keys
                uint4 u;
                uint32_t* w = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int c0 = q8 * 8 + 2 * t;
                    const float a = c0 < AT_L ? e[c0 < 132 ? c0 : 131] * inv : 0.f;
                    const float b = c0 + 1 < AT_L ? e[c0 + 1 < 132 ? c0 + 1 : 131] * inv : 0.f;
                    __nv_bfloat162 hb = __floats2bfloat162_rn(a, b);
                    w[t] = *reinterpret_cast<uint32_t*>(&hb);
                }
                *reinterpret_cast<uint4*>(sP + (q8 >> 3) * AT_CHUNK_P + sw128(i, q8 & 7)) = u;
            }
            // orders this thread's TMEM reads (S of this item, O of the previous one) before the MMAs that the arrival
            // below allows to overwrite them, and its P writes before the async-proxy readers (MMA, TMA store)
            tc_fence_before();
            fence_proxy_async_smem();
            mbar_arrive(bar_p);
            mbar_wait(bar_o, ph);
            tc_fence_after();
            __nv_bfloat16* orow = p.out + (size_t)(row0 + 1 + i) * p.ld_out + h * AT_HD;
            uint32_t o0[32], o1[32];
            tmem_ld_32x32(tS + AT_TM_O, o0);
            tmem_ld_32x32(tS + AT_TM_O + 32, o1);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const uint32_t* r = c == 0 ? o0 : o1;
#pragma unroll
                for (int t = 0; t < 32; t += 8) {
                    uint4 u;
                    uint32_t* w = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
                    for (int z = 0; z < 4; ++z) {
                        __nv_bfloat162 hb = __floats2bfloat162_rn(__uint_as_float(r[t + 2 * z]), __uint_as_float(r[t + 2 * z + 1]));
                        w[z] = *reinterpret_cast<uint32_t*>(&hb);
                    }
                    *reinterpret_cast<uint4*>(orow + c * 32 + t) = u;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc<256>(tmem);
}

int attention_tc_fwd(const EdbAttnDesc& d, cudaStream_t st) {
    if (d.nseq <= 0) return EDB_OK;
    const long long R = (long long)d.nseq * AT_L;
    CUtensorMap mq, mkv, mp;
    EDB_TRY(make_tmap_bf16(&mq, d.qkv, 3LL * d.heads * AT_HD, R, d.ld_qkv, 128));
    EDB_TRY(make_tmap_bf16(&mkv, d.qkv, 3LL * d.heads * AT_HD, R, d.ld_qkv, AT_KP));
    EDB_TRY(make_tmap_bf16(&mp, d.P, AT_PLD, (long long)d.nseq * d.heads * AT_L, AT_PLD, 128));
    AttnTcParams p{};
    p.qkv = (const __nv_bfloat16*)d.qkv; p.ld_qkv = d.ld_qkv;
    p.out = (__nv_bfloat16*)d.out; p.ld_out = d.ld_out;
    p.P = (__nv_bfloat16*)d.P; p.H = d.heads;
    p.scale_log2e = d.scale * 1.4426950408889634f;
    p.idesc_s = make_idesc_bf16(128, AT_KP, 0, 0);
    p.idesc_o = make_idesc_bf16(128, AT_HD, 0, 1);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(attn_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_F_TOTAL);
        if (e != cudaSuccess) return edb_set_error(EDB_ERR_CUDA, cudaGetErrorString(e));
        configured = true;
    }
    const int n_items = d.nseq * d.heads;
    const int slots = 2 * num_sms();            // two resident CTAs per SM
    attn_tc_fwd_kernel<<<n_items < slots ? n_items : slots, 192, AT_F_TOTAL, st>>>(mq, mkv, mp, p, n_items);
    EDB_CHECK_LAUNCH();
    return EDB_OK;
}


// ======================================================================================================== backward
// Per (sequence, head), with the saved bf16 P:   dP = dO V^T;  delta_i = sum_j dP_ij P_ij;  dS = P o (dP - delta) * scale;
// dV = P^T dO;  dQ = dS K;  dK = dS^T Q.
//
// Shared-memory tiles are [144 lines x 128 B] (128-byte swizzle).  Query-indexed tiles (Q, dO, and the lines of P/dS) hold
// the queries in the order  tok1..tok128, tok0, 15 zero lines  -- a reduction dimension may be permuted freely, and this
// puts the 128 patch queries on one M=128 tile.  Key-indexed tiles (K, V, columns of P/dS) are in natural order.  The
// same bytes serve as K-major and as MN-major operands: P/dS [q lines x key columns] is the K-major A of dQ = dS K and the
// MN-major A of dV = P^T dO / dK = dS^T Q.  dS overwrites P in place.  The cls query row and key 128 (the "+1"s) are
// handled by two extra warps on CUDA cores.
constexpr uint32_t BT_TILE = AT_KP * 128;                  // 18432
constexpr uint32_t BT_SQ = 0, BT_SK = BT_TILE, BT_SV = 2 * BT_TILE, BT_SDO = 3 * BT_TILE, BT_SP = 4 * BT_TILE;
constexpr uint32_t BT_DSCOL = 6 * BT_TILE;                 // 144 floats: dS[:, key 128] per query line
constexpr uint32_t BT_BAR = BT_DSCOL + 576;
constexpr uint32_t BT_TOTAL = BT_BAR + 128;
// TMEM (256 columns, so that two CTAs share an SM): dP [0,128) is dead once dS is written; dQ and dK then reuse it
constexpr uint32_t BT_TM_DP = 0, BT_TM_DV = 128, BT_TM_DQ = 0, BT_TM_DK = 64;

struct AttnTcBwdParams {
    const __nv_bfloat16* qkv; long long ld_qkv;
    __nv_bfloat16* d_qkv;
    const __nv_bfloat16* P;
    int H;
    float scale;
    uint32_t idesc_dp, idesc_kk_mn, idesc_k_mn;   // (128x128 K/K), (128x64 MN/MN), (128x64 K/MN)
};

// 64 accumulator columns of this thread's TMEM lane -> bf16 row in HBM, optionally plus a rank-1 term a * v[0..63] with v
// a 128-byte row of a swizzled smem tile (the contribution of the key / query that is not on the MMA tile)
__device__ __forceinline__ void store_row64_bf16(__nv_bfloat16* dst, uint32_t taddr, float a = 0.f, const uint8_t* tile = nullptr,
                                                 int line = 0) {
    // rolled on purpose (here and in the two dS passes below): the backward kernel is executed once per CTA, straight
    // through; fully unrolled it was 150 KB of SASS -- more than the instruction cache -- and 17 % of its stall samples
    // were instruction fetches
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int t = 0; t < 32; t += 8) {
            float v[8];
#pragma unroll
            for (int z = 0; z < 8; ++z) v[z] = __uint_as_float(r[t + z]);
            if (tile != nullptr) {
                const uint4 u = *reinterpret_cast<const uint4*>(tile + sw128(line, c * 4 + t / 8));
                const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                for (int z = 0; z < 4; ++z) {
                    const float2 f = __bfloat1622float2(hh[z]);
                    v[2 * z] = fmaf(a, f.x, v[2 * z]);
                    v[2 * z + 1] = fmaf(a, f.y, v[2 * z + 1]);
                }
            }
            uint4 o;
            uint32_t* w = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
            for (int z = 0; z < 4; ++z) {
                __nv_bfloat162 hb = __floats2bfloat162_rn(v[2 * z], v[2 * z + 1]);
                w[z] = *reinterpret_cast<uint32_t*>(&hb);
            }
            *reinterpret_cast<uint4*>(dst + c * 32 + t) = o;
        }
    }
}

// dot product of two 128-byte rows (64 bf16) of swizzled smem tiles
__device__ __forceinline__ float dot_rows64(const uint8_t* ta, int la, const uint8_t* tb, int lb) {
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const uint4 ua = *reinterpret_cast<const uint4*>(ta + sw128(la, c));
        const uint4 ub = *reinterpret_cast<const uint4*>(tb + sw128(lb, c));
        const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&ua);
        const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&ub);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float2 fa = __bfloat1622float2(ha[t]), fb = __bfloat1622float2(hb[t]);
            acc += fa.x * fb.x + fa.y * fb.y;
        }
    }
    return acc;
}

// Query lines: 0..127 = tokens 1..128, 128 = token 0 (cls), 129..143 zero.  Keys 0..127 sit on the MMA tiles; key 128 is
// handled on CUDA cores: its score column by the row-owner threads (one dot product each), its dK / dV row by warp 6.
//
// PERSISTENT like the forward kernel: a CTA loops over items; V and dO of item n+1 are requested as soon as the dP / dV
// MMAs of item n have retired and every CUDA-core reader has passed them ("early" release), Q, K and P when the dQ / dK
// MMAs have retired and the last readers of K row 128 / the Q tile are done ("late" release).  The 15 zero padding lines
// of the query-indexed tiles are written once: the TMA boxes only ever touch lines 0..128.
constexpr uint32_t BT_TX_VDO = BT_TILE + (128 * 128 + 128);                        // V, dO
constexpr uint32_t BT_TX_QKP = (128 * 128 + 128) + BT_TILE + 2 * (128 * 128 + 128);  // Q, K, 2 chunks of P

__global__ void __launch_bounds__(224, 2)
attn_tc_bwd_kernel(const __grid_constant__ CUtensorMap map_q128, const __grid_constant__ CUtensorMap map_q1,
                   const __grid_constant__ CUtensorMap map_kv, const __grid_constant__ CUtensorMap map_do128,
                   const __grid_constant__ CUtensorMap map_do1, const __grid_constant__ CUtensorMap map_p128,
                   const __grid_constant__ CUtensorMap map_p1, const AttnTcBwdParams p, const int n_items) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];   // 128B-swizzled tiles need a 1024-byte aligned base
    uint8_t* smem = smem_raw;
    uint8_t* sQ = smem + BT_SQ;
    uint8_t* sK = smem + BT_SK;
    uint8_t* sV = smem + BT_SV;
    uint8_t* sdO = smem + BT_SDO;
    uint8_t* sP = smem + BT_SP;            // 2 chunks of 64 keys (keys 0..127); becomes dS
    float* dscol = reinterpret_cast<float*>(smem + BT_DSCOL);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BT_BAR);
    uint64_t *bar_vdo = bars, *bar_qkp = bars + 1, *bar_dp = bars + 2, *bar_dv = bars + 3, *bar_ds = bars + 4,
             *bar_dq = bars + 5, *bar_dk = bars + 6, *bar_early = bars + 7, *bar_late = bars + 8, *bar_tmfree = bars + 9;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 10);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int HC = p.H * AT_HD;
    const int stride = gridDim.x;

    // zero the 15 padding lines (129..143) of the query-indexed tiles
    for (int t = threadIdx.x; t < 4 * 15 * 8; t += 224) {
        const int tile = t / 120, rem = t % 120, line = 129 + rem / 8, c = rem % 8;
        uint8_t* base = tile == 0 ? sQ : (tile == 1 ? sdO : sP + (tile - 2) * BT_TILE);
        *reinterpret_cast<uint4*>(base + line * 128 + c * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async_smem();
    if (warp == 4) {
        if (lane == 0) {
            tma_prefetch_desc(&map_q128); tma_prefetch_desc(&map_q1); tma_prefetch_desc(&map_kv);
            tma_prefetch_desc(&map_do128); tma_prefetch_desc(&map_do1); tma_prefetch_desc(&map_p128);
            tma_prefetch_desc(&map_p1);
            mbar_init(bar_vdo, 1);
            mbar_init(bar_qkp, 1);
            mbar_init(bar_dp, 1);
            mbar_init(bar_dv, 1);
            mbar_init(bar_ds, 128 + 1);
            mbar_init(bar_dq, 1);
            mbar_init(bar_dk, 1);
            mbar_init(bar_early, 128 + 2);
            mbar_init(bar_late, 128 + 2);
            mbar_init(bar_tmfree, 128);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<256>(tmem_ptr);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;

    if (warp == 4) {
        if (lane == 0) {
            auto load_vdo = [&](int blk) {
                const int s = blk / p.H, h = blk - s * p.H, row0 = s * AT_L;
                mbar_expect_tx(bar_vdo, BT_TX_VDO);
                tma_load_2d(sdO, &map_do128, bar_vdo, h * AT_HD, row0 + 1);
                tma_load_2d(sdO + 128 * 128, &map_do1, bar_vdo, h * AT_HD, row0);
                tma_load_2d(sV, &map_kv, bar_vdo, 2 * HC + h * AT_HD, row0);
            };
            auto load_qkp = [&](int blk) {
                const int s = blk / p.H, h = blk - s * p.H, row0 = s * AT_L;
                mbar_expect_tx(bar_qkp, BT_TX_QKP);
                tma_load_2d(sQ, &map_q128, bar_qkp, h * AT_HD, row0 + 1);
                tma_load_2d(sQ + 128 * 128, &map_q1, bar_qkp, h * AT_HD, row0);
                tma_load_2d(sK, &map_kv, bar_qkp, HC + h * AT_HD, row0);
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    tma_load_2d(sP + c * BT_TILE, &map_p128, bar_qkp, c * 64, blk * AT_L + 1);
                    tma_load_2d(sP + c * BT_TILE + 128 * 128, &map_p1, bar_qkp, c * 64, blk * AT_L);
                }
            };
            const uint32_t aq = smem_u32(sQ), ak = smem_u32(sK), av = smem_u32(sV), ado = smem_u32(sdO), ap = smem_u32(sP);
            if ((int)blockIdx.x < n_items) {
                load_vdo(blockIdx.x);
                load_qkp(blockIdx.x);
            }
            uint32_t ph = 0;
            bool first = true;
            for (int blk = blockIdx.x; blk < n_items; blk += stride, ph ^= 1) {
                const int nxt = blk + stride;
                mbar_wait(bar_vdo, ph);
                mbar_wait(bar_qkp, ph);
                if (!first) mbar_wait(bar_tmfree, ph ^ 1);   // dQ / dK of the previous item have left TMEM
                first = false;
                tc_fence_after();
                // dP[q 0..127][key 0..127] = dO V^T
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    tc_mma_bf16(tmem + BT_TM_DP, make_smem_desc(ado + k * 32, 0, 1024), make_smem_desc(av + k * 32, 0, 1024),
                                p.idesc_dp, k > 0);
                tc_commit(bar_dp);
                // dV[key 0..127] = P^T dO   (A: P MN-major, reduction over the 144 query lines)
#pragma unroll
                for (int k = 0; k < AT_KP / 16; ++k)
                    tc_mma_bf16(tmem + BT_TM_DV, make_smem_desc(ap + k * 2048, BT_TILE, 1024),
                                make_smem_desc(ado + k * 2048, BT_TILE, 1024), p.idesc_kk_mn, k > 0);
                tc_commit(bar_dv);
                mbar_wait(bar_dv, ph);            // both MMAs have read V / dO ...
                mbar_wait(bar_early, ph);         // ... and so have the row owners, warp 5 and warp 6
                if (nxt < n_items) load_vdo(nxt);
                mbar_wait(bar_ds, ph);
                tc_fence_after();
                // dQ[q 0..127] = dS K over keys 0..127   (key 128 is added by the epilogue threads)
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    tc_mma_bf16(tmem + BT_TM_DQ, make_smem_desc(ap + (k >> 2) * BT_TILE + (k & 3) * 32, 0, 1024),
                                make_smem_desc(ak + k * 2048, BT_TILE, 1024), p.idesc_k_mn, k > 0);
                tc_commit(bar_dq);
                // dK[key 0..127] = dS^T Q   (reduction over the 144 query lines)
#pragma unroll
                for (int k = 0; k < AT_KP / 16; ++k)
                    tc_mma_bf16(tmem + BT_TM_DK, make_smem_desc(ap + k * 2048, BT_TILE, 1024),
                                make_smem_desc(aq + k * 2048, BT_TILE, 1024), p.idesc_kk_mn, k > 0);
                tc_commit(bar_dk);
                mbar_wait(bar_dk, ph);            // the MMAs have read Q, K and dS ...
                mbar_wait(bar_late, ph);          // ... and the CUDA-core readers of K row 128 / the Q tile are done
                if (nxt < n_items) load_qkp(nxt);
            }
        }
    } else if (warp == 5) {
        // ---------------- cls query (token 0 = query line 128) on CUDA cores
        uint32_t ph = 0;
        for (int blk = blockIdx.x; blk < n_items; blk += stride, ph ^= 1) {
            const int s = blk / p.H, h = blk - s * p.H, row0 = s * AT_L;
            const __nv_bfloat16* Pg = p.P + (size_t)blk * AT_L * AT_PLD;
            const float p0_128 = __bfloat162float(Pg[128]);
            mbar_wait(bar_vdo, ph);
            float g[AT_HD];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint4 u = *reinterpret_cast<const uint4*>(sdO + sw128(128, c));
                const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const float2 f = __bfloat1622float2(hh[t]);
                    g[c * 8 + 2 * t] = f.x; g[c * 8 + 2 * t + 1] = f.y;
                }
            }
            float ds[5], pv[5];
#pragma unroll
            for (int jj = 0; jj < 5; ++jj) {
                const int j = jj * 32 + lane;
                float acc = 0.f;
                if (j < AT_L) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const uint4 u = *reinterpret_cast<const uint4*>(sV + sw128(j, c));
                        const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            const float2 f = __bfloat1622float2(hh[t]);
                            acc += g[c * 8 + 2 * t] * f.x + g[c * 8 + 2 * t + 1] * f.y;
                        }
                    }
                }
                ds[jj] = acc;                        // dP_0j for now
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_early);   // done with V and dO
            mbar_wait(bar_qkp, ph);
            float dsum = 0.f;
#pragma unroll
            for (int jj = 0; jj < 5; ++jj) {
                const int j = jj * 32 + lane;
                pv[jj] = 0.f;
                if (j < AT_L)
                    pv[jj] = j < 128 ? lds_bf16(sP + (j >> 6) * BT_TILE + sw128(128, (j & 63) >> 3) + (j & 7) * 2) : p0_128;
                dsum += ds[jj] * pv[jj];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
            mbar_wait(bar_dv, ph);                   // line 128 of P is no longer read by the dV MMA: overwrite it with dS
#pragma unroll
            for (int jj = 0; jj < 5; ++jj) {
                const int j = jj * 32 + lane;
                const __nv_bfloat16 d16 = __float2bfloat16(pv[jj] * (ds[jj] - dsum) * p.scale);
                ds[jj] = j < AT_L ? __bfloat162float(d16) : 0.f;
                if (j < 128)
                    *reinterpret_cast<__nv_bfloat16*>(sP + (j >> 6) * BT_TILE + sw128(128, (j & 63) >> 3) + (j & 7) * 2) = d16;
                else if (j == 128)
                    dscol[128] = ds[jj];
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_ds);
            float o0 = 0.f, o1 = 0.f;
#pragma unroll
            for (int jj = 0; jj < 5; ++jj) {
                const int jn = (jj < 4) ? 32 : 1;
                _Pragma("unroll 8") for (int t = 0; t < jn; ++t) {
                    const int j = jj * 32 + t;
                    const float dj = __shfl_sync(0xffffffffu, ds[jj], t);
                    const float2 f = lds_bf162(sK + sw128(j, lane >> 2) + (lane & 3) * 4);
                    o0 += dj * f.x;
                    o1 += dj * f.y;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_late);    // done with K
            *reinterpret_cast<__nv_bfloat162*>(p.d_qkv + (size_t)row0 * p.ld_qkv + h * AT_HD + 2 * lane) =
                __floats2bfloat162_rn(o0, o1);
        }
    } else if (warp == 6) {
        // ---------------- key 128 on CUDA cores: dV[128] = sum_i P[i][128] dO[i],  dK[128] = sum_i dS[i][128] Q[i]
        uint32_t ph = 0;
        for (int blk = blockIdx.x; blk < n_items; blk += stride, ph ^= 1) {
            const int s = blk / p.H, h = blk - s * p.H, row0 = s * AT_L;
            const __nv_bfloat16* Pg = p.P + (size_t)blk * AT_L * AT_PLD;
            float col[5];
#pragma unroll
            for (int ii = 0; ii < 5; ++ii) {
                const int i = ii * 32 + lane;        // query line i <-> token (i < 128 ? i + 1 : 0)
                col[ii] = i < AT_L ? __bfloat162float(Pg[(size_t)(i < 128 ? i + 1 : 0) * AT_PLD + 128]) : 0.f;
            }
            mbar_wait(bar_vdo, ph);
            float o0 = 0.f, o1 = 0.f;
#pragma unroll
            for (int ii = 0; ii < 5; ++ii) {
                const int in = (ii < 4) ? 32 : 1;
                _Pragma("unroll 8") for (int t = 0; t < in; ++t) {
                    const int i = ii * 32 + t;
                    const float pj = __shfl_sync(0xffffffffu, col[ii], t);
                    const float2 f = lds_bf162(sdO + sw128(i, lane >> 2) + (lane & 3) * 4);
                    o0 += pj * f.x;
                    o1 += pj * f.y;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_early);   // done with dO
            __nv_bfloat16* krow = p.d_qkv + (size_t)(row0 + 128) * p.ld_qkv + h * AT_HD;
            *reinterpret_cast<__nv_bfloat162*>(krow + 2 * HC + 2 * lane) = __floats2bfloat162_rn(o0, o1);
            mbar_wait(bar_qkp, ph);
            mbar_wait(bar_ds, ph);
#pragma unroll
            for (int ii = 0; ii < 5; ++ii) {
                const int i = ii * 32 + lane;
                col[ii] = i < AT_L ? dscol[i] : 0.f;
            }
            o0 = o1 = 0.f;
#pragma unroll
            for (int ii = 0; ii < 5; ++ii) {
                const int in = (ii < 4) ? 32 : 1;
                _Pragma("unroll 8") for (int t = 0; t < in; ++t) {
                    const int i = ii * 32 + t;
                    const float dj = __shfl_sync(0xffffffffu, col[ii], t);
                    const float2 f = lds_bf162(sQ + sw128(i, lane >> 2) + (lane & 3) * 4);
                    o0 += dj * f.x;
                    o1 += dj * f.y;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_late);    // done with Q and the dS column
            *reinterpret_cast<__nv_bfloat162*>(krow + HC + 2 * lane) = __floats2bfloat162_rn(o0, o1);
        }
    } else {
        // ---------------- thread i: query line i (token i+1) for dS / dQ, key i (token i) for dK / dV
        const int i = threadIdx.x;
        const uint32_t tB = tmem + (static_cast<uint32_t>(warp * 32) << 16);
        uint32_t ph = 0;
        for (int blk = blockIdx.x; blk < n_items; blk += stride, ph ^= 1) {
            const int s = blk / p.H, h = blk - s * p.H, row0 = s * AT_L;
            const __nv_bfloat16* Pg = p.P + (size_t)blk * AT_L * AT_PLD;
            const float p128 = __bfloat162float(Pg[(size_t)(i + 1) * AT_PLD + 128]);     // P[token i+1][key 128]
            mbar_wait(bar_vdo, ph);
            const float dp128 = dot_rows64(sdO, i, sV, 128);                              // dP[i][128] = dO_i . V_128
            mbar_arrive(bar_early);                  // done with V and dO
            mbar_wait(bar_qkp, ph);
            mbar_wait(bar_dp, ph);
            tc_fence_after();
            float delta = p128 * dp128;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {            // 32 score columns per pass
                uint32_t r[32];
                tmem_ld_32x32(tB + BT_TM_DP + c * 32, r);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint4 u = *reinterpret_cast<const uint4*>(sP + (c >> 1) * BT_TILE + sw128(i, (c & 1) * 4 + q));
                    const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const float2 f = __bfloat1622float2(hh[t]);
                        delta += f.x * __uint_as_float(r[q * 8 + 2 * t]) + f.y * __uint_as_float(r[q * 8 + 2 * t + 1]);
                    }
                }
            }
            const float ds128 = __bfloat162float(__float2bfloat16(p128 * (dp128 - delta) * p.scale));
            dscol[i] = ds128;
            mbar_wait(bar_dv, ph);                   // P is no longer read by the dV MMA: overwrite it with dS
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t r[32];
                tmem_ld_32x32(tB + BT_TM_DP + c * 32, r);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint4* pa = reinterpret_cast<uint4*>(sP + (c >> 1) * BT_TILE + sw128(i, (c & 1) * 4 + q));
                    uint4 u = *pa;
                    __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const float2 f = __bfloat1622float2(hh[t]);
                        hh[t] = __floats2bfloat162_rn(f.x * (__uint_as_float(r[q * 8 + 2 * t]) - delta) * p.scale,
                                                      f.y * (__uint_as_float(r[q * 8 + 2 * t + 1]) - delta) * p.scale);
                    }
                    *pa = u;
                }
            }
            tc_fence_before();
            fence_proxy_async_smem();
            mbar_arrive(bar_ds);
            __nv_bfloat16* krow = p.d_qkv + (size_t)(row0 + i) * p.ld_qkv + h * AT_HD;
            store_row64_bf16(krow + 2 * HC, tB + BT_TM_DV);                                           // dV[key i]
            mbar_wait(bar_dq, ph);
            tc_fence_after();
            store_row64_bf16(p.d_qkv + (size_t)(row0 + 1 + i) * p.ld_qkv + h * AT_HD, tB + BT_TM_DQ, ds128, sK, 128);   // dQ[token i+1]
            mbar_arrive(bar_late);                   // done with K row 128
            mbar_wait(bar_dk, ph);
            tc_fence_after();
            store_row64_bf16(krow + HC, tB + BT_TM_DK);                                               // dK[key i]
            tc_fence_before();
            mbar_arrive(bar_tmfree);                 // this thread's lanes of dV / dQ / dK have left TMEM
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc<256>(tmem);
}

int attention_tc_bwd(const EdbAttnDesc& d, cudaStream_t st) {
    if (d.nseq <= 0) return EDB_OK;
    if (d.out == nullptr || (d.ld_out % 8) != 0 || (reinterpret_cast<uintptr_t>(d.out) & 15) != 0)
        return edb_set_error(EDB_ERR_SHAPE, "attention backward (tensor-core path): `out` = the forward output is required "
                                            "(16-byte aligned rows); delta_i = dO_i . O_i is computed from it");
    const long long R = (long long)d.nseq * AT_L;
    const long long W3 = 3LL * d.heads * AT_HD, W1 = (long long)d.heads * AT_HD;
    CUtensorMap mq128, mq1, mkv, mdo128, mdo1, mp128, mp1;
    EDB_TRY(make_tmap_bf16(&mq128, d.qkv, W3, R, d.ld_qkv, 128));
    EDB_TRY(make_tmap_bf16(&mq1, d.qkv, W3, R, d.ld_qkv, 1));
    EDB_TRY(make_tmap_bf16(&mkv, d.qkv, W3, R, d.ld_qkv, AT_KP));
    EDB_TRY(make_tmap_bf16(&mdo128, d.d_out, W1, R, d.ld_dout, 128));
    EDB_TRY(make_tmap_bf16(&mdo1, d.d_out, W1, R, d.ld_dout, 1));
    EDB_TRY(make_tmap_bf16(&mp128, d.P, AT_PLD, (long long)d.nseq * d.heads * AT_L, AT_PLD, 128));
    EDB_TRY(make_tmap_bf16(&mp1, d.P, AT_PLD, (long long)d.nseq * d.heads * AT_L, AT_PLD, 1));
    AttnTcBwdParams p{};
    p.qkv = (const __nv_bfloat16*)d.qkv; p.ld_qkv = d.ld_qkv; p.d_qkv = (__nv_bfloat16*)d.d_qkv;
    p.P = (const __nv_bfloat16*)d.P;
    p.H = d.heads; p.scale = d.scale;
    p.idesc_dp = make_idesc_bf16(128, 128, 0, 0);
    p.idesc_kk_mn = make_idesc_bf16(128, AT_HD, 1, 1);
    p.idesc_k_mn = make_idesc_bf16(128, AT_HD, 0, 1);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(attn_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BT_TOTAL);
        if (e != cudaSuccess) return edb_set_error(EDB_ERR_CUDA, cudaGetErrorString(e));
        configured = true;
    }
    const int n_items = d.nseq * d.heads;
    const int slots = 2 * num_sms();            // two resident CTAs per SM
    attn_tc_bwd_kernel<<<n_items < slots ? n_items : slots, 224, BT_TOTAL, st>>>(mq128, mq1, mkv, mdo128, mdo1, mp128, mp1, p,
                                                                                n_items);
    EDB_CHECK_LAUNCH();
    return EDB_OK;
}

}  // namespace edb

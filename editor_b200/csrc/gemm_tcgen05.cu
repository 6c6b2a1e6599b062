// Persistent, warp-specialised bf16 GEMM for sm_100a:  D[M,N] = epilogue( sum_k A(m,k) * B(n,k) ).
//
//   warp 0      : TMA producer  (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier full/empty)
//   warp 1      : MMA issuer    (one elected lane issues tcgen05.mma, 128 x BN x 16 per instruction, fp32 in TMEM)
//   warp 2      : TMEM allocator (2 accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1)
//   warps 4..   : epilogue      (8 or 16 warps: tcgen05.ld -> smem transpose -> bias / GELU / residual / GELU' / split-K
//                                reduce -> coalesced global; 16 warps for the arithmetic-heavy and store-heavy variants)
//
// CG = 2 (the default for N > 128, M > 128): the two CTAs of a TPC form a cluster and run ONE tcgen05.mma.cta_group::2 of
// 256 x 256 x 16 per instruction.  Each CTA stages its own 128 rows of A and HALF of the B tile (32 KB per k-block instead
// of 48 KB), which takes the operand traffic L2 -> SM from 11.7 to 7.8 bytes per kFLOP: the 128 x 256 single-CTA tile is
// bound by the ~6300 B/clk L2 fabric (B300_MICROARCH.md "LTS throughput cap"), not by the tensor pipe.  Both producers
// signal the leader's "full" barrier, the leader's commits are multicast to both CTAs' "empty" / "accumulator full"
// barriers, and every epilogue warp of the pair arrives on the leader's "accumulator empty" barrier.
//
// Both operands may be K-major (reduction dim contiguous) or MN-major (reduction dim strided); that covers the
// forward (X W^T), dgrad (dY W) and wgrad (dY^T X) products of every Linear on EDITOR's hot path
// (reference: modeling/backbones/vit_pytorch.py:139-145,184-198,158-168,240-258) without any transposed copies.
#include "ptx.cuh"
#include "abi_internal.h"

#ifndef EDB_AUX_PREFETCH
#define EDB_AUX_PREFETCH 3      // next tile's aux slab -> L2: 0 = off, 1 = cp.async.bulk.prefetch.L2 per row, 2 = prefetch.global.L2
//                                 lines, 3 = one cp.async.bulk.prefetch.tensor.2d per warp (32 x 64 box of a tensor map over aux)
#endif
// Experimental build switches for A/B runs on one box (make OUT=../lib_x EXTRA=-D...; EDB_LIB=.../lib_x/libeditor_b200.so
// python tools/gemm_bench.py); both are OFF in the shipped library (measured together: fc1 forward 0.233 -> 0.230 ms, ~1 %):
//   EDB_BIAS_PRELOAD=1  the bias of chunk c+1 is fetched during chunk c (today every chunk waits for its own bias load:
//                       long-scoreboard stalls at the first FFMA of phase B in profiles/r01_ncu_gemm_fc1.txt)
//   EDB_WAIT_BACKOFF=n  producer / epilogue waits poll with n ns of nanosleep in between (the spin loops are ~10 % of the
//                       issued instructions of the fc1 GEMM)
#ifndef EDB_BIAS_PRELOAD
#define EDB_BIAS_PRELOAD 0
#endif
#ifndef EDB_WAIT_BACKOFF
#define EDB_WAIT_BACKOFF 0
#endif
#ifndef EDB_GELU_EPI_WARPS
#define EDB_GELU_EPI_WARPS 16
#endif

namespace edb {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
// epilogue warps per CTA: 8 (32-column chunks) or 16 (16-column chunks; twice the warps per scheduler to hide the FMA /
// shared-memory latencies of the arithmetic-heavy GELU epilogues)
template <int NEPI> constexpr int gemm_threads() { return 128 + NEPI * 32; }

struct GemmKernelParams {
    int M, N, K;
    int m_tiles, n_tiles, split_k, kb_per_split, kb_total;
    int a_mn, b_mn;
    uint32_t idesc;
    void* D;
    long long ldd;
    int out_f32;
    int epi;
    const float* bias;
    const void* aux;
    long long ld_aux;
    int aux_f32;
    void* out2;
    long long ld_out2;
    float alpha;
    const float* row_scale;
    int scale_group;
    const int* M_dev;   // optional: number of valid rows read on the device (packed HMA rows; no host sync)
    const int* K_dev;   // optional: reduction length read on the device (wgrad over packed rows)
    float* colsum;      // optional (EPI_STORE / EPI_GELU_BWD): column sums of D accumulated with atomics (bias gradient)
    int aux_tmap_ok;    // tmap_aux describes the bf16 aux matrix (EPI_GELU_BWD): slab prefetches go through it
};

__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

// bf16-output epilogues.  GELU and GELU' both come from ONE tanh.approx per element:
//     Phi(x) = (1 + tanh(x P(x^2))) / 2,   P(t) = a + b t + c t^2 fitted to atanh(2 Phi(x) - 1) / x on |x| <= 4
//     gelu(x)  = x/2 + x/2 tanh(u)                                     (max abs error 3.0e-5 + tanh.approx's 2^-11 relative)
//     gelu'(x) = (1 + tanh u)/2 + x/2 (1 - tanh^2 u) (a + 3b t + 5c t^2)   (the exact derivative of the line above; 1.2e-4)
// t is clamped at 36 (tanh is +-1 beyond |x| = 6, and P stays positive).  The forward fc1 epilogue writes BOTH gelu(pre)
// (operand of fc2) and gelu'(pre) (bf16): the backward epilogue (EPI_GELU_BWD) is then a single multiply and the fc2
// dgrad GEMM is paced by its MMAs, not by 12 extra FMAs per element; nothing else on the path needs `pre` itself.
// (~13 issue slots per element for both outputs, against 14 for the previous degree-17 polynomial GELU alone.)  The
// fp32-output path (EDB_PREC_FP32) keeps exact erff and stores `pre`.
__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;\n" : "=f"(y) : "f"(x));
    return y;
}
constexpr float kGeluA = 7.97458247e-01f, kGeluB = 3.70506044e-02f, kGeluC = -3.58791060e-04f;
__device__ __forceinline__ float gelu_fast(float x) {
    const float t = fminf(x * x, 36.0f);
    const float th = tanh_approx(x * fmaf(fmaf(kGeluC, t, kGeluB), t, kGeluA));
    const float hx = 0.5f * x;
    return fmaf(hx, th, hx);
}
__device__ __forceinline__ void gelu_and_grad_fast(float x, float& g, float& d) {
    const float t = fminf(x * x, 36.0f);
    const float th = tanh_approx(x * fmaf(fmaf(kGeluC, t, kGeluB), t, kGeluA));
    const float q = fmaf(fmaf(5.0f * kGeluC, t, 3.0f * kGeluB), t, kGeluA);
    const float hx = 0.5f * x;
    g = fmaf(hx, th, hx);
    d = fmaf(hx * fmaf(-th, th, 1.0f), q, fmaf(0.5f, th, 0.5f));
}

// Packed fp32 arithmetic (sm_100: fma/mul/add.rn.f32x2 -> FFMA2, one issue slot for two fp32 operations, full fp32
// precision).  The GELU epilogues are ISSUE-bound (profiles/r01_ncu_gemm_fc1.txt: 123 M instructions, 57 % issue-active,
// tensor pipe 52 %): the polynomial + derivative arithmetic of two neighbouring columns goes through one instruction.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float a, float b) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// same arithmetic, operation for operation, as gelu_fast / gelu_and_grad_fast on each half (fma.rn.f32x2 rounds each lane
// like fma.rn.f32): the packed and the scalar epilogues produce bit-identical results
__device__ __forceinline__ void gelu_pair_fast(float& x0, float& x1) {
    const f32x2 x = pack2(x0, x1);
    float t0, t1;
    unpack2(mul2(x, x), t0, t1);
    const f32x2 t = pack2(fminf(t0, 36.0f), fminf(t1, 36.0f));
    const f32x2 cB = pack2(kGeluB, kGeluB), cA = pack2(kGeluA, kGeluA), cC = pack2(kGeluC, kGeluC);
    float u0, u1;
    unpack2(mul2(x, fma2(fma2(cC, t, cB), t, cA)), u0, u1);
    const f32x2 th = pack2(tanh_approx(u0), tanh_approx(u1));
    const f32x2 hx = mul2(pack2(0.5f, 0.5f), x);
    unpack2(fma2(hx, th, hx), x0, x1);
}
__device__ __forceinline__ void gelu_and_grad_pair_fast(float& x0, float& x1, float& d0, float& d1) {
    const f32x2 x = pack2(x0, x1);
    float t0, t1;
    unpack2(mul2(x, x), t0, t1);
    const f32x2 t = pack2(fminf(t0, 36.0f), fminf(t1, 36.0f));
    const f32x2 cB = pack2(kGeluB, kGeluB), cA = pack2(kGeluA, kGeluA), cC = pack2(kGeluC, kGeluC);
    float u0, u1;
    unpack2(mul2(x, fma2(fma2(cC, t, cB), t, cA)), u0, u1);
    const float th0 = tanh_approx(u0), th1 = tanh_approx(u1);
    const f32x2 th = pack2(th0, th1), nth = pack2(-th0, -th1);
    const f32x2 q = fma2(fma2(pack2(5.0f * kGeluC, 5.0f * kGeluC), t, pack2(3.0f * kGeluB, 3.0f * kGeluB)), t, cA);
    const f32x2 half = pack2(0.5f, 0.5f), one = pack2(1.0f, 1.0f);
    const f32x2 hx = mul2(half, x);
    unpack2(fma2(hx, th, hx), x0, x1);
    unpack2(fma2(mul2(hx, fma2(nth, th, one)), q, fma2(half, th, half)), d0, d1);
}
#ifndef EDB_GELU_PACKED
#define EDB_GELU_PACKED 1
#endif
__device__ __forceinline__ float4 gelu4_fast(float4 v) {
#if EDB_GELU_PACKED
    gelu_pair_fast(v.x, v.y);
    gelu_pair_fast(v.z, v.w);
    return v;
#else
    return make_float4(gelu_fast(v.x), gelu_fast(v.y), gelu_fast(v.z), gelu_fast(v.w));
#endif
}
// v <- gelu(v), returns gelu'(v)
__device__ __forceinline__ float4 gelu4_and_grad_fast(float4& v) {
    float4 d;
#if EDB_GELU_PACKED
    gelu_and_grad_pair_fast(v.x, v.y, d.x, d.y);
    gelu_and_grad_pair_fast(v.z, v.w, d.z, d.w);
#else
    gelu_and_grad_fast(v.x, v.x, d.x);
    gelu_and_grad_fast(v.y, v.y, d.y);
    gelu_and_grad_fast(v.z, v.z, d.z);
    gelu_and_grad_fast(v.w, v.w, d.w);
#endif
    return d;
}
// v * (saved bf16 gelu' factor)
__device__ __forceinline__ float4 mul_bf16x4(float4 v, uint2 f) {
    const float2 a0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&f.x));
    const float2 a1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&f.y));
    return make_float4(v.x * a0.x, v.y * a0.y, v.z * a1.x, v.w * a1.y);
}
// kept out of line: inlining the erf arithmetic into the unrolled epilogue would overflow the instruction cache
__device__ __noinline__ float4 gelu4_exact(float4 v) {
    return make_float4(gelu_exact(v.x), gelu_exact(v.y), gelu_exact(v.z), gelu_exact(v.w));
}

template <int BN, int STAGES, int CG>
struct GemmSmem {
    static constexpr int kABytes = BM * BK * 2;
    static constexpr int kBBytes = (BN / CG) * BK * 2;     // a CTA of a pair stages half of the B tile
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStagingOffset = STAGES * kStageBytes;           // 8 epilogue warps x 4 KB transpose patches
    static constexpr int kBarOffset = kStagingOffset + 8 * 4096;
    static constexpr int kTotal = kBarOffset + 256 + 1024;  // + alignment slack
};

// 4 consecutive elements with a column predicate (nvalid = number of in-range columns, may exceed 4)
__device__ __forceinline__ float4 ld4g(const float* p, int nvalid, bool vec) {
    if (vec) return *reinterpret_cast<const float4*>(p);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    v.x = p[0];
    if (nvalid > 1) v.y = p[1];
    if (nvalid > 2) v.z = p[2];
    if (nvalid > 3) v.w = p[3];
    return v;
}
__device__ __forceinline__ float4 ld4g(const __nv_bfloat16* p, int nvalid, bool vec) {
    if (vec) {
        const uint2 u = *reinterpret_cast<const uint2*>(p);
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
        const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
        return make_float4(a.x, a.y, b.x, b.y);
    }
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    v.x = __bfloat162float(p[0]);
    if (nvalid > 1) v.y = __bfloat162float(p[1]);
    if (nvalid > 2) v.z = __bfloat162float(p[2]);
    if (nvalid > 3) v.w = __bfloat162float(p[3]);
    return v;
}
__device__ __forceinline__ void st4g(float* p, const float4& v, int nvalid, bool vec) {
    if (vec) { *reinterpret_cast<float4*>(p) = v; return; }
    p[0] = v.x;
    if (nvalid > 1) p[1] = v.y;
    if (nvalid > 2) p[2] = v.z;
    if (nvalid > 3) p[3] = v.w;
}
__device__ __forceinline__ void st4g(__nv_bfloat16* p, const float4& v, int nvalid, bool vec) {
    if (vec) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        uint2 u;
        u.x = *reinterpret_cast<uint32_t*>(&lo);
        u.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(p) = u;
        return;
    }
    p[0] = __float2bfloat16(v.x);
    if (nvalid > 1) p[1] = __float2bfloat16(v.y);
    if (nvalid > 2) p[2] = __float2bfloat16(v.z);
    if (nvalid > 3) p[3] = __float2bfloat16(v.w);
}

template <int BN, int STAGES, int EPI, int NEPI, int CG>
__global__ void __launch_bounds__(gemm_threads<NEPI>(), 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_aux, const GemmKernelParams p) {
    using S = GemmSmem<BN, STAGES, CG>;
    constexpr int BNC = BN / CG;            // rows of B this CTA stages
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;     // 0 = leader (issues the MMAs)
    const int unit = blockIdx.x / CG;       // persistent work index of this CTA (pair)
    const int units = gridDim.x / CG;
    extern __shared__ __align__(1024) uint8_t smem_raw[];   // 128B-swizzled tiles need a 1024-byte aligned base
    uint8_t* smem = smem_raw;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], NEPI * CG);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        if (CG == 2) tmem_alloc_pair<2 * BN>(tmem_ptr);
        else tmem_alloc<2 * BN>(tmem_ptr);
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();        // the peer's barriers are initialised before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    const int M_rt = p.M_dev != nullptr ? *p.M_dev : p.M;
    const int m_tiles = (M_rt + BM * CG - 1) / (BM * CG);     // tiles of the CTA (pair): 128 * CG rows
    int kb_total = p.kb_total, kb_per_split = p.kb_per_split;
    if (p.K_dev != nullptr) {
        kb_total = (*p.K_dev + BK - 1) / BK;
        kb_per_split = (kb_total + p.split_k - 1) / p.split_k;
    }
    const int num_work = m_tiles * p.n_tiles * p.split_k;
    constexpr int GM = 16 / CG;  // m-tiles per raster group (keeps the group's A tiles + all of B resident in L2)

    auto decode = [&](int w, int& tm, int& tn, int& ks) {
        ks = w % p.split_k;
        int t = w / p.split_k;
        const int per_group = GM * p.n_tiles;
        const int g = t / per_group;
        const int within = t - g * per_group;
        const int gsize = min(GM, m_tiles - g * GM);
        tm = g * GM + within % gsize;
        tn = within / gsize;
    };

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int w = unit; w < num_work; w += units) {
                int tm, tn, ks;
                decode(w, tm, tn, ks);
                const int row_a = (tm * CG + (int)rank) * BM;           // this CTA's rows of A
                const int row_b = tn * BN + (int)rank * BNC;            // this CTA's part of the B tile
                const int kb0 = ks * kb_per_split;
                const int kb1 = min(kb_total, kb0 + kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb) {
#if EDB_WAIT_BACKOFF > 0
                    mbar_wait_backoff(&empty_bar[stage], phase ^ 1, EDB_WAIT_BACKOFF);
#else
                    mbar_wait(&empty_bar[stage], phase ^ 1);
#endif
                    uint8_t* sa = smem + stage * S::kStageBytes;
                    uint8_t* sb = sa + S::kABytes;
                    // the leader's barrier collects the bytes of both CTAs
                    if (rank == 0) mbar_expect_tx(&full_bar[stage], S::kStageBytes * CG);
                    const uint32_t fb = (CG == 2) ? mapa_shared(smem_u32(&full_bar[stage]), 0u) : 0u;
                    auto load = [&](void* dst, const CUtensorMap* m, int c0, int c1) {
                        if (CG == 2) tma_load_2d_pair(dst, m, fb, c0, c1);
                        else tma_load_2d(dst, m, &full_bar[stage], c0, c1);
                    };
                    if (!p.a_mn) {
                        load(sa, &tmap_a, kb * BK, row_a);
                    } else {
#pragma unroll
                        for (int j = 0; j < BM / 64; ++j) load(sa + j * (BK * 128), &tmap_a, row_a + j * 64, kb * BK);
                    }
                    if (!p.b_mn) {
                        load(sb, &tmap_b, kb * BK, row_b);
                    } else {
#pragma unroll
                        for (int j = 0; j < BNC / 64; ++j) load(sb + j * (BK * 128), &tmap_b, row_b + j * 64, kb * BK);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (lane == 0 && rank == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            const uint32_t a_lbo = p.a_mn ? BK * 128 : 0, b_lbo = p.b_mn ? BK * 128 : 0;
            const uint32_t a_kstep = p.a_mn ? 2048 : 32, b_kstep = p.b_mn ? 2048 : 32;
            for (int w = unit; w < num_work; w += units) {
                int tm, tn, ks;
                decode(w, tm, tn, ks);
                const int kb0 = ks * kb_per_split;
                const int kb1 = min(kb_total, kb0 + kb_per_split);
                if (kb0 >= kb1) continue;      // empty split (device-side K shorter than the host bound)
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * S::kStageBytes);
                    const uint32_t sb = sa + S::kABytes;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t da = make_smem_desc(sa + k * a_kstep, a_lbo, 1024);
                        const uint64_t db = make_smem_desc(sb + k * b_kstep, b_lbo, 1024);
                        if (CG == 2) tc_mma_bf16_pair(d_tmem, da, db, p.idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                        else tc_mma_bf16(d_tmem, da, db, p.idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                    // frees the smem slot (of both CTAs) once these MMAs retire
                    if (CG == 2) tc_commit_pair(&empty_bar[stage]);
                    else tc_commit(&empty_bar[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                // accumulator complete -> epilogue (of both CTAs)
                if (CG == 2) tc_commit_pair(&tmem_full[acc]);
                else tc_commit(&tmem_full[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        // Each warp drains a 32-row x (BN/2)-column slab of the accumulator in 32-column chunks.  tcgen05.ld hands every
        // lane one ROW; global memory wants lanes on consecutive COLUMNS.  So the chunk is transposed through a private
        // 4 KB swizzled smem patch: phase A (row per lane) TMEM -> smem, phase B (8 lanes per row, 4 rows per access)
        // smem -> bias / GELU / residual / GELU' -> fully coalesced 128-byte global accesses.  The aux operand of chunk
        // c+1 (residual stream / saved pre-activation) is fetched while chunk c is processed.
        const int ew = warp - 4;
        const int quarter = warp & 3;          // TMEM lane quarter this warp may read
        const int part = ew >> 2;              // column part (NEPI/4 parts per tile)
        constexpr int NPART = NEPI / 4;
        constexpr int CW = 256 / NEPI;         // chunk width in columns: 32 (8 warps) or 16 (16 warps)
        constexpr int NCH = BN / NPART / CW;   // chunks per warp
        constexpr int LPR = CW / 4;            // phase B: lanes per row (each lane 4 columns)
        constexpr int RPA = 32 / LPR;          // phase B: rows per access
        constexpr int NIT = 32 / RPA;          // phase B: accesses per chunk
        constexpr int RB = CW * 4;             // bytes per patch row
        constexpr bool kAuxF32 = (EPI == EPI_RESIDUAL);
        constexpr bool kAuxBf16 = (EPI == EPI_GELU_BWD);
        constexpr bool kColsum = (EPI == EPI_GELU_BWD || EPI == EPI_STORE);   // fused bias gradient (optional)
        uint8_t* stg = smem + S::kStagingOffset + ew * (32 * RB);
        const int lrow = lane / LPR;           // phase B: row within a group of RPA
        const int lc4 = lane % LPR;            // phase B: which float4 of the chunk
        // conflict-free 16-byte slot of (row, logical slot) inside the patch
        auto slot = [](int row, int j) { return CW == 32 ? (j ^ (row & 7)) : (j ^ ((row >> 1) & 3)); };
        int acc = 0;
        uint32_t acc_phase = 0;
        float4 axf_c[kAuxF32 ? NIT : 1], axf_n[kAuxF32 ? NIT : 1];     // fp32 aux: current / next chunk
        uint2 axh_c[kAuxBf16 ? NIT : 1], axh_n[kAuxBf16 ? NIT : 1];     // bf16 aux
        float rsc[kAuxF32 ? NIT : 1];                                   // DropPath row scales of this lane's rows
        // INTERIOR tiles (all 32 rows and all columns of the warp's slab in range, 16-byte aligned pitches -- every tile
        // of the backbone GEMMs except the last row tile) take a path without per-element predicates and with
        // incremental row pointers; the generic path below handles ragged edges, odd pitches and the fp32 GELU.
        const bool pitch_ok = (p.ldd % 4 == 0) && (p.out2 == nullptr || p.ld_out2 % 4 == 0) &&
                              (p.aux == nullptr || p.ld_aux % 4 == 0) && !(EPI == EPI_GELU && p.out_f32);
        const size_t d_step = (size_t)RPA * p.ldd, o2_step = (size_t)RPA * p.ld_out2, aux_step = (size_t)RPA * p.ld_aux;
        for (int w = unit; w < num_work; w += units) {
            int tm, tn, ks;
            decode(w, tm, tn, ks);
            if (ks * kb_per_split >= kb_total) continue;      // empty split: the issuer skipped it too
            const int row_base = (tm * CG + (int)rank) * BM + quarter * 32;
            const uint32_t t_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
            const int colw = tn * BN + part * (BN / NPART) + lc4 * 4;
            const bool fast = pitch_ok && (row_base + 32 <= M_rt) && (tn * BN + BN <= p.N);
            const size_t row0 = (size_t)(row_base + lrow);     // this lane's first row in phase B
            auto load_aux = [&](int c) {
                const int col = colw + c * CW;
                if (fast) {
                    if (kAuxF32) {
                        const float* ap = reinterpret_cast<const float*>(p.aux) + row0 * p.ld_aux + col;
#pragma unroll
                        for (int it = 0; it < NIT; ++it) axf_n[it] = *reinterpret_cast<const float4*>(ap + it * aux_step);
                    }
                    if (kAuxBf16) {
                        const __nv_bfloat16* ap = reinterpret_cast<const __nv_bfloat16*>(p.aux) + row0 * p.ld_aux + col;
#pragma unroll
                        for (int it = 0; it < NIT; ++it) axh_n[it] = *reinterpret_cast<const uint2*>(ap + it * aux_step);
                    }
                    return;
                }
                const int nvalid = p.N - col;
                if (kAuxF32) {
#pragma unroll
                    for (int it = 0; it < NIT; ++it) {
                        const int row = row_base + it * RPA + lrow;
                        axf_n[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (row < M_rt && nvalid > 0)
                            axf_n[it] = ld4g(reinterpret_cast<const float*>(p.aux) + (size_t)row * p.ld_aux + col, nvalid,
                                             nvalid >= 4 && (p.ld_aux % 4 == 0));
                    }
                }
                if (kAuxBf16) {
#pragma unroll
                    for (int it = 0; it < NIT; ++it) {
                        const int row = row_base + it * RPA + lrow;
                        axh_n[it] = make_uint2(0u, 0u);
                        if (row < M_rt && nvalid > 0) {
                            const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(p.aux) + (size_t)row * p.ld_aux + col;
                            if (nvalid >= 4 && (p.ld_aux % 4 == 0)) {
                                axh_n[it] = *reinterpret_cast<const uint2*>(src);
                            } else {
                                unsigned short e[4] = {0, 0, 0, 0};
                                for (int q = 0; q < 4; ++q)
                                    if (q < nvalid) e[q] = *reinterpret_cast<const unsigned short*>(src + q);
                                axh_n[it] = make_uint2((uint32_t)e[0] | ((uint32_t)e[1] << 16), (uint32_t)e[2] | ((uint32_t)e[3] << 16));
                            }
                        }
                    }
                }
            };
            load_aux(0);
            if (kAuxBf16 && EDB_AUX_PREFETCH != 0) {
                // the bf16 aux slab (saved gelu') of this warp in the NEXT tile of this CTA goes to L2 now (one row per
                // lane): by the time its chunks are fetched into registers (one chunk ahead) they come from L2, not from
                // HBM -- the register prefetch alone left every chunk waiting on a DRAM round trip.  A/B on one box
                // (tools/gemm_bench.py): fc2 dgrad 0.255 -> 0.230 ms.  NOT done for the fp32 residual epilogues: there the
                // same prefetch (512 B per row) cost 8 % (proj 0.087 -> 0.095 ms, fc2 fwd 0.165 -> 0.178 ms).
                const int wn = w + units;
                if (wn < num_work) {
                    int tm2, tn2, ks2;
                    decode(wn, tm2, tn2, ks2);
                    const int c2 = tn2 * BN + part * (BN / NPART);
                    constexpr int kSlabCols = BN / NPART;
#if EDB_AUX_PREFETCH == 3
                    // ONE tensor-map prefetch per warp for its whole 32-row x 64-column slab (the per-lane bulk prefetch
                    // below compiles to a 32-iteration uniform-register loop: 14 % of this kernel's issued instructions,
                    // profiles/r01 fc2-dgrad capture); rows / columns past the matrix are clipped by the tensor map
                    if (kSlabCols == 64 && p.aux_tmap_ok && lane == 0)
                        tma_prefetch_2d(&tmap_aux, c2, (tm2 * CG + (int)rank) * BM + quarter * 32);
#else
                    const int r2 = (tm2 * CG + (int)rank) * BM + quarter * 32 + lane;
                    const int esz = kAuxF32 ? 4 : 2;
                    if (pitch_ok && r2 < M_rt && c2 + kSlabCols <= p.N && (p.ld_aux * esz) % 16 == 0) {
                        const uint8_t* src = reinterpret_cast<const uint8_t*>(p.aux) + ((size_t)r2 * p.ld_aux + c2) * esz;
#if EDB_AUX_PREFETCH == 1
                        prefetch_l2_bulk(src, kSlabCols * esz);
#elif EDB_AUX_PREFETCH == 2
#pragma unroll
                        for (int b = 0; b < kSlabCols * esz; b += 128) prefetch_l2_line(src + b);
#endif
                    }
#endif
                }
            }
            if (kAuxF32) {          // DropPath scale of each of this lane's rows: once per tile, not once per chunk
#pragma unroll
                for (int it = 0; it < NIT; ++it) {
                    const int row = row_base + it * RPA + lrow;
                    rsc[it] = (p.row_scale != nullptr && row < M_rt) ? p.row_scale[row / p.scale_group] : 1.0f;
                }
            }
#if EDB_BIAS_PRELOAD
            float4 b4n = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias != nullptr && EPI != EPI_ATOMIC && p.N - colw > 0) b4n = ld4g(p.bias + colw, p.N - colw, p.N - colw >= 4);
#endif
#if EDB_WAIT_BACKOFF > 0
            mbar_wait_backoff(&tmem_full[acc], acc_phase, EDB_WAIT_BACKOFF);
#else
            mbar_wait(&tmem_full[acc], acc_phase);
#endif
            tc_fence_after();
            uint32_t r[CW];
#pragma unroll 1
            for (int c = 0; c < NCH; ++c) {
                const int col_t = part * (BN / NPART) + c * CW;
                const int col = colw + c * CW;                      // first of this lane's 4 columns in phase B
                const int nvalid = p.N - col;                        // >= 4: all four columns exist
                const bool vec = nvalid >= 4;
                if (kAuxF32) {
#pragma unroll
                    for (int it = 0; it < NIT; ++it) axf_c[it] = axf_n[it];
                }
                if (kAuxBf16) {
#pragma unroll
                    for (int it = 0; it < NIT; ++it) axh_c[it] = axh_n[it];
                }
                // ---- phase A (the TMEM load of chunk c was issued during phase B of chunk c-1)
                if (c == 0) {
                    if constexpr (CW == 32) tmem_ld_32x32(t_base + col_t, r);
                    else tmem_ld_32x16(t_base + col_t, r);
                }
                if (c + 1 < NCH) load_aux(c + 1);
#if EDB_BIAS_PRELOAD
                const float4 b4 = b4n;
                if (c + 1 < NCH) {
                    const int nv2 = nvalid - CW;
                    b4n = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (p.bias != nullptr && EPI != EPI_ATOMIC && nv2 > 0) b4n = ld4g(p.bias + col + CW, nv2, nv2 >= 4);
                }
#else
                float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.bias != nullptr && EPI != EPI_ATOMIC && nvalid > 0) b4 = ld4g(p.bias + col, nvalid, vec);
#endif
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < LPR; ++j)
                    *reinterpret_cast<uint4*>(stg + lane * RB + (slot(lane, j) << 4)) =
                        make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
                if (c + 1 < NCH) {          // next chunk's accumulator columns travel TMEM -> registers under phase B
                    if constexpr (CW == 32) tmem_ld_32x32(t_base + col_t + CW, r);
                    else tmem_ld_32x16(t_base + col_t + CW, r);
                }
                __syncwarp();
                // ---- phase B
                float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);     // EPI_GELU_BWD: column sums of this lane's rows
                if (fast) {
                    auto ldv = [&](int it) {
                        const int rr = it * RPA + lrow;
                        const uint4 raw = *reinterpret_cast<const uint4*>(stg + rr * RB + (slot(rr, lc4) << 4));
                        return make_float4(fmaf(__uint_as_float(raw.x), p.alpha, b4.x), fmaf(__uint_as_float(raw.y), p.alpha, b4.y),
                                           fmaf(__uint_as_float(raw.z), p.alpha, b4.z), fmaf(__uint_as_float(raw.w), p.alpha, b4.w));
                    };
                    auto st_bf16 = [](__nv_bfloat16* q, const float4& v) {
                        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
                        *reinterpret_cast<uint2*>(q) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
                    };
                    if constexpr (EPI == EPI_ATOMIC) {
                        float* dp = reinterpret_cast<float*>(p.D) + row0 * p.ldd + col;
#pragma unroll
                        for (int it = 0; it < NIT; ++it) {
                            const float4 v = ldv(it);
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(dp + it * d_step), "f"(v.x),
                                         "f"(v.y), "f"(v.z), "f"(v.w)
                                         : "memory");
                        }
                    } else if constexpr (EPI == EPI_RESIDUAL) {
                        float* dp = reinterpret_cast<float*>(p.D) + row0 * p.ldd + col;
#pragma unroll
                        for (int it = 0; it < NIT; ++it) {
                            float4 v = ldv(it);
                            const float4 a = axf_c[kAuxF32 ? it : 0];
                            const float sc = rsc[kAuxF32 ? it : 0];
                            v.x = fmaf(sc, v.x, a.x); v.y = fmaf(sc, v.y, a.y);
                            v.z = fmaf(sc, v.z, a.z); v.w = fmaf(sc, v.w, a.w);
                            *reinterpret_cast<float4*>(dp + it * d_step) = v;
                        }
                    } else if (p.out_f32) {          // EPI_STORE / EPI_GELU_BWD with fp32 output (fp32 GELU: generic path)
                        float* dp = reinterpret_cast<float*>(p.D) + row0 * p.ldd + col;
#pragma unroll
                        for (int it = 0; it < NIT; ++it) {
                            float4 v = ldv(it);
                            if (EPI == EPI_GELU_BWD) v = mul_bf16x4(v, axh_c[kAuxBf16 ? it : 0]);
                            if (kColsum) { cs.x += v.x; cs.y += v.y; cs.z += v.z; cs.w += v.w; }
                            *reinterpret_cast<float4*>(dp + it * d_step) = v;
                        }
                    } else {
                        __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(p.D) + row0 * p.ldd + col;
                        if (EPI == EPI_GELU && p.out2 != nullptr) {
                            __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out2) + row0 * p.ld_out2 + col;
#pragma unroll
                            for (int it = 0; it < NIT; ++it) {
                                float4 v = ldv(it);
                                const float4 d = gelu4_and_grad_fast(v);
                                st_bf16(op + it * o2_step, d);
                                st_bf16(dp + it * d_step, v);
                            }
                        } else {
#pragma unroll
                            for (int it = 0; it < NIT; ++it) {
                                float4 v = ldv(it);
                                if (EPI == EPI_GELU) v = gelu4_fast(v);
                                if (EPI == EPI_GELU_BWD) v = mul_bf16x4(v, axh_c[kAuxBf16 ? it : 0]);
                                if (kColsum) { cs.x += v.x; cs.y += v.y; cs.z += v.z; cs.w += v.w; }
                                st_bf16(dp + it * d_step, v);
                            }
                        }
                    }
                } else if (nvalid > 0) {
                    const bool vD = vec && (p.ldd % 4 == 0), v2 = vec && (p.ld_out2 % 4 == 0);
                    // static register indexing of the prefetched aux operand needs the full unroll; the other
                    // epilogues keep the loop rolled up (instruction-cache footprint)
#pragma unroll(kAuxF32 || kAuxBf16 ? NIT : 1)
                    for (int it = 0; it < NIT; ++it) {
                        const int rr = it * RPA + lrow;
                        const int row = row_base + rr;
                        const uint4 raw = *reinterpret_cast<const uint4*>(stg + rr * RB + (slot(rr, lc4) << 4));
                        float4 v = make_float4(__uint_as_float(raw.x), __uint_as_float(raw.y), __uint_as_float(raw.z),
                                               __uint_as_float(raw.w));
                        if (row >= M_rt) continue;
                        v.x = fmaf(v.x, p.alpha, b4.x); v.y = fmaf(v.y, p.alpha, b4.y);
                        v.z = fmaf(v.z, p.alpha, b4.z); v.w = fmaf(v.w, p.alpha, b4.w);
                        if (EPI == EPI_GELU) {
                            if (p.out_f32) {          // fp32-faithful mode: out2 = pre-activation, exact erf
                                if (p.out2 != nullptr)
                                    st4g(reinterpret_cast<float*>(p.out2) + (size_t)row * p.ld_out2 + col, v, nvalid, v2);
                                v = gelu4_exact(v);
                            } else if (p.out2 != nullptr) {   // training: out2 = gelu'(pre), the factor of the backward
                                const float4 d = gelu4_and_grad_fast(v);
                                st4g(reinterpret_cast<__nv_bfloat16*>(p.out2) + (size_t)row * p.ld_out2 + col, d, nvalid, v2);
                            } else {
                                v = gelu4_fast(v);
                            }
                        } else if (EPI == EPI_RESIDUAL) {
                            const float rs = rsc[kAuxF32 ? it : 0];
                            const float4 a = axf_c[kAuxF32 ? it : 0];
                            v.x = fmaf(rs, v.x, a.x); v.y = fmaf(rs, v.y, a.y);
                            v.z = fmaf(rs, v.z, a.z); v.w = fmaf(rs, v.w, a.w);
                        } else if (EPI == EPI_GELU_BWD) {
                            v = mul_bf16x4(v, axh_c[kAuxBf16 ? it : 0]);
                        }
                        if (kColsum) { cs.x += v.x; cs.y += v.y; cs.z += v.z; cs.w += v.w; }
                        if (EPI == EPI_ATOMIC) {
                            float* o = reinterpret_cast<float*>(p.D) + (size_t)row * p.ldd + col;
                            if (vD) {
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(o), "f"(v.x), "f"(v.y),
                                             "f"(v.z), "f"(v.w)
                                             : "memory");
                            } else {
                                atomicAdd(o, v.x);
                                if (nvalid > 1) atomicAdd(o + 1, v.y);
                                if (nvalid > 2) atomicAdd(o + 2, v.z);
                                if (nvalid > 3) atomicAdd(o + 3, v.w);
                            }
                        } else if (p.out_f32) {
                            st4g(reinterpret_cast<float*>(p.D) + (size_t)row * p.ldd + col, v, nvalid, vD);
                        } else {
                            st4g(reinterpret_cast<__nv_bfloat16*>(p.D) + (size_t)row * p.ldd + col, v, nvalid, vD);
                        }
                    }
                }
                if (kColsum && p.colsum != nullptr) {
                    // fused bias gradient: add up the RPA lanes that hold the same 4 columns, one vector red per chunk
#pragma unroll
                    for (int o = LPR; o < 32; o <<= 1) {
                        cs.x += __shfl_xor_sync(0xffffffffu, cs.x, o);
                        cs.y += __shfl_xor_sync(0xffffffffu, cs.y, o);
                        cs.z += __shfl_xor_sync(0xffffffffu, cs.z, o);
                        cs.w += __shfl_xor_sync(0xffffffffu, cs.w, o);
                    }
                    if (lrow == 0 && nvalid > 0) {
                        float* o = p.colsum + col;
                        if (nvalid >= 4) {
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(o), "f"(cs.x), "f"(cs.y),
                                         "f"(cs.z), "f"(cs.w)
                                         : "memory");
                        } else {
                            atomicAdd(o, cs.x);
                            if (nvalid > 1) atomicAdd(o + 1, cs.y);
                            if (nvalid > 2) atomicAdd(o + 2, cs.z);
                        }
                    }
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CG == 2) mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty[acc]), 0u));
                else mbar_arrive(&tmem_empty[acc]);
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (CG == 2) {
        // neither CTA may leave (or free TMEM) while the pair's MMAs, multicast commits or remote arrivals are in flight
        cluster_sync_all();
        if (warp == 2) tmem_dealloc_pair<2 * BN>(tmem_base);
    } else if (warp == 2) {
        tmem_dealloc<2 * BN>(tmem_base);
    }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (fn == nullptr) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<PFN_encodeTiled>(ptr);
    }
    return fn;
}

// 2-D bf16 tensor map: `inner` contiguous elements, `outer` rows of pitch `ld` elements; box = 64 x box_rows, SW128.
int make_tmap_bf16(CUtensorMap* map, const void* base, long long inner, long long outer, long long ld, int box_rows) {
    PFN_encodeTiled enc = get_encode();
    if (enc == nullptr) return edb_set_error(EDB_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * 2) % 16 != 0)
        return edb_set_error(EDB_ERR_ALIGN, "TMA operand needs a 16-byte aligned base and pitch (ld % 8 == 0)");
    cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return edb_set_error(EDB_ERR_CUDA, "cuTensorMapEncodeTiled failed");
    return EDB_OK;
}

static int g_num_sms = 0;
int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

template <int BN, int STAGES, int EPI, int NEPI, int CG>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tx, const GemmKernelParams& p,
                       cudaStream_t stream) {
    using S = GemmSmem<BN, STAGES, CG>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_bf16_kernel<BN, STAGES, EPI, NEPI, CG>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal);
        if (e != cudaSuccess) return edb_set_error(EDB_ERR_CUDA, cudaGetErrorString(e));
        configured = true;
    }
    const int num_work = p.m_tiles * p.n_tiles * p.split_k;
    const int units = num_sms() / CG;          // CTAs, or CTA pairs (one per TPC)
    const int grid = (num_work < units ? num_work : units) * CG;
    if (CG == 1) {
        gemm_bf16_kernel<BN, STAGES, EPI, NEPI, CG><<<grid, gemm_threads<NEPI>(), S::kTotal, stream>>>(ta, tb, tx, p);
    } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid, 1, 1);
        cfg.blockDim = dim3(gemm_threads<NEPI>(), 1, 1);
        cfg.dynamicSmemBytes = S::kTotal;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CG;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_bf16_kernel<BN, STAGES, EPI, NEPI, CG>, ta, tb, tx, p);
        if (e != cudaSuccess) return edb_set_error(EDB_ERR_CUDA, cudaGetErrorString(e));
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return edb_set_error(EDB_ERR_CUDA, cudaGetErrorString(e));
    return EDB_OK;
}

// 0 = automatic (CTA pairs where the shape allows), 1 = single-CTA tiles only (A/B comparisons, tests)
static int g_gemm_mode = 0;
int gemm_set_mode(int mode) {
    if (mode < 0 || mode > 1) return edb_set_error(EDB_ERR_UNSUPPORTED, "gemm mode: 0 = auto (CTA pairs), 1 = single CTA");
    g_gemm_mode = mode;
    return EDB_OK;
}

int gemm_bf16(const EdbGemmDesc& g, cudaStream_t stream) {
    if (g.M <= 0 || g.N <= 0 || g.K <= 0) return edb_set_error(EDB_ERR_SHAPE, "gemm: non-positive dimension");
    const int BN = (g.N > 128) ? 256 : 128;
    // CTA pairs (256 x 256 per MMA) wherever there is more than one 128-row tile to pair up
    const int CG = (BN == 256 && g.M > BM && g_gemm_mode == 0) ? 2 : 1;
    GemmKernelParams p{};
    p.M = g.M; p.N = g.N; p.K = g.K;
    p.m_tiles = (g.M + BM * CG - 1) / (BM * CG);
    p.n_tiles = (g.N + BN - 1) / BN;
    p.kb_total = (g.K + BK - 1) / BK;
    int split = g.split_k > 0 ? g.split_k : 1;
    if (split > p.kb_total) split = p.kb_total;
    p.kb_per_split = (p.kb_total + split - 1) / split;
    p.split_k = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
    if (p.split_k > 1 && g.epilogue != EPI_ATOMIC)
        return edb_set_error(EDB_ERR_SHAPE, "gemm: split_k > 1 requires the atomic-accumulate epilogue");
    if (g.epilogue == EPI_ATOMIC && !g.out_f32)
        return edb_set_error(EDB_ERR_SHAPE, "gemm: atomic-accumulate epilogue needs an fp32 output");
    p.a_mn = g.a_mn_major; p.b_mn = g.b_mn_major;
    p.idesc = make_idesc_bf16(BM * CG, BN, g.a_mn_major, g.b_mn_major);
    p.D = g.D; p.ldd = g.ldd; p.out_f32 = g.out_f32; p.epi = g.epilogue;
    p.bias = g.bias; p.aux = g.aux; p.ld_aux = g.ld_aux; p.aux_f32 = g.aux_f32;
    p.out2 = g.out2; p.ld_out2 = g.ld_out2; p.alpha = g.alpha;
    p.row_scale = g.row_scale; p.scale_group = g.scale_group > 0 ? g.scale_group : 1;
    p.M_dev = g.M_dev; p.K_dev = g.K_dev;
    p.colsum = g.colsum;
    if (g.colsum != nullptr && g.epilogue != EPI_GELU_BWD && g.epilogue != EPI_STORE)
        return edb_set_error(EDB_ERR_UNSUPPORTED, "gemm: colsum is fused into the store and GELU-backward epilogues only");
    if (g.epilogue == EPI_RESIDUAL && (g.aux == nullptr || !g.aux_f32 || !g.out_f32))
        return edb_set_error(EDB_ERR_SHAPE, "gemm: the residual epilogue needs fp32 aux and fp32 output");
    if (g.epilogue == EPI_GELU_BWD && (g.aux == nullptr || g.aux_f32))
        return edb_set_error(EDB_ERR_SHAPE, "gemm: the GELU-backward epilogue needs the saved bf16 gelu' factor as aux");

    CUtensorMap ta, tb;
    int rc;
    if (!g.a_mn_major) rc = make_tmap_bf16(&ta, g.A, g.K, g.M, g.lda, BM);   // [M rows, K inner]
    else               rc = make_tmap_bf16(&ta, g.A, g.M, g.K, g.lda, BK);   // [K rows, M inner]
    if (rc != EDB_OK) return rc;
    if (!g.b_mn_major) rc = make_tmap_bf16(&tb, g.B, g.K, g.N, g.ldb, BN / CG);
    else               rc = make_tmap_bf16(&tb, g.B, g.N, g.K, g.ldb, BK);
    if (rc != EDB_OK) return rc;
    // bf16 aux (saved gelu'): a 64-column x 32-row box per epilogue warp, used for L2 prefetches of the next tile's slab
    CUtensorMap tx = ta;
    if (g.epilogue == EPI_GELU_BWD && (g.ld_aux * 2) % 16 == 0 && (reinterpret_cast<uintptr_t>(g.aux) & 15) == 0 &&
        make_tmap_bf16(&tx, g.aux, g.N, g.M, g.ld_aux, 32) == EDB_OK)
        p.aux_tmap_ok = 1;
#define EDB_LAUNCH_EPI(E, W)                                                          \
    case E:                                                                           \
        if (CG == 2) return launch_gemm<256, 6, E, W, 2>(ta, tb, tx, p, stream);       \
        if (BN == 256) return launch_gemm<256, 4, E, W, 1>(ta, tb, tx, p, stream);     \
        return launch_gemm<128, 6, E, W, 1>(ta, tb, tx, p, stream);
    switch (g.epilogue) {
        EDB_LAUNCH_EPI(EPI_STORE, 16)
        EDB_LAUNCH_EPI(EPI_GELU, EDB_GELU_EPI_WARPS)
        EDB_LAUNCH_EPI(EPI_RESIDUAL, 8)
        EDB_LAUNCH_EPI(EPI_GELU_BWD, EDB_GELU_EPI_WARPS)
        EDB_LAUNCH_EPI(EPI_ATOMIC, 8)
        default:
            return edb_set_error(EDB_ERR_UNSUPPORTED, "gemm: unknown epilogue");
    }
#undef EDB_LAUNCH_EPI
}

}  // namespace edb

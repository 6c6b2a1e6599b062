// Persistent, warp-specialised bf16 GEMM for sm_100a:  D[M,N] = epilogue( sum_k A(m,k) * B(n,k) ).
//
//   warp 0      : TMA producer  (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier full/empty)
//   warp 1      : MMA issuer    (one elected lane issues tcgen05.mma, 128 x BN x 16 per instruction, fp32 in TMEM)
//   warp 2      : TMEM allocator (2 accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1)
//   warps 4..11 : epilogue      (tcgen05.ld -> bias / GELU / residual / GELU' / split-K reduce -> global)
//
// Both operands may be K-major (reduction dim contiguous) or MN-major (reduction dim strided); that covers the
// forward (X W^T), dgrad (dY W) and wgrad (dY^T X) products of every Linear on EDITOR's hot path
// (reference: modeling/backbones/vit_pytorch.py:139-145,184-198,158-168,240-258) without any transposed copies.
#include "ptx.cuh"
#include "abi_internal.h"

namespace edb {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int kNumEpiWarps = 8;
constexpr int kThreads = 128 + kNumEpiWarps * 32;

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
};

__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
    const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
    return cdf + x * pdf;
}

template <int BN, int STAGES>
struct GemmSmem {
    static constexpr int kABytes = BM * BK * 2;
    static constexpr int kBBytes = BN * BK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kBarOffset = STAGES * kStageBytes;
    static constexpr int kTotal = kBarOffset + 256 + 1024;  // + alignment slack
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const GemmKernelParams p) {
    using S = GemmSmem<BN, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
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
            mbar_init(&tmem_empty[i], kNumEpiWarps);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<2 * BN>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    const int num_work = p.m_tiles * p.n_tiles * p.split_k;
    constexpr int GM = 16;  // m-tiles per raster group (keeps the group's A tiles + all of B resident in L2)

    auto decode = [&](int w, int& tm, int& tn, int& ks) {
        ks = w % p.split_k;
        int t = w / p.split_k;
        const int per_group = GM * p.n_tiles;
        const int g = t / per_group;
        const int within = t - g * per_group;
        const int gsize = min(GM, p.m_tiles - g * GM);
        tm = g * GM + within % gsize;
        tn = within / gsize;
    };

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
                int tm, tn, ks;
                decode(w, tm, tn, ks);
                const int kb0 = ks * p.kb_per_split;
                const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * S::kStageBytes;
                    uint8_t* sb = sa + S::kABytes;
                    mbar_expect_tx(&full_bar[stage], S::kStageBytes);
                    if (!p.a_mn) {
                        tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BK, tm * BM);
                    } else {
#pragma unroll
                        for (int j = 0; j < BM / 64; ++j)
                            tma_load_2d(sa + j * (BK * 128), &tmap_a, &full_bar[stage], tm * BM + j * 64, kb * BK);
                    }
                    if (!p.b_mn) {
                        tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BK, tn * BN);
                    } else {
#pragma unroll
                        for (int j = 0; j < BN / 64; ++j)
                            tma_load_2d(sb + j * (BK * 128), &tmap_b, &full_bar[stage], tn * BN + j * 64, kb * BK);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            const uint32_t a_lbo = p.a_mn ? BK * 128 : 0, b_lbo = p.b_mn ? BK * 128 : 0;
            const uint32_t a_kstep = p.a_mn ? 2048 : 32, b_kstep = p.b_mn ? 2048 : 32;
            for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
                int tm, tn, ks;
                decode(w, tm, tn, ks);
                const int kb0 = ks * p.kb_per_split;
                const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
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
                        tc_mma_bf16(d_tmem, da, db, p.idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                    tc_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int ew = warp - 4;
        const int quarter = warp & 3;  // TMEM lane quarter this warp may read
        const int half = ew >> 2;      // column half
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
            int tm, tn, ks;
            decode(w, tm, tn, ks);
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const int row = tm * BM + quarter * 32 + lane;
            const bool row_ok = row < p.M;
            const uint32_t t_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
#pragma unroll 1
            for (int c = 0; c < BN / 2; c += 32) {
                const int col_t = half * (BN / 2) + c;
                const int col0 = tn * BN + col_t;
                uint32_t r[32];
                tmem_ld_32x32(t_base + col_t, r);
                tmem_ld_wait();
                if (!row_ok || col0 >= p.N) continue;
                float v[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * p.alpha;
                const bool full = (col0 + 32 <= p.N);
                if (p.bias != nullptr && p.epi != EPI_ATOMIC) {
                    if (full) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + i));
                            v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (col0 + i < p.N) v[i] += __ldg(p.bias + col0 + i);
                    }
                }
                if (p.epi == EPI_GELU) {
                    if (p.out2 != nullptr) {
                        if (p.out_f32) {
                            float* o = reinterpret_cast<float*>(p.out2) + (size_t)row * p.ld_out2 + col0;
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                if (col0 + i < p.N) o[i] = v[i];
                        } else {
                            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out2) + (size_t)row * p.ld_out2 + col0;
                            if (full) {
#pragma unroll
                                for (int i = 0; i < 32; i += 8) {
                                    uint4 pk;
                                    __nv_bfloat162 t0 = __floats2bfloat162_rn(v[i], v[i + 1]);
                                    __nv_bfloat162 t1 = __floats2bfloat162_rn(v[i + 2], v[i + 3]);
                                    __nv_bfloat162 t2 = __floats2bfloat162_rn(v[i + 4], v[i + 5]);
                                    __nv_bfloat162 t3 = __floats2bfloat162_rn(v[i + 6], v[i + 7]);
                                    pk.x = *reinterpret_cast<uint32_t*>(&t0);
                                    pk.y = *reinterpret_cast<uint32_t*>(&t1);
                                    pk.z = *reinterpret_cast<uint32_t*>(&t2);
                                    pk.w = *reinterpret_cast<uint32_t*>(&t3);
                                    *reinterpret_cast<uint4*>(o + i) = pk;
                                }
                            } else {
#pragma unroll
                                for (int i = 0; i < 32; ++i)
                                    if (col0 + i < p.N) o[i] = __float2bfloat16(v[i]);
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = gelu_exact(v[i]);
                } else if (p.epi == EPI_RESIDUAL) {
                    const float* a = reinterpret_cast<const float*>(p.aux) + (size_t)row * p.ld_aux + col0;
                    const float rsc = p.row_scale ? p.row_scale[row / p.scale_group] : 1.0f;
                    if (full) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const float4 a4 = *reinterpret_cast<const float4*>(a + i);
                            v[i] = a4.x + rsc * v[i]; v[i + 1] = a4.y + rsc * v[i + 1];
                            v[i + 2] = a4.z + rsc * v[i + 2]; v[i + 3] = a4.w + rsc * v[i + 3];
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (col0 + i < p.N) v[i] = a[i] + rsc * v[i];
                    }
                } else if (p.epi == EPI_GELU_BWD) {
                    if (p.aux_f32) {
                        const float* a = reinterpret_cast<const float*>(p.aux) + (size_t)row * p.ld_aux + col0;
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (col0 + i < p.N) v[i] *= gelu_grad(a[i]);
                    } else {
                        const __nv_bfloat16* a =
                            reinterpret_cast<const __nv_bfloat16*>(p.aux) + (size_t)row * p.ld_aux + col0;
                        if (full) {
#pragma unroll
                            for (int i = 0; i < 32; i += 8) {
                                const uint4 pk = *reinterpret_cast<const uint4*>(a + i);
                                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&pk);
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const float2 f = __bfloat1622float2(h[j]);
                                    v[i + 2 * j] *= gelu_grad(f.x);
                                    v[i + 2 * j + 1] *= gelu_grad(f.y);
                                }
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                if (col0 + i < p.N) v[i] *= gelu_grad(__bfloat162float(a[i]));
                        }
                    }
                }
                // ---- store
                if (p.epi == EPI_ATOMIC) {
                    float* o = reinterpret_cast<float*>(p.D) + (size_t)row * p.ldd + col0;
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (col0 + i < p.N) atomicAdd(o + i, v[i]);
                } else if (p.out_f32) {
                    float* o = reinterpret_cast<float*>(p.D) + (size_t)row * p.ldd + col0;
                    if (full && (p.ldd % 4 == 0)) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4)
                            *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (col0 + i < p.N) o[i] = v[i];
                    }
                } else {
                    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.D) + (size_t)row * p.ldd + col0;
                    if (full && (p.ldd % 8 == 0)) {
#pragma unroll
                        for (int i = 0; i < 32; i += 8) {
                            uint4 pk;
                            __nv_bfloat162 t0 = __floats2bfloat162_rn(v[i], v[i + 1]);
                            __nv_bfloat162 t1 = __floats2bfloat162_rn(v[i + 2], v[i + 3]);
                            __nv_bfloat162 t2 = __floats2bfloat162_rn(v[i + 4], v[i + 5]);
                            __nv_bfloat162 t3 = __floats2bfloat162_rn(v[i + 6], v[i + 7]);
                            pk.x = *reinterpret_cast<uint32_t*>(&t0);
                            pk.y = *reinterpret_cast<uint32_t*>(&t1);
                            pk.z = *reinterpret_cast<uint32_t*>(&t2);
                            pk.w = *reinterpret_cast<uint32_t*>(&t3);
                            *reinterpret_cast<uint4*>(o + i) = pk;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (col0 + i < p.N) o[i] = __float2bfloat16(v[i]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<2 * BN>(tmem_base);
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

template <int BN, int STAGES>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmKernelParams& p, cudaStream_t stream) {
    using S = GemmSmem<BN, STAGES>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_bf16_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             S::kTotal);
        if (e != cudaSuccess) return edb_set_error(EDB_ERR_CUDA, cudaGetErrorString(e));
        configured = true;
    }
    const int num_work = p.m_tiles * p.n_tiles * p.split_k;
    const int grid = num_work < num_sms() ? num_work : num_sms();
    gemm_bf16_kernel<BN, STAGES><<<grid, kThreads, S::kTotal, stream>>>(ta, tb, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return edb_set_error(EDB_ERR_CUDA, cudaGetErrorString(e));
    return EDB_OK;
}

int gemm_bf16(const EdbGemmDesc& g, cudaStream_t stream) {
    if (g.M <= 0 || g.N <= 0 || g.K <= 0) return edb_set_error(EDB_ERR_SHAPE, "gemm: non-positive dimension");
    const int BN = (g.N > 128) ? 256 : 128;
    GemmKernelParams p{};
    p.M = g.M; p.N = g.N; p.K = g.K;
    p.m_tiles = (g.M + BM - 1) / BM;
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
    p.idesc = make_idesc_bf16(BM, BN, g.a_mn_major, g.b_mn_major);
    p.D = g.D; p.ldd = g.ldd; p.out_f32 = g.out_f32; p.epi = g.epilogue;
    p.bias = g.bias; p.aux = g.aux; p.ld_aux = g.ld_aux; p.aux_f32 = g.aux_f32;
    p.out2 = g.out2; p.ld_out2 = g.ld_out2; p.alpha = g.alpha;
    p.row_scale = g.row_scale; p.scale_group = g.scale_group > 0 ? g.scale_group : 1;
    if (g.epilogue == EPI_RESIDUAL && (g.aux == nullptr || !g.aux_f32 || !g.out_f32))
        return edb_set_error(EDB_ERR_SHAPE, "gemm: the residual epilogue needs fp32 aux and fp32 output");

    CUtensorMap ta, tb;
    int rc;
    if (!g.a_mn_major) rc = make_tmap_bf16(&ta, g.A, g.K, g.M, g.lda, BM);   // [M rows, K inner]
    else               rc = make_tmap_bf16(&ta, g.A, g.M, g.K, g.lda, BK);   // [K rows, M inner]
    if (rc != EDB_OK) return rc;
    if (!g.b_mn_major) rc = make_tmap_bf16(&tb, g.B, g.K, g.N, g.ldb, BN);
    else               rc = make_tmap_bf16(&tb, g.B, g.N, g.K, g.ldb, BK);
    if (rc != EDB_OK) return rc;
    if (BN == 256) return launch_gemm<256, 4>(ta, tb, p, stream);
    return launch_gemm<128, 6>(ta, tb, p, stream);
}

}  // namespace edb

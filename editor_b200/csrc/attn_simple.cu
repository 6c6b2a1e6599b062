// Variable-length multi-head attention on CUDA cores (fp32 math; fp32 or bf16 storage), one CTA per (sequence, head).
// Used (a) for the fp32-faithful path (EDB_PREC_FP32) of Attention.forward (vit_pytorch.py:184-198), whose softmax maps
// feed the bit-exact token selection, and (b) for HMA's packed AttentionMask (vit_pytorch.py:240-258): after packing
// only kept tokens exist, so masked_fill(-65504)/"* mask" reduce to plain attention over 1+n_sel (or 3(1+n_sel)) tokens
// (SURVEY.md App. A-5).  The tensor-core (tcgen05) kernel for the 129-token backbone lives in attn_tc.cu.
#include "abi_internal.h"

namespace edb {

constexpr int HD = 64;  // head dim

template <typename T> struct SmemPad;
template <> struct SmemPad<float> { static constexpr int kStride = HD + 1; };          // odd number of 32-bit words
template <> struct SmemPad<__nv_bfloat16> { static constexpr int kStride = HD + 2; };  // 33 words

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ void from_f(float& d, float v) { d = v; }
__device__ __forceinline__ void from_f(__nv_bfloat16& d, float v) { d = __float2bfloat16(v); }

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

struct AttnArgs {
    const void* qkv; long long ld_qkv;      // rows x (3*H*64): [q | k | v], head h at column h*64 of each third
    void* out; long long ld_out;            // rows x (H*64)
    void* P; long long p_rows; long long ldp;  // [(seq*H + h)][p_rows][ldp] post-softmax maps (optional in fwd)
    const int* seq_off;                     // nseq+1 row offsets, or nullptr -> seq s starts at s*fixed_len
    int fixed_len, nseq, H, max_len;
    float scale;
    // backward only
    const void* d_out; long long ld_dout;
    void* d_qkv;
};

template <typename T, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) attn_simple_fwd_kernel(const AttnArgs a) {
    constexpr int ST = SmemPad<T>::kStride;
    extern __shared__ uint8_t smem_raw[];
    const int s = blockIdx.x / a.H, h = blockIdx.x % a.H;
    const int off = a.seq_off ? a.seq_off[s] : s * a.fixed_len;
    const int L = a.seq_off ? a.seq_off[s + 1] - off : a.fixed_len;
    if (L <= 0) return;
    T* Ks = reinterpret_cast<T*>(smem_raw);
    T* Vs = Ks + (size_t)a.max_len * ST;
    float* pbuf = reinterpret_cast<float*>(Vs + (size_t)a.max_len * ST);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const T* qkv = reinterpret_cast<const T*>(a.qkv);
    const int HC = a.H * HD;
    for (int i = threadIdx.x; i < L * HD; i += WARPS * 32) {
        const int r = i / HD, d = i % HD;
        const T* row = qkv + (size_t)(off + r) * a.ld_qkv + h * HD + d;
        Ks[r * ST + d] = row[HC];
        Vs[r * ST + d] = row[2 * HC];
    }
    __syncthreads();
    float* pw = pbuf + warp * a.max_len;
    T* Pg = reinterpret_cast<T*>(a.P);
    T* out = reinterpret_cast<T*>(a.out);
    for (int i = warp; i < L; i += WARPS) {
        float q[HD];
        const T* qrow = qkv + (size_t)(off + i) * a.ld_qkv + h * HD;
#pragma unroll
        for (int d = 0; d < HD; ++d) q[d] = to_f(qrow[d]);
        float sc[8];
        float mx = -INFINITY;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            const int j = jj * 32 + lane;
            float acc = -INFINITY;
            if (j < L) {
                acc = 0.f;
                const T* kr = Ks + j * ST;
#pragma unroll
                for (int d = 0; d < HD; ++d) acc += q[d] * to_f(kr[d]);
                acc *= a.scale;
            }
            sc[jj] = acc;
            mx = fmaxf(mx, acc);
        }
        mx = wmax(mx);
        float sum = 0.f;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            sc[jj] = (jj * 32 + lane < L) ? expf(sc[jj] - mx) : 0.f;
            sum += sc[jj];
        }
        const float inv = 1.0f / wsum(sum);
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            const int j = jj * 32 + lane;
            float p = sc[jj] * inv;
            if (Pg != nullptr && j < a.ldp) {
                T pv;
                from_f(pv, j < L ? p : 0.f);
                Pg[((size_t)blockIdx.x * a.p_rows + i) * a.ldp + j] = pv;
                p = to_f(pv);  // P.V uses the stored (possibly bf16-rounded) probabilities, like the backward will
            }
            if (j < L) pw[j] = p;
        }
        __syncwarp();
        float o0 = 0.f, o1 = 0.f;
        for (int j = 0; j < L; ++j) {
            const float p = pw[j];
            o0 += p * to_f(Vs[j * ST + lane]);
            o1 += p * to_f(Vs[j * ST + lane + 32]);
        }
        T* orow = out + (size_t)(off + i) * a.ld_out + h * HD;
        from_f(orow[lane], o0);
        from_f(orow[lane + 32], o1);
        __syncwarp();
    }
}

template <typename T, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) attn_simple_bwd_kernel(const AttnArgs a) {
    // Two phases share two shared-memory tiles: phase A holds (K, V) and produces delta_i and dQ_i, one query row per warp;
    // phase B reloads the tiles with (Q, dO) and produces dK_j, dV_j, one key row per warp (dS column recomputed).
    constexpr int ST = SmemPad<T>::kStride;
    extern __shared__ uint8_t smem_raw[];
    const int s = blockIdx.x / a.H, h = blockIdx.x % a.H;
    const int off = a.seq_off ? a.seq_off[s] : s * a.fixed_len;
    const int L = a.seq_off ? a.seq_off[s + 1] - off : a.fixed_len;
    if (L <= 0) return;
    const int ML = a.max_len;
    T* T0 = reinterpret_cast<T*>(smem_raw);            // K, then Q
    T* T1 = T0 + (size_t)ML * ST;                      // V, then dO
    float* delta = reinterpret_cast<float*>(T1 + (size_t)ML * ST);
    float* pbuf = delta + ML;                          // [WARPS][2][ML]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const T* qkv = reinterpret_cast<const T*>(a.qkv);
    const T* dO = reinterpret_cast<const T*>(a.d_out);
    const T* Pg = reinterpret_cast<const T*>(a.P) + (size_t)blockIdx.x * a.p_rows * a.ldp;
    T* dqkv = reinterpret_cast<T*>(a.d_qkv);
    const int HC = a.H * HD;
    for (int i = threadIdx.x; i < L * HD; i += WARPS * 32) {
        const int r = i / HD, d = i % HD;
        const T* row = qkv + (size_t)(off + r) * a.ld_qkv + h * HD + d;
        T0[r * ST + d] = row[HC];
        T1[r * ST + d] = row[2 * HC];
    }
    __syncthreads();
    float* p1 = pbuf + (size_t)warp * 2 * ML;
    float* p2 = p1 + ML;
    for (int i = warp; i < L; i += WARPS) {
        float g[HD];
        const T* grow = dO + (size_t)(off + i) * a.ld_dout + h * HD;
#pragma unroll
        for (int d = 0; d < HD; ++d) g[d] = to_f(grow[d]);
        float dsum = 0.f;
        float dp[8], pp[8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            const int j = jj * 32 + lane;
            dp[jj] = 0.f; pp[jj] = 0.f;
            if (j < L) {
                float acc = 0.f;
                const T* vr = T1 + j * ST;
#pragma unroll
                for (int d = 0; d < HD; ++d) acc += g[d] * to_f(vr[d]);
                dp[jj] = acc;
                pp[jj] = to_f(Pg[(size_t)i * a.ldp + j]);
                dsum += acc * pp[jj];
            }
        }
        dsum = wsum(dsum);
        if (lane == 0) delta[i] = dsum;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            const int j = jj * 32 + lane;
            if (j < L) p1[j] = pp[jj] * (dp[jj] - dsum) * a.scale;
        }
        __syncwarp();
        float o0 = 0.f, o1 = 0.f;
        for (int j = 0; j < L; ++j) {
            const float ds = p1[j];
            o0 += ds * to_f(T0[j * ST + lane]);
            o1 += ds * to_f(T0[j * ST + lane + 32]);
        }
        T* orow = dqkv + (size_t)(off + i) * a.ld_qkv + h * HD;
        from_f(orow[lane], o0);
        from_f(orow[lane + 32], o1);
        __syncwarp();
    }
    __syncthreads();
    for (int i = threadIdx.x; i < L * HD; i += WARPS * 32) {
        const int r = i / HD, d = i % HD;
        T0[r * ST + d] = qkv[(size_t)(off + r) * a.ld_qkv + h * HD + d];
        T1[r * ST + d] = dO[(size_t)(off + r) * a.ld_dout + h * HD + d];
    }
    __syncthreads();
    for (int j = warp; j < L; j += WARPS) {
        float v[HD];
        const T* vrow = qkv + (size_t)(off + j) * a.ld_qkv + 2 * HC + h * HD;
#pragma unroll
        for (int d = 0; d < HD; ++d) v[d] = to_f(vrow[d]);
        for (int ii = 0; ii < 8; ++ii) {
            const int i = ii * 32 + lane;
            if (i < L) {
                float acc = 0.f;
                const T* gr = T1 + i * ST;
#pragma unroll
                for (int d = 0; d < HD; ++d) acc += v[d] * to_f(gr[d]);
                const float p = to_f(Pg[(size_t)i * a.ldp + j]);
                p1[i] = p * (acc - delta[i]) * a.scale;
                p2[i] = p;
            }
        }
        __syncwarp();
        float k0 = 0.f, k1 = 0.f, v0 = 0.f, v1 = 0.f;
        for (int i = 0; i < L; ++i) {
            const float ds = p1[i], p = p2[i];
            k0 += ds * to_f(T0[i * ST + lane]);
            k1 += ds * to_f(T0[i * ST + lane + 32]);
            v0 += p * to_f(T1[i * ST + lane]);
            v1 += p * to_f(T1[i * ST + lane + 32]);
        }
        T* krow = dqkv + (size_t)(off + j) * a.ld_qkv + HC + h * HD;
        from_f(krow[lane], k0);
        from_f(krow[lane + 32], k1);
        from_f(krow[HC + lane], v0);
        from_f(krow[HC + lane + 32], v1);
        __syncwarp();
    }
}

constexpr int kSimpleWarps = 8;

template <typename T>
static int launch_simple(const AttnArgs& a, bool bwd, cudaStream_t st) {
    constexpr int ST = SmemPad<T>::kStride;
    size_t smem;
    if (!bwd) smem = (size_t)2 * a.max_len * ST * sizeof(T) + (size_t)kSimpleWarps * a.max_len * sizeof(float);
    else smem = (size_t)2 * a.max_len * ST * sizeof(T) + (size_t)(1 + 2 * kSimpleWarps) * a.max_len * sizeof(float);
    if (smem > 227 * 1024) return edb_set_error(EDB_ERR_SHAPE, "attention: sequence too long for shared memory");
    auto kern = bwd ? attn_simple_bwd_kernel<T, kSimpleWarps> : attn_simple_fwd_kernel<T, kSimpleWarps>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return edb_set_error(EDB_ERR_CUDA, cudaGetErrorString(e));
    kern<<<a.nseq * a.H, kSimpleWarps * 32, smem, st>>>(a);
    EDB_CHECK_LAUNCH();
    return EDB_OK;
}

int attention_simple(const EdbAttnDesc& d, bool bwd, cudaStream_t st) {
    if (d.nseq <= 0) return EDB_OK;
    if (d.max_len <= 0 || d.max_len > 256) return edb_set_error(EDB_ERR_SHAPE, "attention: max_len must be in 1..256");
    if (d.heads <= 0) return edb_set_error(EDB_ERR_SHAPE, "attention: heads");
    if (bwd && (d.P == nullptr || d.d_out == nullptr || d.d_qkv == nullptr))
        return edb_set_error(EDB_ERR_SHAPE, "attention backward needs P, d_out and d_qkv");
    AttnArgs a{};
    a.qkv = d.qkv; a.ld_qkv = d.ld_qkv; a.out = d.out; a.ld_out = d.ld_out;
    a.P = d.P; a.p_rows = d.p_rows; a.ldp = d.ldp;
    a.seq_off = d.seq_off; a.fixed_len = d.fixed_len; a.nseq = d.nseq; a.H = d.heads; a.max_len = d.max_len;
    a.scale = d.scale; a.d_out = d.d_out; a.ld_dout = d.ld_dout; a.d_qkv = d.d_qkv;
    if (d.f32) return launch_simple<float>(a, bwd, st);
    return launch_simple<__nv_bfloat16>(a, bwd, st);
}

}  // namespace edb

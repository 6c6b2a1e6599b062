// Spatial-Frequency Token Selection and the packing that turns HMA's masked attention into dense var-len work.
//
//   freq_counts      Frequency.py:42-56,65-81 : positive-pixel count per 16x16 window of the 3-modality / 3-channel mean
//                    (Haar DWT -> average -> IDWT is the identity on the mean)        HBM-bound
//   topk_rank_mask   Frequency.py:58-63, SFTS.py:155-158 : torch.topk -> sort -> scatter_ as a 128-bit set, with the CUDA
//                    tie rule (strictly greater first, then equal values by ascending index)
//   rollout_topk     SFTS.py:145-162 : cls row of A_11...A_0 as a row-vector chain, per-head top-k, OR into the index
//   index_finalize   SFTS.py:183-190 + make_model.py:197-198 : popcounts, packed row offsets
//   sfts_pack        SFTS.py:208-222 : gather kept rows (cls + selected) into packed per-modality matrices, BCC loss
//   joint gather / pool  vit_pytorch.py:324, make_model.py:186-203
#include "abi_internal.h"

namespace edb {

constexpr int DT = 768;
constexpr int NP = 128;  // patch tokens per image in every shipped config (16x8 or 8x16)

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// ------------------------------------------------------------------------------------------ frequency counts
// grid = (B, H/16); block = 256.  Each thread sums 4 consecutive pixels x 9 planes.
__global__ void __launch_bounds__(256) freq_counts_kernel(const float* __restrict__ rgb, const float* __restrict__ ni,
                                                          const float* __restrict__ ti, int H, int W,
                                                          int* __restrict__ counts) {
    __shared__ int cnt[32];
    const int b = blockIdx.x, py = blockIdx.y, nx = W / 16;
    if (threadIdx.x < 32) cnt[threadIdx.x] = 0;
    __syncthreads();
    const int tpr = W / 4;                 // threads per pixel row
    const int rows_per_it = 256 / tpr;     // W = 128 -> 8, W = 256 -> 4
    const int xq = threadIdx.x % tpr, yr = threadIdx.x / tpr;
    const size_t plane = (size_t)H * W;
    int c = 0;
    for (int y = yr; y < 16; y += rows_per_it) {
        const size_t o = (size_t)b * 3 * plane + (size_t)(py * 16 + y) * W + xq * 4;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const float4 a = ld4(rgb + o + ch * plane), d = ld4(ni + o + ch * plane), e = ld4(ti + o + ch * plane);
            // (x+y+z)/3 per channel, then the channel mean (Frequency.py:71-74,44); only the sign is used
            s.x += (a.x + d.x + e.x); s.y += (a.y + d.y + e.y); s.z += (a.z + d.z + e.z); s.w += (a.w + d.w + e.w);
        }
        c += (s.x > 0.f) + (s.y > 0.f) + (s.z > 0.f) + (s.w > 0.f);
    }
    atomicAdd(&cnt[xq / 4], c);
    __syncthreads();
    if (threadIdx.x < nx) counts[(size_t)b * (H / 16) * nx + py * nx + threadIdx.x] = cnt[threadIdx.x];
}

int freq_counts(const float* rgb, const float* ni, const float* ti, int B, int H, int W, int* counts, cudaStream_t st) {
    if (B <= 0) return EDB_OK;
    if (H % 16 || W % 16 || (W != 128 && W != 256 && W != 64) || (H / 16) * (W / 16) != NP)
        return edb_set_error(EDB_ERR_SHAPE, "freq_counts: image must give 128 patch tokens (256x128 or 128x256)");
    freq_counts_kernel<<<dim3(B, H / 16), 256, 0, st>>>(rgb, ni, ti, H, W, counts);
    EDB_CHECK_LAUNCH();
    return EDB_OK;
}

// ------------------------------------------------------------------------------------------ top-k as a bit set
// selected(j) <=> #{i : v_i > v_j  or (v_i == v_j and i < j)} < k      (CUDA torch.topk tie rule, tools/probe_topk.py)
template <typename T>
__device__ __forceinline__ bool rank_select(const T* vals, int j, int k) {
    const T v = vals[j];
    int rank = 0;
#pragma unroll 8
    for (int i = 0; i < NP; ++i) {
        const T u = vals[i];
        rank += (u > v) || (u == v && i < j);
    }
    return rank < k;
}

template <typename T>
__global__ void __launch_bounds__(NP) topk_mask_kernel(const T* __restrict__ vals, long long ld, int k,
                                                       unsigned* __restrict__ mask, int accumulate) {
    __shared__ T sv[NP];
    const int row = blockIdx.x, j = threadIdx.x;
    sv[j] = vals[(size_t)row * ld + j];
    __syncthreads();
    const unsigned bits = __ballot_sync(0xffffffffu, rank_select(sv, j, k));
    if ((j & 31) == 0) {
        if (accumulate) atomicOr(&mask[row * 4 + (j >> 5)], bits);
        else mask[row * 4 + (j >> 5)] = bits;
    }
}

int topk_mask(const void* vals, int vals_f32, long long ld, int rows, int n, int k, unsigned* mask, int accumulate,
              cudaStream_t st) {
    if (rows <= 0) return EDB_OK;
    if (n != NP || k < 0 || k > NP) return edb_set_error(EDB_ERR_SHAPE, "topk_mask: rows must have 128 entries, 0<=k<=128");
    if (vals_f32) topk_mask_kernel<float><<<rows, NP, 0, st>>>((const float*)vals, ld, k, mask, accumulate);
    else topk_mask_kernel<int><<<rows, NP, 0, st>>>((const int*)vals, ld, k, mask, accumulate);
    EDB_CHECK_LAUNCH();
    return EDB_OK;
}

// ------------------------------------------------------------------------------------------ attention rollout + top-k
struct RolloutArgs {
    const void* maps[16];   // layer l: [(seq*H + h)][p_rows][ldp]
    int layers, H, B, k;
    long long p_rows, ldp;
    unsigned* index;        // [B][4]  OR target (sequence s = m*B + b contributes to sample b)
    unsigned* mod_mask;     // [S][4]  optional per-sequence result (tests)
    float* rows_out;        // [S*H][128] optional rollout rows (tests)
};

__device__ __forceinline__ float ldv(const float* p) { return *p; }
__device__ __forceinline__ float ldv(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <typename T>
__global__ void __launch_bounds__(160) rollout_topk_kernel(const RolloutArgs a) {
    __shared__ float r[2][132];
    __shared__ float sv[NP];
    const int blk = blockIdx.x;           // seq*H + h
    const int j = threadIdx.x;
    const int NT = NP + 1;
    const size_t base = (size_t)blk * a.p_rows * a.ldp;
    const T* last = reinterpret_cast<const T*>(a.maps[a.layers - 1]) + base;
    if (j < NT) r[0][j] = ldv(last + j);  // row 0 of the last layer's map
    __syncthreads();
    int cur = 0;
    for (int l = a.layers - 2; l >= 0; --l) {
        const T* m = reinterpret_cast<const T*>(a.maps[l]) + base;
        float acc = 0.f;
        if (j < NT) {
            int i = 0;
            for (; i + 8 <= NT; i += 8) {
                float v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = ldv(m + (size_t)(i + u) * a.ldp + j);
#pragma unroll
                for (int u = 0; u < 8; ++u) acc += r[cur][i + u] * v[u];
            }
            for (; i < NT; ++i) acc += r[cur][i] * ldv(m + (size_t)i * a.ldp + j);
            r[cur ^ 1][j] = acc;
        }
        __syncthreads();
        cur ^= 1;
    }
    if (j >= 1 && j < NT) {
        sv[j - 1] = r[cur][j];
        if (a.rows_out) a.rows_out[(size_t)blk * NP + j - 1] = r[cur][j];
    }
    __syncthreads();
    if (j < NP) {
        const unsigned bits = __ballot_sync(0xffffffffu, rank_select(sv, j, a.k));
        if ((j & 31) == 0) {
            const int s = blk / a.H;
            atomicOr(&a.index[(s % a.B) * 4 + (j >> 5)], bits);
            if (a.mod_mask) atomicOr(&a.mod_mask[s * 4 + (j >> 5)], bits);
        }
    }
}

// bf16 maps: 16-byte loads.  119 threads = 17 column groups (8 keys each, 136 = pitch) x 7 row slots; each thread
// accumulates its 8 columns over rows i = slot, slot+7, ...; partial sums meet in shared memory.  ~4x more bytes in flight
// per CTA than the scalar kernel (the chain over layers is latency-bound otherwise).
__global__ void __launch_bounds__(128) rollout_topk_bf16_kernel(const RolloutArgs a) {
    __shared__ float r[132];
    __shared__ float part[7][136];
    __shared__ float sv[NP];
    const int blk = blockIdx.x;
    const int t = threadIdx.x;
    const int NT = NP + 1;
    const size_t base = (size_t)blk * a.p_rows * a.ldp;
    const __nv_bfloat16* last = reinterpret_cast<const __nv_bfloat16*>(a.maps[a.layers - 1]) + base;
    for (int j = t; j < NT; j += 128) r[j] = __bfloat162float(last[j]);
    __syncthreads();
    const int cg = t % 17, slot = t / 17;          // slot 7 (threads 119..127) idles in the load phase
    for (int l = a.layers - 2; l >= 0; --l) {
        const __nv_bfloat16* m = reinterpret_cast<const __nv_bfloat16*>(a.maps[l]) + base + cg * 8;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (slot < 7) {
            uint4 v[4];
            int i = slot;
            for (; i + 21 < NT; i += 28) {
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const uint4*>(m + (size_t)(i + 7 * u) * a.ldp);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float ri = r[i + 7 * u];
                    const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&v[u]);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float2 f = __bfloat1622float2(hh[q]);
                        acc[2 * q] += ri * f.x;
                        acc[2 * q + 1] += ri * f.y;
                    }
                }
            }
            for (; i < NT; i += 7) {
                const uint4 w = *reinterpret_cast<const uint4*>(m + (size_t)i * a.ldp);
                const float ri = r[i];
                const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&w);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float2 f = __bfloat1622float2(hh[q]);
                    acc[2 * q] += ri * f.x;
                    acc[2 * q + 1] += ri * f.y;
                }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) part[slot][cg * 8 + q] = acc[q];
        }
        __syncthreads();
        for (int j = t; j < NT; j += 128) {
            float s = 0.f;
#pragma unroll
            for (int q = 0; q < 7; ++q) s += part[q][j];
            r[j] = s;
        }
        __syncthreads();
    }
    if (t < NP) {
        sv[t] = r[t + 1];
        if (a.rows_out) a.rows_out[(size_t)blk * NP + t] = r[t + 1];
    }
    __syncthreads();
    if (t < NP) {
        const unsigned bits = __ballot_sync(0xffffffffu, rank_select(sv, t, a.k));
        if ((t & 31) == 0) {
            const int s = blk / a.H;
            atomicOr(&a.index[(s % a.B) * 4 + (t >> 5)], bits);
            if (a.mod_mask) atomicOr(&a.mod_mask[s * 4 + (t >> 5)], bits);
        }
    }
}

int rollout_topk(const void* const* maps, int layers, int maps_f32, int nseq, int B, int heads, long long p_rows,
                 long long ldp, int k, unsigned* index, unsigned* mod_mask, float* rows_out, cudaStream_t st) {
    if (nseq <= 0) return EDB_OK;
    if (layers < 1 || layers > 16) return edb_set_error(EDB_ERR_SHAPE, "rollout: 1..16 layers");
    if (p_rows < NP + 1 || ldp < NP + 1) return edb_set_error(EDB_ERR_SHAPE, "rollout: maps must be at least 129x129");
    RolloutArgs a{};
    for (int l = 0; l < layers; ++l) a.maps[l] = maps[l];
    a.layers = layers; a.H = heads; a.B = B; a.k = k; a.p_rows = p_rows; a.ldp = ldp;
    a.index = index; a.mod_mask = mod_mask; a.rows_out = rows_out;
    if (maps_f32) rollout_topk_kernel<float><<<nseq * heads, 160, 0, st>>>(a);
    else if (ldp == 136) rollout_topk_bf16_kernel<<<nseq * heads, 128, 0, st>>>(a);
    else rollout_topk_kernel<__nv_bfloat16><<<nseq * heads, 160, 0, st>>>(a);
    EDB_CHECK_LAUNCH();
    return EDB_OK;
}

// ------------------------------------------------------------------------------------------ offsets of the packed layout
// seq_off[b] = sum_{b'<b} (1 + popcount(index[b']));  seq_off3[b] = 3*seq_off[b]  (joint layout)
__global__ void index_finalize_kernel(const unsigned* __restrict__ index, int B, int* __restrict__ seq_off,
                                      int* __restrict__ seq_off3) {
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int b = 0; b < B; ++b) {
            seq_off[b] = acc;
            seq_off3[b] = 3 * acc;
            acc += 1 + __popc(index[b * 4]) + __popc(index[b * 4 + 1]) + __popc(index[b * 4 + 2]) + __popc(index[b * 4 + 3]);
        }
        seq_off[B] = acc;
        seq_off3[B] = 3 * acc;
    }
}

int index_finalize(const unsigned* index, int B, int* seq_off, int* seq_off3, cudaStream_t st) {
    index_finalize_kernel<<<1, 32, 0, st>>>(index, B, seq_off, seq_off3);
    EDB_CHECK_LAUNCH();
    return EDB_OK;
}

__device__ __forceinline__ bool token_kept(const unsigned* idx4, int t, int& rank) {
    // t = 0 is the cls token (always kept, vit_pytorch.py:310); patch p = t-1
    if (t == 0) { rank = 0; return true; }
    const int p = t - 1, w = p >> 5, bit = p & 31;
    int below = 0;
    for (int i = 0; i < w; ++i) below += __popc(idx4[i]);
    below += __popc(idx4[w] & ((1u << bit) - 1u));
    rank = 1 + below;
    return (idx4[w] >> bit) & 1u;
}

__device__ __forceinline__ float block_sum_192(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float s = 0.f;
    if (threadIdx.x == 0)
        for (int w = 0; w < 6; ++w) s += red[w];
    return s;  // valid on thread 0
}

// grid = (129, B), block = 192 (one float4 per thread).  tokens: [3][B][129][768]; packed: [3][cap][768]
__global__ void __launch_bounds__(192) sfts_pack_fwd_kernel(const float* __restrict__ tokens, const unsigned* __restrict__ index,
                                                            const int* __restrict__ seq_off, int B, long long cap,
                                                            float* __restrict__ packed, float* __restrict__ loss_bcc,
                                                            float inv_denom) {
    __shared__ float red[6];
    const int t = blockIdx.x, b = blockIdx.y, c = threadIdx.x * 4;
    const size_t mstride = (size_t)B * (NP + 1) * DT;
    const float* src = tokens + ((size_t)b * (NP + 1) + t) * DT + c;
    int rank;
    const bool kept = token_kept(index + b * 4, t, rank);
    if (kept) {
        float* dst = packed + ((size_t)seq_off[b] + rank) * DT + c;
#pragma unroll
        for (int m = 0; m < 3; ++m)
            *reinterpret_cast<float4*>(dst + (size_t)m * cap * DT) = ld4(src + m * mstride);
    } else if (loss_bcc != nullptr) {
        const float4 R = ld4(src), N = ld4(src + mstride), T = ld4(src + 2 * mstride);
        float v = 0.f;
#define SQ(a, b) ((a - b) * (a - b))
        v += SQ(R.x, N.x) + SQ(R.y, N.y) + SQ(R.z, N.z) + SQ(R.w, N.w);
        v += SQ(R.x, T.x) + SQ(R.y, T.y) + SQ(R.z, T.z) + SQ(R.w, T.w);
        v += SQ(N.x, T.x) + SQ(N.y, T.y) + SQ(N.z, T.z) + SQ(N.w, T.w);
#undef SQ
        const float s = block_sum_192(v, red);
        if (threadIdx.x == 0) atomicAdd(loss_bcc, s * inv_denom);
    }
}

int sfts_pack_fwd(const float* tokens, const unsigned* index, const int* seq_off, int B, long long cap, float* packed,
                  float* loss_bcc, cudaStream_t st) {
    if (B <= 0) return EDB_OK;
    sfts_pack_fwd_kernel<<<dim3(NP + 1, B), 192, 0, st>>>(tokens, index, seq_off, B, cap, packed, loss_bcc,
                                                            1.0f / ((float)B * NP * DT));
    EDB_CHECK_LAUNCH();
    return EDB_OK;
}

// d tokens = kept rows: d packed;  background rows: g_loss * 2/(B*128*768) * (2R - N - T) etc.  (overwrites d_tokens)
__global__ void __launch_bounds__(192) sfts_pack_bwd_kernel(const float* __restrict__ tokens, const unsigned* __restrict__ index,
                                                            const int* __restrict__ seq_off, int B, long long cap,
                                                            const float* __restrict__ d_packed,
                                                            const float* __restrict__ g_loss, float coef,
                                                            float* __restrict__ d_tokens) {
    const int t = blockIdx.x, b = blockIdx.y, c = threadIdx.x * 4;
    const size_t mstride = (size_t)B * (NP + 1) * DT;
    const size_t o = ((size_t)b * (NP + 1) + t) * DT + c;
    int rank;
    const bool kept = token_kept(index + b * 4, t, rank);
    if (kept) {
        const float* src = d_packed + ((size_t)seq_off[b] + rank) * DT + c;
#pragma unroll
        for (int m = 0; m < 3; ++m)
            *reinterpret_cast<float4*>(d_tokens + o + m * mstride) = ld4(src + (size_t)m * cap * DT);
    } else {
        float4 dR = make_float4(0.f, 0.f, 0.f, 0.f), dN = dR, dT = dR;
        if (g_loss != nullptr) {
            const float k = coef * g_loss[0];
            const float4 R = ld4(tokens + o), N = ld4(tokens + o + mstride), T = ld4(tokens + o + 2 * mstride);
            dR = make_float4(k * (2 * R.x - N.x - T.x), k * (2 * R.y - N.y - T.y), k * (2 * R.z - N.z - T.z), k * (2 * R.w - N.w - T.w));
            dN = make_float4(k * (2 * N.x - R.x - T.x), k * (2 * N.y - R.y - T.y), k * (2 * N.z - R.z - T.z), k * (2 * N.w - R.w - T.w));
            dT = make_float4(k * (2 * T.x - R.x - N.x), k * (2 * T.y - R.y - N.y), k * (2 * T.z - R.z - N.z), k * (2 * T.w - R.w - N.w));
        }
        *reinterpret_cast<float4*>(d_tokens + o) = dR;
        *reinterpret_cast<float4*>(d_tokens + o + mstride) = dN;
        *reinterpret_cast<float4*>(d_tokens + o + 2 * mstride) = dT;
    }
}

int sfts_pack_bwd(const float* tokens, const unsigned* index, const int* seq_off, int B, long long cap,
                  const float* d_packed, const float* g_loss, float* d_tokens, cudaStream_t st) {
    if (B <= 0) return EDB_OK;
    sfts_pack_bwd_kernel<<<dim3(NP + 1, B), 192, 0, st>>>(tokens, index, seq_off, B, cap, d_packed, g_loss,
                                                            2.0f / ((float)B * NP * DT), d_tokens);
    EDB_CHECK_LAUNCH();
    return EDB_OK;
}

// ------------------------------------------------------------------------------------------ per-modality <-> joint layout
// joint row 3*off_b + m*len_b + r  <->  modality m row off_b + r     (vit_pytorch.py:324 torch.cat(dim=1), packed)
// dir 0: mod -> joint, dir 1: joint -> mod.   grid = (max_len, B, 3)
__global__ void __launch_bounds__(192) joint_gather_kernel(float* __restrict__ mod, long long cap, float* __restrict__ joint,
                                                           const int* __restrict__ seq_off, int dir) {
    const int r = blockIdx.x, b = blockIdx.y, m = blockIdx.z, c = threadIdx.x * 4;
    const int off = seq_off[b], len = seq_off[b + 1] - off;
    if (r >= len) return;
    float* pm = mod + ((size_t)m * cap + off + r) * DT + c;
    float* pj = joint + ((size_t)3 * off + (size_t)m * len + r) * DT + c;
    if (dir == 0) *reinterpret_cast<float4*>(pj) = ld4(pm);
    else *reinterpret_cast<float4*>(pm) = ld4(pj);
}

int joint_gather(float* mod, long long cap, float* joint, const int* seq_off, int B, int max_len, int dir, cudaStream_t st) {
    if (B <= 0 || max_len <= 0) return EDB_OK;
    joint_gather_kernel<<<dim3(max_len, B, 3), 192, 0, st>>>(mod, cap, joint, seq_off, dir);
    EDB_CHECK_LAUNCH();
    return EDB_OK;
}

// ------------------------------------------------------------------------------------------ pooling (make_model.py:186-203)
// x: joint layout after out_norm.  cls_out/patch_mean: [3][B][768];  num[b] = #RGB patch rows with sum != 0.
// grid (B, 3): one CTA per (sample, modality); every CTA recounts the RGB rows (55 x 768 floats), four rows in flight
// per thread in the column sums
__global__ void __launch_bounds__(256) pool_fwd_kernel(const float* __restrict__ x, const int* __restrict__ seq_off, int B,
                                                       float* __restrict__ cls_out, float* __restrict__ patch_mean,
                                                       int* __restrict__ num) {
    __shared__ int cnt;
    const int b = blockIdx.x, m = blockIdx.y;
    const int off = seq_off[b], len = seq_off[b + 1] - off;
    const float* base = x + (size_t)3 * off * DT;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = 1 + warp; r < len; r += 8) {
        const float* row = base + (size_t)r * DT;
        float s = 0.f;
#pragma unroll 4
        for (int c = lane; c < DT; c += 32) s += row[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0 && s != 0.f) atomicAdd(&cnt, 1);
    }
    __syncthreads();
    const int n = cnt;
    if (threadIdx.x == 0 && m == 0) num[b] = n;
    const float inv = 1.0f / (float)n;   // n == 0 gives inf/nan exactly like the reference's division by zero
    const float* mb = base + (size_t)m * len * DT;
    for (int c = threadIdx.x; c < DT; c += 256) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int r = 1;
        for (; r + 3 < len; r += 4) {
            s0 += mb[(size_t)r * DT + c];
            s1 += mb[(size_t)(r + 1) * DT + c];
            s2 += mb[(size_t)(r + 2) * DT + c];
            s3 += mb[(size_t)(r + 3) * DT + c];
        }
        for (; r < len; ++r) s0 += mb[(size_t)r * DT + c];
        cls_out[((size_t)m * B + b) * DT + c] = mb[c];
        patch_mean[((size_t)m * B + b) * DT + c] = ((s0 + s1) + (s2 + s3)) * inv;
    }
}

__global__ void __launch_bounds__(192) pool_bwd_kernel(const float* __restrict__ d_cls, const float* __restrict__ d_patch,
                                                       const int* __restrict__ seq_off, const int* __restrict__ num, int B,
                                                       float* __restrict__ dx) {
    const int r = blockIdx.x, b = blockIdx.y, m = blockIdx.z, c = threadIdx.x * 4;
    const int off = seq_off[b], len = seq_off[b + 1] - off;
    if (r >= len) return;
    float* o = dx + ((size_t)3 * off + (size_t)m * len + r) * DT + c;
    const size_t src = ((size_t)m * B + b) * DT + c;
    if (r == 0) {
        *reinterpret_cast<float4*>(o) = ld4(d_cls + src);
    } else {
        const float inv = 1.0f / (float)num[b];
        const float4 g = ld4(d_patch + src);
        *reinterpret_cast<float4*>(o) = make_float4(g.x * inv, g.y * inv, g.z * inv, g.w * inv);
    }
}

int pool_fwd(const float* x, const int* seq_off, int B, float* cls_out, float* patch_mean, int* num, cudaStream_t st) {
    if (B <= 0) return EDB_OK;
    pool_fwd_kernel<<<dim3(B, 3), 256, 0, st>>>(x, seq_off, B, cls_out, patch_mean, num);
    EDB_CHECK_LAUNCH();
    return EDB_OK;
}

int pool_bwd(const float* d_cls, const float* d_patch, const int* seq_off, const int* num, int B, int max_len, float* dx,
             cudaStream_t st) {
    if (B <= 0) return EDB_OK;
    pool_bwd_kernel<<<dim3(max_len, B, 3), 192, 0, st>>>(d_cls, d_patch, seq_off, num, B, dx);
    EDB_CHECK_LAUNCH();
    return EDB_OK;
}

// rows[b] = x[seq_off[b]*mult + add_len*len_b]  -- gathers / scatters the cls rows of a packed matrix (OCFR input, vit_pytorch.py:319-323)
__global__ void __launch_bounds__(192) cls_rows_kernel(float* __restrict__ packed, long long cap, const int* __restrict__ seq_off,
                                                       int B, float* __restrict__ rows, int dir) {
    const int b = blockIdx.x, m = blockIdx.y, c = threadIdx.x * 4;
    float* p = packed + ((size_t)m * cap + seq_off[b]) * DT + c;
    float* r = rows + ((size_t)m * B + b) * DT + c;
    if (dir == 0) *reinterpret_cast<float4*>(r) = ld4(p);
    else {  // accumulate gradient into the packed rows
        const float4 g = ld4(r), o = ld4(p);
        *reinterpret_cast<float4*>(p) = make_float4(o.x + g.x, o.y + g.y, o.z + g.z, o.w + g.w);
    }
}

int cls_rows(float* packed, long long cap, const int* seq_off, int B, float* rows, int dir, cudaStream_t st) {
    if (B <= 0) return EDB_OK;
    cls_rows_kernel<<<dim3(B, 3), 192, 0, st>>>(packed, cap, seq_off, B, rows, dir);
    EDB_CHECK_LAUNCH();
    return EDB_OK;
}

}  // namespace edb

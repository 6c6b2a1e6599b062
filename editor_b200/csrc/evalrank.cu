// Retrieval evaluation on the GPU: the step after the eval forward in the reference's
// do_train / do_inference loops -- R1_mAP_eval.compute / R1_mAP.compute (utils/metrics.py:209-283):
//   F.normalize(feats)                      (:255-256)  -> eval_normalize_kernel   (HBM-bound, one warp per row)
//   euclidean_distance(qf, gf)              (:12-18)    -> eval_distmat_kernel     (fp32 FFMA tiles; the row norms are
//                                                          accumulated from the same shared-memory tiles)
//   eval_func / eval_func_msrv              (:133-191 / :36-130) -> eval_rank_kernel: one CTA per query, NO argsort.
// The reference sorts every distance row (np.argsort) and walks it in Python.  CMC and AP only need the rank of each
// correct match among the kept gallery items:  rank_j = 1 + #{kept k : d_k < d_j or (d_k == d_j and k < j)}  and the
// number c_j of correct matches at or before it;  AP = (1/R) sum_j c_j / rank_j,  first_rank = min_j rank_j  (integer
// logic; AP in fp64 like numpy).  Equal distances rank by ascending gallery index (np.argsort(kind="stable")).
#include "abi_internal.h"

namespace edb {

__global__ void __launch_bounds__(256) eval_normalize_kernel(float* __restrict__ x, long long ld, int N, int F, float eps) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= N) return;
    float* p = x + (size_t)row * ld;
    float s = 0.f;
    for (int c = lane; c < F; c += 32) s = fmaf(p[c], p[c], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float inv = 1.0f / fmaxf(sqrtf(s), eps);
    for (int c = lane; c < F; c += 32) p[c] *= inv;
}

// dist[q][g] = |q|^2 + |g|^2 - 2 q.g ; 64 x 64 tile per CTA, 16-deep k slices, 4 x 4 outputs per thread
constexpr int ED_T = 64, ED_K = 16;
__global__ void __launch_bounds__(256) eval_distmat_kernel(const float* __restrict__ qf, long long ldq, int Q,
                                                           const float* __restrict__ gf, long long ldg, int G, int F,
                                                           float* __restrict__ dist, long long ldd) {
    __shared__ float sa[ED_K][ED_T + 4], sb[ED_K][ED_T + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int q0 = blockIdx.y * ED_T, g0 = blockIdx.x * ED_T;
    float acc[4][4] = {}, na[4] = {}, nb[4] = {};
    const int lr = threadIdx.x >> 2, lk = (threadIdx.x & 3) * 4;       // loader: row 0..63, k offset 0,4,8,12
    for (int k0 = 0; k0 < F; k0 += ED_K) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int k = k0 + lk + t;
            sa[lk + t][lr] = (q0 + lr < Q && k < F) ? qf[(size_t)(q0 + lr) * ldq + k] : 0.f;
            sb[lk + t][lr] = (g0 + lr < G && k < F) ? gf[(size_t)(g0 + lr) * ldg + k] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < ED_K; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = sa[k][ty * 4 + i]; b[i] = sb[k][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                na[i] = fmaf(a[i], a[i], na[i]);
                nb[i] = fmaf(b[i], b[i], nb[i]);
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int q = q0 + ty * 4 + i;
        if (q >= Q) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int g = g0 + tx * 4 + j;
            if (g < G) dist[(size_t)q * ldd + g] = fmaf(-2.0f, acc[i][j], na[i] + nb[j]);
        }
    }
}

constexpr int ER_MAXM = 2048;   // correct matches per query held in shared memory
constexpr int ER_TILE = 2048;   // gallery slice staged in shared memory

__global__ void __launch_bounds__(256) eval_rank_kernel(const float* __restrict__ dist, long long ldd, int G,
                                                        const long long* __restrict__ q_pid, const long long* __restrict__ g_pid,
                                                        const long long* __restrict__ q_key, const long long* __restrict__ g_key,
                                                        double* __restrict__ ap, int* __restrict__ first_rank,
                                                        int* __restrict__ overflow) {
    __shared__ float m_d[ER_MAXM];
    __shared__ int m_i[ER_MAXM];
    __shared__ float t_d[ER_TILE];
    __shared__ unsigned char t_keep[ER_TILE];
    __shared__ int warp_cnt[8];
    __shared__ int s_total;
    __shared__ double red[256];
    __shared__ int redi[256];
    const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* drow = dist + (size_t)q * ldd;
    const long long pid = q_pid[q], key = q_key[q];
    if (tid == 0) s_total = 0;
    __syncthreads();
    // ---- ordered compaction of the correct matches (same pid, not removed) in ascending gallery index
    for (int base = 0; base < G; base += 256) {
        const int k = base + tid;
        bool is_m = false;
        if (k < G) {
            const bool same = g_pid[k] == pid;
            is_m = same && !(g_key[k] == key);          // removed: same pid AND same camera (scene); correct: same pid, kept
        }
        const unsigned bal = __ballot_sync(0xffffffffu, is_m);
        if (lane == 0) warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int off = s_total;
        for (int w = 0; w < warp; ++w) off += warp_cnt[w];
        if (is_m) {
            const int pos = off + __popc(bal & ((1u << lane) - 1u));
            if (pos < ER_MAXM) { m_d[pos] = drow[k]; m_i[pos] = k; }
        }
        __syncthreads();
        if (tid == 0) {
            int t = s_total;
            for (int w = 0; w < 8; ++w) t += warp_cnt[w];
            s_total = t;
        }
        __syncthreads();
    }
    const int R = s_total;
    if (R == 0) {                                       // query identity absent from the gallery (:165-167)
        if (tid == 0) { ap[q] = 0.0; first_rank[q] = -1; }
        return;
    }
    if (R > ER_MAXM) {
        if (tid == 0) { ap[q] = 0.0; first_rank[q] = -1; atomicAdd(overflow, 1); }
        return;
    }
    double ap_part = 0.0;
    int best = 0x7fffffff;
    for (int mb = 0; mb < R; mb += 256) {               // 256 matches per pass, one per thread
        const int j = mb + tid;
        const bool active = j < R;
        const float dj = active ? m_d[j] : 0.f;
        const int ij = active ? m_i[j] : 0;
        int rank = 1;
        for (int base = 0; base < G; base += ER_TILE) {
            __syncthreads();
            for (int t = tid; t < ER_TILE && base + t < G; t += 256) {
                const int k = base + t;
                t_d[t] = drow[k];
                t_keep[t] = !(g_pid[k] == pid && g_key[k] == key);
            }
            __syncthreads();
            const int n = min(ER_TILE, G - base);
            if (active) {
                for (int t = 0; t < n; ++t) {
                    const float dk = t_d[t];
                    rank += (t_keep[t] && (dk < dj || (dk == dj && base + t < ij))) ? 1 : 0;
                }
            }
        }
        if (active) {
            int c = 1;
            for (int t = 0; t < R; ++t) {
                const float dk = m_d[t];
                c += (dk < dj || (dk == dj && m_i[t] < ij)) ? 1 : 0;
            }
            ap_part += (double)c / (double)rank;
            best = min(best, rank);
        }
    }
    red[tid] = ap_part;
    redi[tid] = best;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {                 // fixed-order tree: deterministic fp64 sum
        if (tid < s) { red[tid] += red[tid + s]; redi[tid] = min(redi[tid], redi[tid + s]); }
        __syncthreads();
    }
    if (tid == 0) { ap[q] = red[0] / (double)R; first_rank[q] = redi[0]; }
}

int eval_normalize(float* feats, long long ld, int N, int F, float eps, cudaStream_t st) {
    if (N <= 0 || F <= 0) return EDB_OK;
    eval_normalize_kernel<<<(N + 7) / 8, 256, 0, st>>>(feats, ld, N, F, eps);
    EDB_CHECK_LAUNCH();
    return EDB_OK;
}

int eval_distmat(const float* qf, long long ldq, int Q, const float* gf, long long ldg, int G, int F, float* dist,
                 long long ldd, cudaStream_t st) {
    if (Q <= 0 || G <= 0) return EDB_OK;
    if (F <= 0) return edb_set_error(EDB_ERR_SHAPE, "eval_distmat: empty feature dimension");
    dim3 grid((G + ED_T - 1) / ED_T, (Q + ED_T - 1) / ED_T);
    if (grid.y > 65535) return edb_set_error(EDB_ERR_SHAPE, "eval_distmat: more than 4 M queries");
    eval_distmat_kernel<<<grid, 256, 0, st>>>(qf, ldq, Q, gf, ldg, G, F, dist, ldd);
    EDB_CHECK_LAUNCH();
    return EDB_OK;
}

int eval_rank(const float* dist, long long ldd, int Q, int G, const long long* q_pid, const long long* g_pid,
              const long long* q_key, const long long* g_key, double* ap, int* first_rank, int* overflow, cudaStream_t st) {
    if (Q <= 0) return EDB_OK;
    if (G <= 0) return edb_set_error(EDB_ERR_SHAPE, "eval_rank: empty gallery");
    eval_rank_kernel<<<Q, 256, 0, st>>>(dist, ldd, G, q_pid, g_pid, q_key, g_key, ap, first_rank, overflow);
    EDB_CHECK_LAUNCH();
    return EDB_OK;
}

}  // namespace edb

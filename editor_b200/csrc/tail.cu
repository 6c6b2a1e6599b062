// The [B, 768..2304] tail of EDITOR.forward and the loss that follows it -- tiny tensors, one or two CTAs of work each,
// written as plain CUDA so that no ATen kernel remains on the training step:
//   bn1d_fwd / bn1d_bwd      nn.BatchNorm1d in training mode (make_model.py:114-141, BNNeck), running-stat update included
//   ocfr_*                   OCFR.forward/update/compute_intra_loss (fusion_part/OCFR.py:22-84) for P x K contiguous labels
//   ce_smooth                CrossEntropyLabelSmooth (layers/softmax_loss.py:23-34), loss and d(logits)
//   pairdist / triplet_*     euclidean_dist + hard_example_mining + SoftMarginLoss (layers/triplet_loss.py:16-31,51-105,122-136)
#include "abi_internal.h"

namespace edb {

__device__ __forceinline__ float warp_sum_t(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------ BatchNorm1d (training)
// grid = F/32, block = (32, 8): thread (fx, ry) owns feature blockIdx.x*32+fx and rows ry, ry+8, ...
__global__ void __launch_bounds__(256) bn1d_fwd_kernel(const float* __restrict__ x, long long ldx, int B, int F,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       float* __restrict__ run_mean, float* __restrict__ run_var,
                                                       float momentum, float eps, float* __restrict__ y, long long ldy,
                                                       float* __restrict__ save_mean, float* __restrict__ save_invstd) {
    __shared__ float red[8][33];
    const int fx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int f = blockIdx.x * 32 + fx;
    float s = 0.f;
    for (int r = ry; r < B; r += 8) s += x[(size_t)r * ldx + f];
    red[ry][fx] = s;
    __syncthreads();
    float mean = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) mean += red[k][fx];
    mean /= (float)B;
    __syncthreads();
    float q = 0.f;
    for (int r = ry; r < B; r += 8) {
        const float d = x[(size_t)r * ldx + f] - mean;
        q += d * d;
    }
    red[ry][fx] = q;
    __syncthreads();
    float var = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) var += red[k][fx];
    var /= (float)B;
    const float invstd = rsqrtf(var + eps);
    const float g = gamma[f], b = beta[f];
    for (int r = ry; r < B; r += 8) y[(size_t)r * ldy + f] = (x[(size_t)r * ldx + f] - mean) * invstd * g + b;
    if (ry == 0) {
        save_mean[f] = mean;
        save_invstd[f] = invstd;
        const float unbiased = B > 1 ? var * (float)B / (float)(B - 1) : var;
        run_mean[f] = (1.f - momentum) * run_mean[f] + momentum * mean;
        run_var[f] = (1.f - momentum) * run_var[f] + momentum * unbiased;
    }
}

__global__ void __launch_bounds__(256) bn1d_bwd_kernel(const float* __restrict__ dy, long long lddy, const float* __restrict__ x,
                                                       long long ldx, int B, int F, const float* __restrict__ gamma,
                                                       const float* __restrict__ save_mean, const float* __restrict__ save_invstd,
                                                       float* __restrict__ dx, long long lddx, float* __restrict__ dgamma,
                                                       float* __restrict__ dbeta) {
    __shared__ float red1[8][33], red2[8][33];
    const int fx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int f = blockIdx.x * 32 + fx;
    const float mean = save_mean[f], invstd = save_invstd[f];
    float s1 = 0.f, s2 = 0.f;
    for (int r = ry; r < B; r += 8) {
        const float d = dy[(size_t)r * lddy + f];
        s1 += d;
        s2 += d * (x[(size_t)r * ldx + f] - mean) * invstd;
    }
    red1[ry][fx] = s1;
    red2[ry][fx] = s2;
    __syncthreads();
    float db = 0.f, dg = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { db += red1[k][fx]; dg += red2[k][fx]; }
    const float g = gamma[f], invB = 1.0f / (float)B;
    for (int r = ry; r < B; r += 8) {
        const float xh = (x[(size_t)r * ldx + f] - mean) * invstd;
        dx[(size_t)r * lddx + f] = g * invstd * (dy[(size_t)r * lddy + f] - db * invB - xh * dg * invB);
    }
    if (ry == 0) {
        dgamma[f] += dg;
        dbeta[f] += db;
    }
}

int bn1d_fwd(const float* x, long long ldx, int B, int F, const float* gamma, const float* beta, float* run_mean,
             float* run_var, float momentum, float eps, float* y, long long ldy, float* save_mean, float* save_invstd,
             cudaStream_t st) {
    if (B <= 0) return EDB_OK;
    if (F % 32) return edb_set_error(EDB_ERR_SHAPE, "bn1d: feature count must be a multiple of 32");
    bn1d_fwd_kernel<<<F / 32, 256, 0, st>>>(x, ldx, B, F, gamma, beta, run_mean, run_var, momentum, eps, y, ldy, save_mean,
                                            save_invstd);
    EDB_CHECK_LAUNCH();
    return EDB_OK;
}

int bn1d_bwd(const float* dy, long long lddy, const float* x, long long ldx, int B, int F, const float* gamma,
             const float* save_mean, const float* save_invstd, float* dx, long long lddx, float* dgamma, float* dbeta,
             cudaStream_t st) {
    if (B <= 0) return EDB_OK;
    if (F % 32) return edb_set_error(EDB_ERR_SHAPE, "bn1d: feature count must be a multiple of 32");
    bn1d_bwd_kernel<<<F / 32, 256, 0, st>>>(dy, lddy, x, ldx, B, F, gamma, save_mean, save_invstd, dx, lddx, dgamma, dbeta);
    EDB_CHECK_LAUNCH();
    return EDB_OK;
}

// ------------------------------------------------------------------------------------------ OCFR
constexpr int OD = 768;
struct OcfrPtrs { float* c[3]; };

// grid (B, 3): fn = x / max(||x||, 1e-12)   (F.normalize, OCFR.py:46-49)
__global__ void __launch_bounds__(192) ocfr_norm_kernel(const float* __restrict__ x, int B, float* __restrict__ fn,
                                                        float* __restrict__ inv_norm) {
    __shared__ float red[6];
    const int b = blockIdx.x, m = blockIdx.y, c = threadIdx.x * 4;
    const size_t o = ((size_t)m * B + b) * OD + c;
    const float4 v = *reinterpret_cast<const float4*>(x + o);
    float s = warp_sum_t(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < 6; ++w) tot += red[w];
    const float inv = 1.0f / fmaxf(sqrtf(tot), 1e-12f);
    *reinterpret_cast<float4*>(fn + o) = make_float4(v.x * inv, v.y * inv, v.z * inv, v.w * inv);
    if (threadIdx.x == 0) inv_norm[m * B + b] = inv;
}

// grid (C, 3): centres[c] = mom * mean_{b: label_b = c} fn_b + (1 - mom) * centres[c]  for the ids present (OCFR.py:22-29,71-84)
__global__ void __launch_bounds__(192) ocfr_update_kernel(const float* __restrict__ fn, const long long* __restrict__ label,
                                                          int B, OcfrPtrs cen, float mom) {
    extern __shared__ int lab[];
    const int cls = blockIdx.x, m = blockIdx.y, c = threadIdx.x * 4;
    for (int i = threadIdx.x; i < B; i += 192) lab[i] = (int)label[i];
    __syncthreads();
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int cnt = 0;
    for (int b = 0; b < B; ++b) {
        if (lab[b] != cls) continue;
        const float4 v = *reinterpret_cast<const float4*>(fn + ((size_t)m * B + b) * OD + c);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        ++cnt;
    }
    if (cnt == 0) return;
    const float inv = 1.0f / (float)cnt;
    float* p = cen.c[m] + (size_t)cls * OD + c;
    const float4 o = *reinterpret_cast<const float4*>(p);
    *reinterpret_cast<float4*>(p) = make_float4(mom * acc.x * inv + (1.f - mom) * o.x, mom * acc.y * inv + (1.f - mom) * o.y,
                                                mom * acc.z * inv + (1.f - mom) * o.z, mom * acc.w * inv + (1.f - mom) * o.w);
}

// grid (B, 3): loss += sum (centre[label_b] - fn_b)^2 / (B*768)   (three nn.MSELoss, OCFR.py:31-42,55-57)
// with g_loss != nullptr instead writes dx = (dfn - fn (fn . dfn)) * inv_norm, dfn = g * 2 (fn - c) / (B*768)
__global__ void __launch_bounds__(192) ocfr_loss_kernel(const float* __restrict__ fn, const float* __restrict__ inv_norm,
                                                        const long long* __restrict__ label, int B, OcfrPtrs cen,
                                                        float* __restrict__ loss, const float* __restrict__ g_loss,
                                                        float* __restrict__ dx) {
    __shared__ float red[6];
    const int b = blockIdx.x, m = blockIdx.y, c = threadIdx.x * 4;
    const size_t o = ((size_t)m * B + b) * OD + c;
    const float4 f = *reinterpret_cast<const float4*>(fn + o);
    const float4 ce = *reinterpret_cast<const float4*>(cen.c[m] + (size_t)label[b] * OD + c);
    const float4 d = make_float4(f.x - ce.x, f.y - ce.y, f.z - ce.z, f.w - ce.w);
    const float denom = 1.0f / ((float)B * OD);
    float part = g_loss == nullptr ? (d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w)
                                   : (f.x * d.x + f.y * d.y + f.z * d.z + f.w * d.w);
    part = warp_sum_t(part);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < 6; ++w) tot += red[w];
    if (g_loss == nullptr) {
        if (threadIdx.x == 0) atomicAdd(loss, tot * denom);
    } else {
        const float k = g_loss[0] * 2.0f * denom, inv = inv_norm[m * B + b];
        // dfn = k*d ;  fn . dfn = k * tot
        *reinterpret_cast<float4*>(dx + o) = make_float4(k * (d.x - f.x * tot) * inv, k * (d.y - f.y * tot) * inv,
                                                         k * (d.z - f.z * tot) * inv, k * (d.w - f.w * tot) * inv);
    }
}

int ocfr_fwd(const float* x, const long long* label, int B, int C, float* c0, float* c1, float* c2, float mom, float* fn,
             float* inv_norm, float* loss, cudaStream_t st) {
    if (B <= 0) return EDB_OK;
    OcfrPtrs cp{{c0, c1, c2}};
    ocfr_norm_kernel<<<dim3(B, 3), 192, 0, st>>>(x, B, fn, inv_norm);
    EDB_CHECK_LAUNCH();
    ocfr_update_kernel<<<dim3(C, 3), 192, B * sizeof(int), st>>>(fn, label, B, cp, mom);
    EDB_CHECK_LAUNCH();
    ocfr_loss_kernel<<<dim3(B, 3), 192, 0, st>>>(fn, inv_norm, label, B, cp, loss, nullptr, nullptr);
    EDB_CHECK_LAUNCH();
    return EDB_OK;
}

int ocfr_bwd(const float* fn, const float* inv_norm, const long long* label, int B, float* c0, float* c1, float* c2,
             const float* g_loss, float* dx, cudaStream_t st) {
    if (B <= 0) return EDB_OK;
    OcfrPtrs cp{{c0, c1, c2}};
    ocfr_loss_kernel<<<dim3(B, 3), 192, 0, st>>>(fn, inv_norm, label, B, cp, nullptr, g_loss, dx);
    EDB_CHECK_LAUNCH();
    return EDB_OK;
}

// ------------------------------------------------------------------------------------------ label-smoothed cross entropy
// one warp per row: loss += (1/B) sum_c -t_c log p_c, t = (1-eps) onehot + eps/C;  dlogits = (p - t) / B
__global__ void __launch_bounds__(256) ce_smooth_kernel(const float* __restrict__ logits, long long ld, const long long* __restrict__ label,
                                                        int B, int C, float eps, float* __restrict__ loss,
                                                        float* __restrict__ dlogits, long long ldd) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= B) return;
    const float* z = logits + (size_t)row * ld;
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, z[c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float se = 0.f, sz = 0.f;
    for (int c = lane; c < C; c += 32) {
        se += expf(z[c] - mx);
        sz += z[c];
    }
    se = warp_sum_t(se);
    sz = warp_sum_t(sz);
    const float lse = mx + logf(se);
    const int y = (int)label[row];
    const float invB = 1.0f / (float)B;
    if (lane == 0) {
        // -sum_c t_c (z_c - lse) = lse - (1-eps) z_y - (eps/C) sum_c z_c
        const float l = lse - (1.f - eps) * z[y] - (eps / (float)C) * sz;
        atomicAdd(loss, l * invB);
    }
    if (dlogits != nullptr) {
        for (int c = lane; c < C; c += 32) {
            const float p = expf(z[c] - lse);
            const float t = (c == y ? (1.f - eps) : 0.f) + eps / (float)C;
            dlogits[(size_t)row * ldd + c] = (p - t) * invB;
        }
    }
}

int ce_smooth(const float* logits, long long ld, const long long* label, int B, int C, float eps, float* loss,
              float* dlogits, long long ldd, cudaStream_t st) {
    if (B <= 0) return EDB_OK;
    ce_smooth_kernel<<<(B + 7) / 8, 256, 0, st>>>(logits, ld, label, B, C, eps, loss, dlogits, ldd);
    EDB_CHECK_LAUNCH();
    return EDB_OK;
}

// ------------------------------------------------------------------------------------------ batch-hard soft-margin triplet
// Gram matrix in feature slices: part[z][i][j] = sum_{f in slice z} x_i[f] x_j[f]  (fp32).  The feature dimension is cut
// into TR_SLICES slices (one CTA each per 16x16 tile) so that the 64 tiles of a 128-batch become 512 CTAs with short
// dependent-load chains; the slices are added in a fixed order by the mining kernel (deterministic).
constexpr int TR_SLICES = 8;
__global__ void __launch_bounds__(256) pairdist_kernel(const float* __restrict__ x, long long ld, int B, int F,
                                                       float* __restrict__ part) {
    __shared__ float xi[16][65], xj[16][65];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int i = blockIdx.y * 16 + ty, j = blockIdx.x * 16 + tx;
    const int per = ((F + TR_SLICES - 1) / TR_SLICES + 63) / 64 * 64;
    const int fa = blockIdx.z * per, fb = min(F, fa + per);
    float dot = 0.f;
    for (int f0 = fa; f0 < fb; f0 += 64) {
        for (int t = threadIdx.x; t < 16 * 64; t += 256) {
            const int r = t >> 6, c = t & 63;
            const int gi = blockIdx.y * 16 + r, gj = blockIdx.x * 16 + r;
            xi[r][c] = (gi < B && f0 + c < fb) ? x[(size_t)gi * ld + f0 + c] : 0.f;
            xj[r][c] = (gj < B && f0 + c < fb) ? x[(size_t)gj * ld + f0 + c] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < 64; ++c) dot += xi[ty][c] * xj[tx][c];
        __syncthreads();
    }
    if (i < B && j < B) part[((size_t)blockIdx.z * B + i) * B + j] = dot;
}

// one warp per anchor i: G = sum of the slices; dist[i][j] = sqrt(max(G_ii + G_jj - 2 G_ij, 1e-12)) (euclidean_dist,
// triplet_loss.py:16-31); hardest positive / negative (first index on ties, like torch.max / torch.min);
// loss = mean log(1 + exp(-(d_an - d_ap))), and the two gradient coefficients  cp = dL/d(d_ap) / d_ap,
// cn = dL/d(d_an) / d_an   (0 where the clamp is active)
__global__ void __launch_bounds__(128) triplet_mine_kernel(const float* __restrict__ part, const long long* __restrict__ label,
                                                           int B, float* __restrict__ loss, int* __restrict__ pidx,
                                                           int* __restrict__ nidx, float* __restrict__ cp,
                                                           float* __restrict__ cn) {
    const int i = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= B) return;
    auto gram = [&](int a, int b) {
        float g = 0.f;
#pragma unroll
        for (int z = 0; z < TR_SLICES; ++z) g += part[((size_t)z * B + a) * B + b];
        return g;
    };
    const long long li = label[i];
    const float gii = gram(i, i);
    float ap = -INFINITY, an = INFINITY;
    int p = B, n = B;
    for (int j = lane; j < B; j += 32) {
        const float d = sqrtf(fmaxf(gii + gram(j, j) - 2.f * gram(i, j), 1e-12f));
        if (label[j] == li) { if (d > ap) { ap = d; p = j; } }
        else if (d < an) { an = d; n = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ap2 = __shfl_xor_sync(0xffffffffu, ap, o), an2 = __shfl_xor_sync(0xffffffffu, an, o);
        const int p2 = __shfl_xor_sync(0xffffffffu, p, o), n2 = __shfl_xor_sync(0xffffffffu, n, o);
        if (ap2 > ap || (ap2 == ap && p2 < p)) { ap = ap2; p = p2; }
        if (an2 < an || (an2 == an && n2 < n)) { an = an2; n = n2; }
    }
    if (lane != 0) return;
    if (p >= B) p = i;
    if (n >= B) n = i;
    const float xv = an - ap;
    const float l = xv > 0.f ? log1pf(expf(-xv)) : -xv + log1pf(expf(xv));
    atomicAdd(loss, l / (float)B);
    const float sig = 1.0f / (1.0f + expf(xv));          // sigma(-x) = -dl/dx
    const float dan = -sig / (float)B, dap = sig / (float)B;
    pidx[i] = p; nidx[i] = n;
    cp[i] = ap > 1.0e-6f ? dap / ap : 0.f;
    cn[i] = an > 1.0e-6f ? dan / an : 0.f;
}

// grid (B, ceil(F/256)): dx_i = g * [ cp_i (x_i - x_p) + cn_i (x_i - x_n) + sum_{k: p_k = i} cp_k (x_i - x_k)
//                                     + sum_{k: n_k = i} cn_k (x_i - x_k) ]
__global__ void __launch_bounds__(256) triplet_grad_kernel(const float* __restrict__ x, long long ld, int B, int F,
                                                           const int* __restrict__ pidx, const int* __restrict__ nidx,
                                                           const float* __restrict__ cp, const float* __restrict__ cn,
                                                           const float* __restrict__ g, float* __restrict__ dx, long long ldd,
                                                           int accumulate) {
    extern __shared__ int sidx[];      // [2B] ints then [2B] floats
    int* sp = sidx;
    int* sn = sidx + B;
    float* scp = reinterpret_cast<float*>(sidx + 2 * B);
    float* scn = scp + B;
    for (int k = threadIdx.x; k < B; k += 256) { sp[k] = pidx[k]; sn[k] = nidx[k]; scp[k] = cp[k]; scn[k] = cn[k]; }
    __syncthreads();
    const int i = blockIdx.x;
    const float gs = g[0];
    const int f = blockIdx.y * 256 + threadIdx.x;
    if (f >= F) return;
    const float xi = x[(size_t)i * ld + f];
    float acc = scp[i] * (xi - x[(size_t)sp[i] * ld + f]) + scn[i] * (xi - x[(size_t)sn[i] * ld + f]);
    for (int k = 0; k < B; ++k) {
        if (sp[k] == i) acc += scp[k] * (xi - x[(size_t)k * ld + f]);
        if (sn[k] == i) acc += scn[k] * (xi - x[(size_t)k * ld + f]);
    }
    float* o = dx + (size_t)i * ldd + f;
    *o = accumulate ? *o + gs * acc : gs * acc;
}

size_t triplet_workspace_bytes(int B) { return (size_t)TR_SLICES * B * B * 4 + (size_t)B * 16; }

int triplet_fwd(const float* x, long long ld, const long long* label, int B, int F, float* loss, void* workspace,
                size_t ws_bytes, cudaStream_t st) {
    if (B <= 0) return EDB_OK;
    if (ws_bytes < triplet_workspace_bytes(B)) return edb_set_error(EDB_ERR_WORKSPACE, "triplet: workspace too small");
    float* dist = static_cast<float*>(workspace);
    int* pidx = reinterpret_cast<int*>(dist + (size_t)TR_SLICES * B * B);
    int* nidx = pidx + B;
    float* cp = reinterpret_cast<float*>(nidx + B);
    float* cn = cp + B;
    pairdist_kernel<<<dim3((B + 15) / 16, (B + 15) / 16, TR_SLICES), 256, 0, st>>>(x, ld, B, F, dist);
    EDB_CHECK_LAUNCH();
    triplet_mine_kernel<<<(B + 3) / 4, 128, 0, st>>>(dist, label, B, loss, pidx, nidx, cp, cn);
    EDB_CHECK_LAUNCH();
    return EDB_OK;
}

int triplet_bwd(const float* x, long long ld, int B, int F, const void* workspace, const float* g, float* dx,
                long long ldd, int accumulate, cudaStream_t st) {
    if (B <= 0) return EDB_OK;
    const float* dist = static_cast<const float*>(workspace);
    const int* pidx = reinterpret_cast<const int*>(dist + (size_t)TR_SLICES * B * B);
    const int* nidx = pidx + B;
    const float* cp = reinterpret_cast<const float*>(nidx + B);
    const float* cn = cp + B;
    triplet_grad_kernel<<<dim3(B, (F + 255) / 256), 256, (size_t)B * 16, st>>>(x, ld, B, F, pidx, nidx, cp, cn, g, dx, ldd,
                                                                             accumulate);
    EDB_CHECK_LAUNCH();
    return EDB_OK;
}

// y = a * x (+ y)  over n floats -- scales the pre-computed d(logits) by the upstream gradient
__global__ void scale_kernel(const float* __restrict__ x, const float* __restrict__ a, float* __restrict__ y, size_t n) {
    const float s = a[0];
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] = s * x[i];
}

int scale_by(const float* x, const float* a, float* y, size_t n, cudaStream_t st) {
    if (n == 0) return EDB_OK;
    size_t blocks = (n + 255) / 256;
    if (blocks > 1184) blocks = 1184;
    scale_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, a, y, n);
    EDB_CHECK_LAUNCH();
    return EDB_OK;
}

}  // namespace edb

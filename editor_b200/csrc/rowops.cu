// HBM-bound row kernels of the EDITOR hot path: LayerNorm forward/backward (768-wide rows, one warp per row),
// bias-gradient column sums, fp32->bf16 casts / error-compensated bf16 splits, patch im2col and the token-embedding
// assembly (cls + pos + SIE) with its backward.  Reference: modeling/backbones/vit_pytorch.py:206-220 (LN in Block),
// :265 (LN eps 1e-5 in BlockMask), :455-457 (patch conv), :627-637 (cls/pos/SIE).
#include "abi_internal.h"

namespace edb {

constexpr int D = 768;          // token width of ViT-B/16 (the only width on the path)
constexpr int VPL = D / 128;    // float4 vectors per lane (6)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void store4(float* p, float a, float b, float c, float d) {
    *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
__device__ __forceinline__ void store4(__nv_bfloat16* p, float a, float b, float c, float d) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&lo);
    u.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(p) = u;
}
__device__ __forceinline__ float4 load4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 load4(const __nv_bfloat16* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
}

// ------------------------------------------------------------------------------------------ LayerNorm forward
template <typename OutT>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ x, long long ldx,
                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                     float eps, OutT* __restrict__ y, long long ldy,
                                                     float* __restrict__ mean, float* __restrict__ rstd, int rows,
                                                     const int* __restrict__ rows_dev) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (rows_dev != nullptr) rows = min(rows, *rows_dev);
    if (row >= rows) return;
    const float* xr = x + (size_t)row * ldx;
    float4 v[VPL];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
        v[j] = load4(xr + (j * 32 + lane) * 4);
        s += v[j].x + v[j].y + v[j].z + v[j].w;
    }
    const float mu = warp_sum(s) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
        const float a = v[j].x - mu, b = v[j].y - mu, c = v[j].z - mu, d = v[j].w - mu;
        q += a * a + b * b + c * c + d * d;
    }
    const float rs = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
    OutT* yr = y + (size_t)row * ldy;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
        const int c = (j * 32 + lane) * 4;
        const float4 g = load4(gamma + c), b = load4(beta + c);
        store4(yr + c, (v[j].x - mu) * rs * g.x + b.x, (v[j].y - mu) * rs * g.y + b.y, (v[j].z - mu) * rs * g.z + b.z,
               (v[j].w - mu) * rs * g.w + b.w);
    }
    if (lane == 0) {
        if (mean) mean[row] = mu;
        if (rstd) rstd[row] = rs;
    }
}

int layernorm_fwd(const float* x, long long ldx, const float* gamma, const float* beta, float eps, void* y,
                  long long ldy, int y_f32, float* mean, float* rstd, int rows, int dim, const int* rows_dev,
                  cudaStream_t st) {
    if (dim != D) return edb_set_error(EDB_ERR_SHAPE, "layernorm: only 768-wide rows are supported");
    if (rows <= 0) return EDB_OK;
    if (ldx % 4 || ldy % 4) return edb_set_error(EDB_ERR_ALIGN, "layernorm: row pitch must be a multiple of 4");
    const int grid = (rows + 7) / 8;
    if (y_f32)
        ln_fwd_kernel<float><<<grid, 256, 0, st>>>(x, ldx, gamma, beta, eps, (float*)y, ldy, mean, rstd, rows, rows_dev);
    else
        ln_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(x, ldx, gamma, beta, eps, (__nv_bfloat16*)y, ldy, mean, rstd,
                                                          rows, rows_dev);
    EDB_CHECK_LAUNCH();
    return EDB_OK;
}

// ------------------------------------------------------------------------------------------ LayerNorm backward
// g_out = g_in + dLN(dy);  partial[cta][0:768) = sum dy*xhat (dgamma), [768:1536) = sum dy (dbeta),
// [1536:2304) = sum g_out (bias gradient of the Linear that produced this residual-stream value).
constexpr int LNB_WARPS = 8;

template <typename DyT>
__global__ void __launch_bounds__(LNB_WARPS * 32)
ln_bwd_kernel(const DyT* __restrict__ dy, long long lddy, const float* __restrict__ x, long long ldx,
              const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
              const float* __restrict__ g_in, float* __restrict__ g_out, long long ldg,
              __nv_bfloat16* __restrict__ g_bf16, long long ldgb, float* __restrict__ dgamma, float* __restrict__ dbeta,
              float* __restrict__ dcol, int rows, const float* __restrict__ row_scale, int scale_group,
              const int* __restrict__ rows_dev) {
    __shared__ float red[LNB_WARPS][D];
    if (rows_dev != nullptr) rows = min(rows, *rows_dev);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 gm[VPL];
    float4 acc_g[VPL], acc_b[VPL], acc_c[VPL];
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
        gm[j] = load4(gamma + (j * 32 + lane) * 4);
        acc_g[j] = acc_b[j] = acc_c[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int row = blockIdx.x * LNB_WARPS + warp; row < rows; row += gridDim.x * LNB_WARPS) {
        const float mu = mean[row], rs = rstd[row];
        const float sc = row_scale ? row_scale[row / scale_group] : 1.0f;
        const float* xr = x + (size_t)row * ldx;
        const DyT* dr = dy + (size_t)row * lddy;
        float4 xh[VPL], d[VPL], gi[VPL];
        float s1 = 0.f, s2 = 0.f;
        // the incoming residual gradient is requested together with x and dy: 18 instead of 12 independent 16-byte loads
        // in flight per lane (the kernel runs one 8-warp CTA per SM and is bound by bytes in flight: 0.74 of the HBM
        // copy bandwidth in profiles/r01_ncu_sfts_ln.txt with the g_in loads issued after the row reduction)
        if (g_in) {
#pragma unroll
            for (int j = 0; j < VPL; ++j) gi[j] = load4(g_in + (size_t)row * ldg + (j * 32 + lane) * 4);
        }
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
            const int c = (j * 32 + lane) * 4;
            const float4 xv = load4(xr + c);
            d[j] = load4(dr + c);
            xh[j] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
            acc_g[j].x += d[j].x * xh[j].x; acc_g[j].y += d[j].y * xh[j].y;
            acc_g[j].z += d[j].z * xh[j].z; acc_g[j].w += d[j].w * xh[j].w;
            acc_b[j].x += d[j].x; acc_b[j].y += d[j].y; acc_b[j].z += d[j].z; acc_b[j].w += d[j].w;
            d[j].x *= gm[j].x; d[j].y *= gm[j].y; d[j].z *= gm[j].z; d[j].w *= gm[j].w;
            s1 += d[j].x + d[j].y + d[j].z + d[j].w;
            s2 += d[j].x * xh[j].x + d[j].y * xh[j].y + d[j].z * xh[j].z + d[j].w * xh[j].w;
        }
        const float c1 = warp_sum(s1) * (1.0f / D), c2 = warp_sum(s2) * (1.0f / D);
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
            const int c = (j * 32 + lane) * 4;
            float4 o = make_float4(rs * (d[j].x - c1 - xh[j].x * c2), rs * (d[j].y - c1 - xh[j].y * c2),
                                   rs * (d[j].z - c1 - xh[j].z * c2), rs * (d[j].w - c1 - xh[j].w * c2));
            if (g_in) { o.x += gi[j].x; o.y += gi[j].y; o.z += gi[j].z; o.w += gi[j].w; }
            if (g_out) store4(g_out + (size_t)row * ldg + c, o.x, o.y, o.z, o.w);
            if (g_bf16) store4(g_bf16 + (size_t)row * ldgb + c, sc * o.x, sc * o.y, sc * o.z, sc * o.w);
            acc_c[j].x += sc * o.x; acc_c[j].y += sc * o.y; acc_c[j].z += sc * o.z; acc_c[j].w += sc * o.w;
        }
    }
    // cross-warp reduction of the three column sums, one after the other through the same smem buffer
#pragma unroll 1
    for (int which = 0; which < 3; ++which) {
        __syncthreads();
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
            const float4 a = which == 0 ? acc_g[j] : (which == 1 ? acc_b[j] : acc_c[j]);
            store4(&red[warp][(j * 32 + lane) * 4], a.x, a.y, a.z, a.w);
        }
        __syncthreads();
        float* dst = which == 0 ? dgamma : (which == 1 ? dbeta : dcol);
        if (dst == nullptr) continue;      // uniform across the block
        for (int c = threadIdx.x; c < D; c += LNB_WARPS * 32) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < LNB_WARPS; ++w) s += red[w][c];
            atomicAdd(dst + c, s);         // ~300 CTAs x 768 columns: cheaper than a second launch per LayerNorm
        }
    }
}

size_t layernorm_bwd_workspace_bytes() { return 256; }   // kept in the ABI; the reduction now uses atomics

int layernorm_bwd(const void* dy, long long lddy, int dy_f32, const float* x, long long ldx, const float* mean,
                  const float* rstd, const float* gamma, const float* g_in, float* g_out, long long ldg,
                  void* g_bf16, long long ldgb, float* dgamma, float* dbeta, float* dcol, void* workspace,
                  size_t ws_bytes, int rows, int dim, const float* row_scale, int scale_group, const int* rows_dev,
                  cudaStream_t st) {
    if (scale_group <= 0) scale_group = 1;
    if (dim != D) return edb_set_error(EDB_ERR_SHAPE, "layernorm_bwd: only 768-wide rows are supported");
    if (rows <= 0) return EDB_OK;
    if (ws_bytes < layernorm_bwd_workspace_bytes()) return edb_set_error(EDB_ERR_WORKSPACE, "layernorm_bwd: workspace");
    int grid = num_sms() * 2;
    const int need = (rows + LNB_WARPS - 1) / LNB_WARPS;
    if (grid > need) grid = need;
    (void)workspace;
    if (dy_f32)
        ln_bwd_kernel<float><<<grid, LNB_WARPS * 32, 0, st>>>((const float*)dy, lddy, x, ldx, mean, rstd, gamma, g_in,
                                                              g_out, ldg, (__nv_bfloat16*)g_bf16, ldgb, dgamma, dbeta,
                                                              dcol, rows, row_scale, scale_group, rows_dev);
    else
        ln_bwd_kernel<__nv_bfloat16><<<grid, LNB_WARPS * 32, 0, st>>>((const __nv_bfloat16*)dy, lddy, x, ldx, mean, rstd,
                                                                      gamma, g_in, g_out, ldg, (__nv_bfloat16*)g_bf16,
                                                                      ldgb, dgamma, dbeta, dcol, rows, row_scale, scale_group,
                                                                      rows_dev);
    EDB_CHECK_LAUNCH();
    return EDB_OK;
}

// ------------------------------------------------------------------------------------------ column sums (bias grads)
// out[n] += sum_r src[r][n];  grid = (ceil(N/256), row_splits), 8 warps, each lane owns 8 adjacent columns.
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ src, long long ld, int rows, int N,
                                                     float* __restrict__ out) {
    __shared__ float red[8][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = blockIdx.x * 256 + lane * 8;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (c0 < N) {
        const int step = gridDim.y * 8;
        int r = blockIdx.y * 8 + warp;
        for (; r + step < rows; r += 2 * step) {         // two rows in flight per warp
            const T* p = src + (size_t)r * ld + c0;
            const T* q = p + (size_t)step * ld;
            const float4 a = load4(p), b = load4(p + 4), c = load4(q), d = load4(q + 4);
            acc[0] += a.x + c.x; acc[1] += a.y + c.y; acc[2] += a.z + c.z; acc[3] += a.w + c.w;
            acc[4] += b.x + d.x; acc[5] += b.y + d.y; acc[6] += b.z + d.z; acc[7] += b.w + d.w;
        }
        for (; r < rows; r += step) {
            const T* p = src + (size_t)r * ld + c0;
            const float4 a = load4(p), b = load4(p + 4);
            acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
            acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) red[warp][lane * 8 + i] = acc[i];
    __syncthreads();
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c < N) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
        atomicAdd(out + c, s);
    }
}

int colsum(const void* src, long long ld, int src_f32, int rows, int N, float* out, cudaStream_t st) {
    if (rows <= 0 || N <= 0) return EDB_OK;
    if (N % 8 || ld % 8) return edb_set_error(EDB_ERR_ALIGN, "colsum: N and pitch must be multiples of 8");
    const int gx = (N + 255) / 256;
    int splits = (rows + 255) / 256;
    const int cap = (4 * num_sms() + gx - 1) / gx;      // about four CTAs per SM over the whole grid
    if (splits > cap) splits = cap;
    dim3 grid(gx, splits);
    if (src_f32) colsum_kernel<float><<<grid, 256, 0, st>>>((const float*)src, ld, rows, N, out);
    else colsum_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)src, ld, rows, N, out);
    EDB_CHECK_LAUNCH();
    return EDB_OK;
}

// ------------------------------------------------------------------------------------------ casts and bf16 splits
__global__ void cast_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t n4) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = load4(src + i * 4);
        store4(dst + i * 4, v.x, v.y, v.z, v.w);
    }
}

int cast_f32_bf16(const float* src, void* dst, size_t n, cudaStream_t st) {
    if (n == 0) return EDB_OK;
    if (n % 4) return edb_set_error(EDB_ERR_ALIGN, "cast: element count must be a multiple of 4");
    const size_t n4 = n / 4;
    size_t blocks = (n4 + 255) / 256;
    if (blocks > (size_t)num_sms() * 16) blocks = (size_t)num_sms() * 16;
    cast_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, (__nv_bfloat16*)dst, n4);
    EDB_CHECK_LAUNCH();
    return EDB_OK;
}

__global__ void cast_rows_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int max_rows, int cols4,
                                 const int* __restrict__ rows_dev) {
    const size_t n4 = (size_t)min(max_rows, *rows_dev) * cols4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = load4(src + i * 4);
        store4(dst + i * 4, v.x, v.y, v.z, v.w);
    }
}

int cast_rows_f32_bf16(const float* src, void* dst, int max_rows, int cols, const int* rows_dev, cudaStream_t st) {
    if (max_rows <= 0) return EDB_OK;
    if (cols % 4) return edb_set_error(EDB_ERR_ALIGN, "cast_rows: cols must be a multiple of 4");
    size_t blocks = ((size_t)max_rows * (cols / 4) + 255) / 256;
    if (blocks > (size_t)num_sms() * 8) blocks = (size_t)num_sms() * 8;
    cast_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, (__nv_bfloat16*)dst, max_rows, cols / 4, rows_dev);
    EDB_CHECK_LAUNCH();
    return EDB_OK;
}

__global__ void zero_rows_kernel(uint8_t* base, long long row_bytes, const int* __restrict__ rows_dev, int nrows) {
    uint4* p = reinterpret_cast<uint4*>(base + (size_t)(*rows_dev) * row_bytes);
    const size_t n = (size_t)nrows * row_bytes / 16;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        p[i] = make_uint4(0u, 0u, 0u, 0u);
}

int zero_rows(void* base, long long row_bytes, const int* rows_dev, int nrows, cudaStream_t st) {
    if (nrows <= 0) return EDB_OK;
    if (row_bytes % 16) return edb_set_error(EDB_ERR_ALIGN, "zero_rows: row size must be a multiple of 16 bytes");
    zero_rows_kernel<<<32, 256, 0, st>>>((uint8_t*)base, row_bytes, rows_dev, nrows);
    EDB_CHECK_LAUNCH();
    return EDB_OK;
}

// fp32-faithful GEMMs on the bf16 tensor-core kernel: x = h + m + l (three bf16 pieces, 24 mantissa bits); the six
// products of order <= 2 are obtained from ONE GEMM by concatenating pieces along K:
//   role 0 (A side): [h | h | m | h | l | m]      role 1 (B side): [h | m | h | l | h | m]
// roles 2 / 3: the same A- / B-side piece orders concatenated along ROWS (dst is [6*rows][K]) -- operands whose reduction
// dimension is the row index (MN-major A/B of the dgrad / wgrad products).
__global__ void split_kernel(const float* __restrict__ src, long long ld, int rows, int K,
                             __nv_bfloat16* __restrict__ dst, int role) {
    const int kq = K / 4;
    const size_t total = (size_t)rows * kq;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / kq), c = (int)(i % kq) * 4;
        const float4 v = load4(src + (size_t)r * ld + c);
        const float x[4] = {v.x, v.y, v.z, v.w};
        float h[4], m[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            h[j] = __bfloat162float(__float2bfloat16(x[j]));
            const float r1 = x[j] - h[j];
            m[j] = __bfloat162float(__float2bfloat16(r1));
            l[j] = __bfloat162float(__float2bfloat16(r1 - m[j]));
        }
        const bool by_rows = role >= 2;
        __nv_bfloat16* o = by_rows ? dst + (size_t)r * K + c : dst + (size_t)r * (6 * (size_t)K) + c;
        const size_t step = by_rows ? (size_t)rows * K : (size_t)K;
        const float* seq[6];
        if ((role & 1) == 0) { seq[0] = h; seq[1] = h; seq[2] = m; seq[3] = h; seq[4] = l; seq[5] = m; }
        else                 { seq[0] = h; seq[1] = m; seq[2] = h; seq[3] = l; seq[4] = h; seq[5] = m; }
#pragma unroll
        for (int t = 0; t < 6; ++t) store4(o + (size_t)t * step, seq[t][0], seq[t][1], seq[t][2], seq[t][3]);
    }
}

int split_bf16x3(const float* src, long long ld, int rows, int K, void* dst, int role, cudaStream_t st) {
    if (rows <= 0) return EDB_OK;
    if (K % 8 || ld % 4) return edb_set_error(EDB_ERR_ALIGN, "split: K must be a multiple of 8, pitch of 4");
    const size_t total = (size_t)rows * (K / 4);
    size_t blocks = (total + 255) / 256;
    if (blocks > (size_t)num_sms() * 16) blocks = (size_t)num_sms() * 16;
    split_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, ld, rows, K, (__nv_bfloat16*)dst, role);
    EDB_CHECK_LAUNCH();
    return EDB_OK;
}

// out = dh * gelu'(pre), exact erf (fp32-faithful backward of nn.GELU, vit_pytorch.py:130)
__global__ void gelu_bwd_f32_kernel(const float* __restrict__ dh, const float* __restrict__ pre, float* __restrict__ out,
                                    size_t n4) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 d = load4(dh + i * 4), x = load4(pre + i * 4);
        const float xs[4] = {x.x, x.y, x.z, x.w};
        float gr[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float cdf = 0.5f * (1.0f + erff(xs[j] * 0.70710678118654752f));
            gr[j] = cdf + xs[j] * 0.3989422804014327f * expf(-0.5f * xs[j] * xs[j]);
        }
        store4(out + i * 4, d.x * gr[0], d.y * gr[1], d.z * gr[2], d.w * gr[3]);
    }
}

int gelu_bwd_f32(const float* dh, const float* pre, float* out, size_t n, cudaStream_t st) {
    if (n == 0) return EDB_OK;
    if (n % 4) return edb_set_error(EDB_ERR_ALIGN, "gelu_bwd: element count must be a multiple of 4");
    size_t blocks = (n / 4 + 255) / 256;
    if (blocks > (size_t)num_sms() * 16) blocks = (size_t)num_sms() * 16;
    gelu_bwd_f32_kernel<<<(unsigned)blocks, 256, 0, st>>>(dh, pre, out, n / 4);
    EDB_CHECK_LAUNCH();
    return EDB_OK;
}

// ------------------------------------------------------------------------------------------ patch im2col
// patches[(s*P + py*nx + px)][c*256 + ky*16 + kx] = img_m[b][c][py*16+ky][px*16+kx],  s = m*B + b  (vit_pytorch.py:455-457
// with kernel = stride = 16: the conv is a GEMM over non-overlapping patches).  One thread = 8 consecutive kx.
template <typename OutT>
__global__ void im2col_kernel(const float* __restrict__ i0, const float* __restrict__ i1, const float* __restrict__ i2,
                              int B, int H, int W, OutT* __restrict__ out, long long ldo, int col_off) {
    const int nx = W / 16, P = (H / 16) * nx;
    const size_t total = (size_t)3 * B * P * 96;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int q = (int)(i % 96);
        const size_t row = i / 96;
        const int p = (int)(row % P);
        const int s = (int)(row / P);
        const int m = s / B, b = s % B;
        const int c = q / 32, ky = (q % 32) / 2, half = q & 1;
        const int py = p / nx, px = p % nx;
        const float* img = m == 0 ? i0 : (m == 1 ? i1 : i2);
        const float* src = img + (((size_t)b * 3 + c) * H + py * 16 + ky) * W + px * 16 + half * 8;
        const float4 a = load4(src), d = load4(src + 4);
        OutT* o = out + row * ldo + col_off + q * 8;
        store4(o, a.x, a.y, a.z, a.w);
        store4(o + 4, d.x, d.y, d.z, d.w);
    }
}

int patch_im2col(const float* rgb, const float* ni, const float* ti, int B, int H, int W, void* out, long long ldo,
                 int out_f32, cudaStream_t st) {
    if (B <= 0) return EDB_OK;
    if (H % 16 || W % 16) return edb_set_error(EDB_ERR_SHAPE, "im2col: image sides must be multiples of 16");
    const size_t total = (size_t)3 * B * (H / 16) * (W / 16) * 96;
    size_t blocks = (total + 255) / 256;
    if (blocks > (size_t)num_sms() * 32) blocks = (size_t)num_sms() * 32;
    if (out_f32) im2col_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(rgb, ni, ti, B, H, W, (float*)out, ldo, 0);
    else im2col_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(rgb, ni, ti, B, H, W, (__nv_bfloat16*)out, ldo, 0);
    EDB_CHECK_LAUNCH();
    return EDB_OK;
}

// ------------------------------------------------------------------------------------------ token embedding
// x[s][0] = cls + pos[0] + coe*sie[cam[b]];  x[s][1+p] = patch_out[s*P+p] (+bias already added by the GEMM) + pos[1+p]
// + coe*sie[cam[b]]   (vit_pytorch.py:627-633)
__global__ void embed_kernel(const float* __restrict__ patch_out, const float* __restrict__ cls,
                             const float* __restrict__ pos, const float* __restrict__ sie,
                             const long long* __restrict__ cam, float coe, int S, int B, int P, float* __restrict__ x) {
    const int N = P + 1;
    const size_t total = (size_t)S * N * (D / 4);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % (D / 4)) * 4;
        const size_t row = i / (D / 4);
        const int t = (int)(row % N), s = (int)(row / N);
        float4 v = t == 0 ? load4(cls + c) : load4(patch_out + ((size_t)s * P + t - 1) * D + c);
        const float4 pe = load4(pos + (size_t)t * D + c);
        v.x += pe.x; v.y += pe.y; v.z += pe.z; v.w += pe.w;
        if (sie != nullptr) {
            const float4 se = load4(sie + (size_t)cam[s % B] * D + c);
            v.x += coe * se.x; v.y += coe * se.y; v.z += coe * se.z; v.w += coe * se.w;
        }
        store4(x + row * D + c, v.x, v.y, v.z, v.w);
    }
}

int embed_assemble(const float* patch_out, const float* cls, const float* pos, const float* sie, const long long* cam,
                   float coe, int S, int B, int P, float* x, cudaStream_t st) {
    if (S <= 0) return EDB_OK;
    const size_t total = (size_t)S * (P + 1) * (D / 4);
    size_t blocks = (total + 255) / 256;
    if (blocks > (size_t)num_sms() * 32) blocks = (size_t)num_sms() * 32;
    embed_kernel<<<(unsigned)blocks, 256, 0, st>>>(patch_out, cls, pos, sie, cam, coe, S, B, P, x);
    EDB_CHECK_LAUNCH();
    return EDB_OK;
}

// backward: dpos[t] += sum_s g[s][t];  dpatch(bf16)[s*P+t-1] = g[s][t];  dsie[cam] += coe * sum_t g[s][t]
template <typename OutT>
__global__ void __launch_bounds__(192) embed_bwd_pos_kernel(const float* __restrict__ g, int S, int N,
                                                            float* __restrict__ dpos, OutT* __restrict__ dpatch) {
    const int t = blockIdx.x, c = threadIdx.x * 4;
    const int s0 = blockIdx.y, sstep = gridDim.y;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = s0; s < S; s += sstep) {
        const float4 v = load4(g + ((size_t)s * N + t) * D + c);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        if (t > 0 && dpatch != nullptr) store4(dpatch + ((size_t)s * (N - 1) + t - 1) * D + c, v.x, v.y, v.z, v.w);
    }
    float* o = dpos + (size_t)t * D + c;
    atomicAdd(o, acc.x); atomicAdd(o + 1, acc.y); atomicAdd(o + 2, acc.z); atomicAdd(o + 3, acc.w);
}

__global__ void __launch_bounds__(192) embed_bwd_sie_kernel(const float* __restrict__ g, int B, int N,
                                                            const long long* __restrict__ cam, float coe,
                                                            float* __restrict__ dsie) {
    const int s = blockIdx.x, c = threadIdx.x * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = 0; t < N; ++t) {
        const float4 v = load4(g + ((size_t)s * N + t) * D + c);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    float* o = dsie + (size_t)cam[s % B] * D + c;
    atomicAdd(o, coe * acc.x); atomicAdd(o + 1, coe * acc.y); atomicAdd(o + 2, coe * acc.z); atomicAdd(o + 3, coe * acc.w);
}

int embed_assemble_bwd(const float* g, int S, int B, int P, const long long* cam, float coe, float* dpos, float* dsie,
                       void* dpatch, int dpatch_f32, cudaStream_t st) {
    if (S <= 0) return EDB_OK;
    int ysplit = S < 8 ? S : 8;
    if (dpatch_f32) embed_bwd_pos_kernel<float><<<dim3(P + 1, ysplit), 192, 0, st>>>(g, S, P + 1, dpos, (float*)dpatch);
    else embed_bwd_pos_kernel<__nv_bfloat16><<<dim3(P + 1, ysplit), 192, 0, st>>>(g, S, P + 1, dpos, (__nv_bfloat16*)dpatch);
    EDB_CHECK_LAUNCH();
    if (dsie != nullptr) {
        embed_bwd_sie_kernel<<<S, 192, 0, st>>>(g, B, P + 1, cam, coe, dsie);
        EDB_CHECK_LAUNCH();
    }
    return EDB_OK;
}

// ------------------------------------------------------------------------------------------ fused SGD-momentum
// torch.optim.SGD semantics of solver/make_optimizer.py:6-22 over the flat arena: g += wd*p; buf = mu*buf + g (buf = g on
// the first step); p -= lr*buf, with lr*bias_lr_factor and wd_bias on 64-element chunks flagged as bias.  Also refreshes the
// bf16 shadow of the parameters and scales the incoming gradient (1/world_size after the allreduce).
__global__ void sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf,
                           __nv_bfloat16* __restrict__ p16, const unsigned char* __restrict__ flags, size_t n4, float lr,
                           float mu, float wd, float wd_bias, float bias_lr_factor, float gscale, int first) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned char f = flags[i >> 4];
        if (f & 2) continue;                        // frozen / padding chunk
        const float l = (f & 1) ? lr * bias_lr_factor : lr, w = (f & 1) ? wd_bias : wd;
        float4 pv = load4(p + i * 4);
        const float4 gv = load4(g + i * 4);
        float4 bv = first ? make_float4(0.f, 0.f, 0.f, 0.f) : load4(buf + i * 4);
        const float gx = gv.x * gscale + w * pv.x, gy = gv.y * gscale + w * pv.y, gz = gv.z * gscale + w * pv.z,
                    gw = gv.w * gscale + w * pv.w;
        bv.x = first ? gx : mu * bv.x + gx; bv.y = first ? gy : mu * bv.y + gy;
        bv.z = first ? gz : mu * bv.z + gz; bv.w = first ? gw : mu * bv.w + gw;
        pv.x -= l * bv.x; pv.y -= l * bv.y; pv.z -= l * bv.z; pv.w -= l * bv.w;
        store4(buf + i * 4, bv.x, bv.y, bv.z, bv.w);
        store4(p + i * 4, pv.x, pv.y, pv.z, pv.w);
        store4(p16 + i * 4, pv.x, pv.y, pv.z, pv.w);
    }
}

int sgd_step(float* p, const float* g, float* buf, void* p16, const unsigned char* flags, size_t n, float lr, float mu,
             float wd, float wd_bias, float bias_lr_factor, float gscale, int first, cudaStream_t st) {
    if (n == 0) return EDB_OK;
    if (n % 64) return edb_set_error(EDB_ERR_ALIGN, "sgd: arena length must be a multiple of 64");
    const size_t n4 = n / 4;
    size_t blocks = (n4 + 255) / 256;
    if (blocks > (size_t)num_sms() * 16) blocks = (size_t)num_sms() * 16;
    sgd_kernel<<<(unsigned)blocks, 256, 0, st>>>(p, g, buf, (__nv_bfloat16*)p16, flags, n4, lr, mu, wd, wd_bias,
                                                bias_lr_factor, gscale, first);
    EDB_CHECK_LAUNCH();
    return EDB_OK;
}

}  // namespace edb

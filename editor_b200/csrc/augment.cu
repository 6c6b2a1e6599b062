// Training-time input pipeline of the reference on the GPU, for uint8 HWC images that were decoded on
// the host:   T.Resize(SIZE_TRAIN, interpolation=3) -> RandomHorizontalFlip -> Pad(PADDING) -> RandomCrop(SIZE_TRAIN) ->
// ToTensor -> Normalize -> RandomErasing(mode='pixel', max_count=1)      (data/datasets/make_dataloader.py:245-253, :55-140)
// for the three modality images of every sample (data/datasets/bases.py:100-103), 3*B images per launch.
//
//   aug_hpass_kernel   horizontal pass of Pillow's bicubic resample (libImaging/Resample.c: 22-bit fixed-point coefficients,
//                      rounded and clipped to uint8 between the passes) -> uint8 [Hs][W][3]; skipped when Ws == W
//   aug_main_kernel    vertical pass + flip + zero padding + crop + /255 + (x - mean) / std + erase rectangle filled with
//                      N(0,1) noise (given, or Philox4x32-10 + Box-Muller per element) -> fp32 [B][3][H][W] per modality
//
// HBM-bound byte work: 3 B/pixel in, 12 B/pixel out; one thread per output pixel, x fastest (coalesced 4-byte stores per
// channel plane).  The random DRAWS (flip, crop offsets, erase rectangle) are per-image scalars supplied by the host
// (editor_b200/data.py) -- parity with torchvision is defined for given draws.
#include "abi_internal.h"

namespace edb {

constexpr int AUG_PREC = 22;

struct AugTables {
    const int* hb; const int* hk; int ksh;      // horizontal: bounds [W][2] (first tap, taps), coefficients [W][ksh]
    const int* vb; const int* vk; int ksv;      // vertical:   bounds [H][2], coefficients [H][ksv]
};

__device__ __forceinline__ int clip8(int v) { return min(max(v >> AUG_PREC, 0), 255); }

// one thread per (image, row, out column): 3 channels
__global__ void __launch_bounds__(256) aug_hpass_kernel(const uint8_t* __restrict__ s0, const uint8_t* __restrict__ s1,
                                                        const uint8_t* __restrict__ s2, int B, int Hs, int Ws, int W,
                                                        AugTables t, uint8_t* __restrict__ tmp) {
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long total = 3LL * B * Hs * W;
    if (idx >= total) return;
    const int x = (int)(idx % W);
    const long long r = idx / W;
    const int y = (int)(r % Hs);
    const int img = (int)(r / Hs);                 // m * B + b
    const int m = img / B, b = img - m * B;
    const uint8_t* src = (m == 0 ? s0 : (m == 1 ? s1 : s2)) + ((size_t)b * Hs + y) * Ws * 3;
    const int x0 = t.hb[2 * x], n = t.hb[2 * x + 1];
    const int* k = t.hk + (size_t)x * t.ksh;
    int a0 = 1 << (AUG_PREC - 1), a1 = a0, a2 = a0;
    for (int i = 0; i < n; ++i) {
        const uint8_t* px = src + (size_t)(x0 + i) * 3;
        const int kv = k[i];
        a0 += px[0] * kv; a1 += px[1] * kv; a2 += px[2] * kv;
    }
    uint8_t* o = tmp + (((size_t)img * Hs + y) * W + x) * 3;
    o[0] = (uint8_t)clip8(a0); o[1] = (uint8_t)clip8(a1); o[2] = (uint8_t)clip8(a2);
}

// Philox4x32-10 (Salmon et al. 2011), counter = (element index lo, hi, image, 0), key = the image's seed
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__global__ void __launch_bounds__(256) aug_main_kernel(const uint8_t* __restrict__ s0, const uint8_t* __restrict__ s1,
                                                       const uint8_t* __restrict__ s2, const uint8_t* __restrict__ tmp,
                                                       int B, int Hs, int Ws, int H, int W, int pad, AugTables t,
                                                       float m0, float m1, float m2, float d0, float d1, float d2,
                                                       const EdbAugImage* __restrict__ prm, const float* __restrict__ noise,
                                                       float* __restrict__ o0, float* __restrict__ o1, float* __restrict__ o2) {
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long total = 3LL * B * H * W;
    if (idx >= total) return;
    const int x = (int)(idx % W);
    const long long r = idx / W;
    const int y = (int)(r % H);
    const int img = (int)(r / H);
    const int m = img / B, b = img - m * B;
    const EdbAugImage p = prm[img];
    float* out = (m == 0 ? o0 : (m == 1 ? o1 : o2)) + (size_t)b * 3 * H * W + (size_t)y * W + x;
    const size_t plane = (size_t)H * W;
    if (p.e_h > 0 && y >= p.e_top && y < p.e_top + p.e_h && x >= p.e_left && x < p.e_left + p.e_w) {
        // RandomErasing 'pixel' mode: per-element N(0,1) (make_dataloader.py:42-47,118-124)
        if (noise != nullptr) {
            const float* nz = noise + (size_t)img * 3 * plane + (size_t)y * W + x;
            out[0] = nz[0]; out[plane] = nz[plane]; out[2 * plane] = nz[2 * plane];
        } else {
            uint32_t rnd[4];
            philox4x32_10((uint32_t)(y * W + x), 0u, (uint32_t)img, 0u, p.seed_lo, p.seed_hi, rnd);
            // Box-Muller on (0,1] uniforms
            const float u0 = ((float)(rnd[0] >> 8) + 1.0f) * (1.0f / 16777216.0f), u1 = (float)(rnd[1] >> 8) * (1.0f / 16777216.0f);
            const float u2 = ((float)(rnd[2] >> 8) + 1.0f) * (1.0f / 16777216.0f), u3 = (float)(rnd[3] >> 8) * (1.0f / 16777216.0f);
            const float r0 = sqrtf(-2.0f * logf(u0)), r1 = sqrtf(-2.0f * logf(u2));
            float s, c;
            sincospif(2.0f * u1, &s, &c);
            out[0] = r0 * c; out[plane] = r0 * s;
            out[2 * plane] = r1 * cospif(2.0f * u3);
        }
        return;
    }
    // crop offsets are in the padded image; the padded border is 0 BEFORE ToTensor / Normalize
    const int ry = y + p.top - pad;
    int rx = x + p.left - pad;
    int v0 = 0, v1 = 0, v2 = 0;
    if (ry >= 0 && ry < H && rx >= 0 && rx < W) {
        if (p.flip) rx = W - 1 - rx;
        const uint8_t* base;          // image after the horizontal pass: [Hs][W][3]
        if (Ws == W) base = (m == 0 ? s0 : (m == 1 ? s1 : s2)) + (size_t)b * Hs * W * 3;
        else base = tmp + (size_t)img * Hs * W * 3;
        if (Hs == H) {
            const uint8_t* px = base + ((size_t)ry * W + rx) * 3;
            v0 = px[0]; v1 = px[1]; v2 = px[2];
        } else {
            const int y0 = t.vb[2 * ry], n = t.vb[2 * ry + 1];
            const int* k = t.vk + (size_t)ry * t.ksv;
            int a0 = 1 << (AUG_PREC - 1), a1 = a0, a2 = a0;
            for (int i = 0; i < n; ++i) {
                const uint8_t* px = base + ((size_t)(y0 + i) * W + rx) * 3;
                const int kv = k[i];
                a0 += px[0] * kv; a1 += px[1] * kv; a2 += px[2] * kv;
            }
            v0 = clip8(a0); v1 = clip8(a1); v2 = clip8(a2);
        }
    }
    // ToTensor: uint8 -> float32 / 255;  Normalize: (x - mean) / std   (IEEE divisions: bit-identical to torchvision)
    out[0] = __fdiv_rn(__fdiv_rn((float)v0, 255.0f) - m0, d0);
    out[plane] = __fdiv_rn(__fdiv_rn((float)v1, 255.0f) - m1, d1);
    out[2 * plane] = __fdiv_rn(__fdiv_rn((float)v2, 255.0f) - m2, d2);
}

size_t augment_workspace_bytes(int B, int Hs, int Ws, int W) { return Ws == W ? 0 : (size_t)3 * B * Hs * W * 3; }

int augment_u8(const uint8_t* s0, const uint8_t* s1, const uint8_t* s2, int B, int Hs, int Ws, int H, int W, int pad,
               const int* hb, const int* hk, int ksh, const int* vb, const int* vk, int ksv, const float* mean,
               const float* stdv, const EdbAugImage* params, const float* noise, float* o0, float* o1, float* o2,
               void* workspace, size_t ws_bytes, cudaStream_t st) {
    if (B <= 0) return EDB_OK;
    if (Hs <= 0 || Ws <= 0 || H <= 0 || W <= 0 || pad < 0) return edb_set_error(EDB_ERR_SHAPE, "augment: bad geometry");
    if ((Ws != W && (hb == nullptr || hk == nullptr)) || (Hs != H && (vb == nullptr || vk == nullptr)))
        return edb_set_error(EDB_ERR_SHAPE, "augment: resample tables missing for a resized axis");
    if (ws_bytes < augment_workspace_bytes(B, Hs, Ws, W)) return edb_set_error(EDB_ERR_WORKSPACE, "augment: workspace too small");
    AugTables t{hb, hk, ksh, vb, vk, ksv};
    if (Ws != W) {
        const long long total = 3LL * B * Hs * W;
        aug_hpass_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(s0, s1, s2, B, Hs, Ws, W, t, (uint8_t*)workspace);
        EDB_CHECK_LAUNCH();
    }
    const long long total = 3LL * B * H * W;
    aug_main_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(s0, s1, s2, (const uint8_t*)workspace, B, Hs, Ws, H, W, pad, t,
                                                                    mean[0], mean[1], mean[2], stdv[0], stdv[1], stdv[2], params,
                                                                    noise, o0, o1, o2);
    EDB_CHECK_LAUNCH();
    return EDB_OK;
}

}  // namespace edb

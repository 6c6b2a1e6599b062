/* editor_b200 -- C ABI of the B200-native EDITOR hot path (libeditor_b200.so).
 *
 * The reference (924973292/EDITOR) has no FFI layer: its hot path is the Python module surface
 * `modeling.make_model(cfg, num_class, camera_num)` -> `EDITOR.forward` (modeling/make_model.py:150-258,371-374).
 * This header is the boundary a maintainer binds instead of the ATen calls listed in SURVEY.md section 2b;
 * INTEGRATION.md shows the ctypes stub.  Every entry point cites the reference code it replaces.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no C++ or torch types.  All pointers are DEVICE pointers unless a
 *     parameter name ends in `_host`.  Tensors are row-major and contiguous unless a pitch (`ld*`) is given.
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), never allocates or frees device
 *     memory and never synchronises; scratch comes from the caller (`workspace`, sized by *_workspace_bytes).
 *   - return value: EDB_OK (0) or a negative EDB_ERR_*; edb_last_error() gives the thread-local message.
 *   - `prec`: EDB_PREC_BF16 = bf16 tensor-core operands, fp32 accumulate, fp32 residual stream (what the
 *     reference's autocast training loop does, engine/processor.py:79);  EDB_PREC_FP32 = fp32-faithful
 *     (3-way bf16 operand split on the same tensor-core kernel, fp32 attention) for the fp32 eval path
 *     (engine/processor.py:176-186).
 */
#ifndef EDITOR_B200_H
#define EDITOR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EDB_OK 0
#define EDB_ERR_SHAPE (-1)
#define EDB_ERR_ALIGN (-2)
#define EDB_ERR_CUDA (-3)
#define EDB_ERR_WORKSPACE (-4)
#define EDB_ERR_UNSUPPORTED (-5)

#define EDB_PREC_BF16 0
#define EDB_PREC_FP32 1

/* GEMM epilogues */
#define EPI_STORE 0    /* D = alpha*acc + bias                                                     */
#define EPI_GELU 1     /* D = gelu_erf(pre), pre = acc + bias (vit_pytorch.py:139-145); optional out2: bf16 D -> out2 =
                        * gelu'(pre), the factor EPI_GELU_BWD multiplies by; fp32 D (EDB_PREC_FP32) -> out2 = pre         */
#define EPI_RESIDUAL 2 /* D(f32) = aux(f32) + acc + bias; D may alias aux           (vit_pytorch.py:217-219) */
#define EPI_GELU_BWD 3 /* D = acc * aux, aux = the bf16 gelu'(pre) saved by EPI_GELU (backward of nn.GELU)       */
#define EPI_ATOMIC 4   /* D(f32) += acc   (split-K partial sums, D pre-zeroed by the caller)        */

int edb_version(void);
const char* edb_last_error(void);

/* D[M,N] = epilogue( sum_k A(m,k) * B(n,k) ), bf16 operands, fp32 accumulation on tcgen05 tensor cores.
 * a_mn_major = 0: A stored [M][lda] (k contiguous);   1: A stored [K][lda] (m contiguous).  Same for B with N.
 * Replaces every nn.Linear forward / dgrad / wgrad on the path (vit_pytorch.py:133-136,181-183,235-237). */
typedef struct EdbGemmDesc {
    int M, N, K;
    const void* A; long long lda; int a_mn_major;
    const void* B; long long ldb; int b_mn_major;
    void* D; long long ldd; int out_f32;
    int epilogue;
    const float* bias;                 /* [N] or NULL */
    const void* aux; long long ld_aux; int aux_f32;
    void* out2; long long ld_out2;     /* EPI_GELU only; same dtype as D */
    float alpha;
    int split_k;                       /* >1 only with EPI_ATOMIC */
    const float* row_scale;            /* EPI_RESIDUAL only: D = aux + row_scale[row/scale_group]*(acc+bias) -- DropPath   */
    int scale_group;                   /* (vit_pytorch.py:52-69,217-219); NULL = 1.0                                      */
    const int* M_dev;                  /* optional DEVICE ints overriding M / K at run time (M, K then are upper bounds used  */
    const int* K_dev;                  /* for the launch): row counts of the packed HMA matrices are produced on the GPU     */
    float* colsum;                     /* optional, EPI_STORE / EPI_GELU_BWD: colsum[n] += sum_m D[m][n] (fp32, before the  */
                                       /* bf16 rounding of D) -- the bias gradient of the Linear whose output gradient D is  */
                                       /* (vit_pytorch.py:133-136), fused so that D is not read back from HBM for it         */
} EdbGemmDesc;

int edb_gemm_bf16(const EdbGemmDesc* desc, void* stream);
/* Tile mode of edb_gemm_bf16 (process-wide): 0 = automatic -- CTA pairs running tcgen05.mma.cta_group::2 (256 x 256 x 16)
 * wherever N > 128 and M > 128 --, 1 = single-CTA 128 x 256 tiles only.  Results are bit-identical in both modes (same
 * accumulation order); the switch exists for A/B timing and for the tests that prove exactly that. */
int edb_gemm_set_mode(int mode);

/* ---- row kernels (HBM-bound) ------------------------------------------------------------------------------ */

/* y = LayerNorm(x) over 768-wide rows; optional per-row mean / rstd for the backward.
 * Replaces nn.LayerNorm in Block (vit_pytorch.py:206,211,643, eps 1e-6) and BlockMask (:265-296, eps 1e-5). */
int edb_layernorm_fwd(const float* x, long long ldx, const float* gamma, const float* beta, float eps, void* y,
                      long long ldy, int y_f32, float* mean, float* rstd, int rows, int dim, const int* rows_dev,
                      void* stream);
/* rows_dev (optional, everywhere it appears): DEVICE int with the actual row count; `rows` is then the launch bound. */

/* g_out = g_in + dLN(dy) (g_in may be NULL, may alias g_out); optional bf16 copy of g_out; dgamma/dbeta/dcol are
 * ACCUMULATED (+=), dcol = column sums of g_out (the bias gradient of the Linear feeding this residual value). */
size_t edb_layernorm_bwd_workspace_bytes(void);
int edb_layernorm_bwd(const void* dy, long long lddy, int dy_f32, const float* x, long long ldx, const float* mean,
                      const float* rstd, const float* gamma, const float* g_in, float* g_out, long long ldg,
                      void* g_bf16, long long ldgb, float* dgamma, float* dbeta, float* dcol, void* workspace,
                      size_t ws_bytes, int rows, int dim, const float* row_scale, int scale_group, const int* rows_dev,
                      void* stream);
/* row_scale (optional): g_bf16 and dcol carry row_scale[row/scale_group]*g_out -- the gradient entering a DropPath-scaled
 * branch (vit_pytorch.py:52-69). */

/* out[n] += sum_r src[r][n]  (bias gradients of nn.Linear, vit_pytorch.py:133-136,181-183) */
int edb_colsum(const void* src, long long ld, int src_f32, int rows, int n, float* out, void* stream);

/* fp32 -> bf16 (weights once per step, activations) */
int edb_cast_f32_bf16(const float* src, void* dst, size_t n, void* stream);
/* same over rows_dev[0] rows of `cols` contiguous elements (upper bound max_rows) */
int edb_cast_rows_f32_bf16(const float* src, void* dst, int max_rows, int cols, const int* rows_dev, void* stream);
/* zero `nrows` rows of `row_bytes` bytes starting at row rows_dev[0]: the K-padding of split-K wgrads over packed rows */
int edb_zero_rows(void* base, long long row_bytes, const int* rows_dev, int nrows, void* stream);

/* Fused SGD-momentum over the flat parameter arena (torch.optim.SGD semantics of
 * solver/make_optimizer.py:6-22: one group per tensor, bias lr x BIAS_LR_FACTOR, weight decay added to the gradient).
 * flags: one byte per 64-element chunk (bit0 = bias group, bit1 = skip).  Also rewrites the bf16 shadow p16 and applies
 * gscale to the gradient (1/world_size after the data-parallel allreduce). */
int edb_sgd_step(float* p, const float* g, float* buf, void* p16, const unsigned char* flags, size_t n, float lr,
                 float momentum, float wd, float wd_bias, float bias_lr_factor, float gscale, int first, void* stream);

/* fp32 -> 3-piece bf16 split laid out along K for the fp32-faithful GEMM: dst is [rows][6*K] bf16;
 * role 0 = A-side order, role 1 = B-side order; roles 2 / 3: the same orders concatenated along rows, dst [6*rows][K]
 * (operands whose reduction dimension is the row index: dgrad weights, wgrad activations).  EDB_PREC_FP32 path only. */
int edb_split_bf16x3(const float* src, long long ld, int rows, int K, void* dst, int role, void* stream);

/* patches[(m*B+b)*P + p][c*256+ky*16+kx] of the three modality images [B,3,H,W]: the k16/s16 patch conv as a GEMM
 * operand (PatchEmbed_overlap.forward, vit_pytorch.py:455-457). */
int edb_patch_im2col(const float* rgb, const float* ni, const float* ti, int B, int H, int W, void* out, long long ldo,
                     int out_f32, void* stream);

/* x[s][t] = (t==0 ? cls : patch_out[s*P+t-1]) + pos[t] + coe*sie[cam[s%B]]   (Trans.forward, vit_pytorch.py:627-633);
 * sie may be NULL.  cam is int64[B]. */
int edb_embed_assemble(const float* patch_out, const float* cls, const float* pos, const float* sie,
                       const long long* cam, float coe, int S, int B, int P, float* x, void* stream);
/* dpos += sum_s g;  dsie[cam] += coe*sum_t g;  dpatch (bf16 or fp32, [S*P][768]) = g[:,1:]  */
int edb_embed_assemble_bwd(const float* g, int S, int B, int P, const long long* cam, float coe, float* dpos,
                           float* dsie, void* dpatch, int dpatch_f32, void* stream);
/* out = dh * gelu'(pre) with exact erf: backward of nn.GELU in the fp32-faithful mode (vit_pytorch.py:130,139-145) */
int edb_gelu_bwd_f32(const float* dh, const float* pre, float* out, size_t n, void* stream);

/* ---- attention ---------------------------------------------------------------------------------------------- */

/* softmax(q k^T * scale) v per (sequence, head) on a packed [rows][3*heads*64] qkv matrix; optionally stores the
 * post-softmax maps P (the reference returns them, vit_pytorch.py:195-196; SFTS consumes them, SFTS.py:145-153).
 * Sequences: seq_off (nseq+1 int32 row offsets, device) or fixed_len.  impl: 0 = tensor-core kernel when the shape
 * allows (bf16, fixed_len 129), 1 = CUDA-core kernel (any length <= 256, fp32 or bf16 storage), 2 = tensor-core
 * kernel for packed var-len sequences (bf16, length <= 256, P stored as [ceil(max_len/128)*128][ldp] blocks with
 * ldp = 128 or 256, zero outside the sequence).
 * Also serves AttentionMask (vit_pytorch.py:240-258) on packed kept tokens. */
typedef struct EdbAttnDesc {
    const void* qkv; long long ld_qkv;
    void* out; long long ld_out;            /* forward: O.  backward: unused (may be NULL) */
    void* P; long long p_rows; long long ldp;
    const int* seq_off; int fixed_len; int nseq; int heads; int max_len;
    float scale;
    int f32;            /* storage type of qkv/out/P/d_out/d_qkv: 1 = fp32, 0 = bf16 */
    int impl;
    const void* d_out; long long ld_dout;   /* backward */
    void* d_qkv;                            /* backward, same pitch as qkv */
    long long total_rows;                   /* impl 2: rows of the packed qkv / out matrices */
} EdbAttnDesc;
int edb_attention_fwd(const EdbAttnDesc* desc, void* stream);
int edb_attention_bwd(const EdbAttnDesc* desc, void* stream);

/* ---- SFTS: token selection -------------------------------------------------------------------------------- */

/* counts[b][p] = #pixels of 16x16 window p whose 3-modality/3-channel mean is > 0
 * (Frequency_based_Token_Selection.forward/.mask, Frequency.py:42-56,65-81; Haar round trip == identity). */
int edb_freq_counts(const float* rgb, const float* ni, const float* ti, int B, int H, int W, int* counts, void* stream);

/* 128-bit set of torch.topk(vals, k) per row of 128 values (int32 or fp32), CUDA tie rule; store or OR into mask[row][4]
 * (Frequency.py:58-63; SFTS.py:155-158). */
int edb_topk_mask(const void* vals, int vals_f32, long long ld, int rows, int n, int k, unsigned* mask, int accumulate,
                  void* stream);

/* Part_Attention.forward (SFTS.py:145-162): cls row of A_{L-1}...A_0 per (sequence, head), per-head top-k, OR into
 * index[s % B] (SFTS.py:187-190).  maps: host array of `layers` device pointers, each [(s*heads+h)][p_rows][ldp]. */
int edb_rollout_topk(const void* const* maps_host, int layers, int maps_f32, int nseq, int B, int heads,
                     long long p_rows, long long ldp, int k, unsigned* index, unsigned* mod_mask, float* rows_out,
                     void* stream);

/* seq_off[b] = sum_{b'<b}(1 + popcount(index[b'])), seq_off3 = 3*seq_off (B+1 entries each, device). */
int edb_index_finalize(const unsigned* index, int B, int* seq_off, int* seq_off3, void* stream);

/* SFTS.forward (SFTS.py:208-222) fused with packing: kept rows (cls + selected) of tokens[3][B][129][768] go to
 * packed[3][cap][768] at row seq_off[b]+rank; loss_bcc (optional, pre-zeroed) += the three background MSEs. */
int edb_sfts_pack_fwd(const float* tokens, const unsigned* index, const int* seq_off, int B, long long cap,
                      float* packed, float* loss_bcc, void* stream);
int edb_sfts_pack_bwd(const float* tokens, const unsigned* index, const int* seq_off, int B, long long cap,
                      const float* d_packed, const float* g_loss, float* d_tokens, void* stream);

/* per-modality packed rows <-> joint rows (torch.cat([RGB,NIR,TIR],dim=1), vit_pytorch.py:324); dir 0 mod->joint */
int edb_joint_gather(float* mod, long long cap, float* joint, const int* seq_off, int B, int max_len, int dir,
                     void* stream);

/* make_model.py:186-203: cls rows, patch sums / count(RGB rows != 0) from the joint packed HMA output */
int edb_pool_fwd(const float* x, const int* seq_off, int B, float* cls_out, float* patch_mean, int* num, void* stream);
int edb_pool_bwd(const float* d_cls, const float* d_patch, const int* seq_off, const int* num, int B, int max_len,
                 float* dx, void* stream);

/* rows[m][b] = packed[m][seq_off[b]] (dir 0) or packed[m][seq_off[b]] += rows[m][b] (dir 1): HMA cls tokens for OCFR
 * (vit_pytorch.py:319-323). */
int edb_cls_rows(float* packed, long long cap, const int* seq_off, int B, float* rows, int dir, void* stream);

/* ---- tail of EDITOR.forward and the loss (tiny [B, 768..2304] tensors) -------------------------------------- */

/* nn.BatchNorm1d, training mode (BNNeck: FUSE_BN / AL_BN / BACKBONE_BN, make_model.py:114-141,165-171,209): batch
 * statistics, running_mean/var updated in place (momentum 0.1, unbiased variance), mean / invstd saved for bwd. */
int edb_bn1d_fwd(const float* x, long long ldx, int B, int F, const float* gamma, const float* beta, float* run_mean,
                 float* run_var, float momentum, float eps, float* y, long long ldy, float* save_mean,
                 float* save_invstd, void* stream);
/* dgamma / dbeta are accumulated (+=) */
int edb_bn1d_bwd(const float* dy, long long lddy, const float* x, long long ldx, int B, int F, const float* gamma,
                 const float* save_mean, const float* save_invstd, float* dx, long long lddx, float* dgamma,
                 float* dbeta, void* stream);

/* OCFR.forward (fusion_part/OCFR.py:44-84) on x = cls tokens [3][B][768]: L2-normalise, per-id batch centres, EMA into the
 * three [C][768] memory banks (before the loss), loss (pre-zeroed) += sum_m MSE(centres[label], fn).  fn / inv_norm are
 * kept for the backward. */
int edb_ocfr_fwd(const float* x, const long long* label, int B, int C, float* c_rgb, float* c_nir, float* c_tir,
                 float momentum, float* fn, float* inv_norm, float* loss, void* stream);
int edb_ocfr_bwd(const float* fn, const float* inv_norm, const long long* label, int B, float* c_rgb, float* c_nir,
                 float* c_tir, const float* g_loss, float* dx, void* stream);

/* CrossEntropyLabelSmooth (layers/softmax_loss.py:23-34): loss (pre-zeroed) += value; dlogits (optional) = d loss/d logits */
int edb_ce_smooth(const float* logits, long long ld, const long long* label, int B, int C, float eps, float* loss,
                  float* dlogits, long long ldd, void* stream);

/* TripletLoss(margin=None) (layers/triplet_loss.py:16-31,51-105,122-136): fp32 pairwise distances, batch-hard mining,
 * SoftMarginLoss.  fwd fills the workspace (distances, indices, coefficients) that bwd consumes. */
size_t edb_triplet_workspace_bytes(int B);
int edb_triplet_fwd(const float* x, long long ld, const long long* label, int B, int F, float* loss, void* workspace,
                    size_t ws_bytes, void* stream);
int edb_triplet_bwd(const float* x, long long ld, int B, int F, const void* workspace, const float* g_loss, float* dx,
                    long long ldd, int accumulate, void* stream);

/* y[i] = a[0] * x[i] (a on the device): upstream-gradient scaling of a pre-computed gradient */
int edb_scale_by(const float* x, const float* a, float* y, size_t n, void* stream);

/* ---- retrieval evaluation (the step after the eval forward, utils/metrics.py) ------------------------ */

/* feats[n][:] /= max(|feats[n]|_2, eps), in place: F.normalize(feats, dim=1, p=2) of R1_mAP_eval.compute
 * (utils/metrics.py:255-256; eps = 1e-12 is torch's default) */
int edb_eval_normalize(float* feats, long long ld, int n, int f, float eps, void* stream);
/* dist[q][g] = |qf_q|^2 + |gf_g|^2 - 2 qf_q . gf_g, fp32: euclidean_distance (utils/metrics.py:12-18; squared, no sqrt) */
int edb_eval_distmat(const float* qf, long long ldq, int q, const float* gf, long long ldg, int g, int f, float* dist,
                     long long ldd, void* stream);
/* eval_func (utils/metrics.py:133-191; key = camera id) / eval_func_msrv (:36-130; key = scene id) without the argsort:
 * per query, gallery items with the query's pid AND key are removed; ap[q] = average precision (fp64), first_rank[q] =
 * 1-based rank of the first correct match among the kept items, -1 when the query identity is absent from the gallery
 * (the reference skips such queries).  Equal distances rank by ascending gallery index.  *overflow is incremented for
 * queries with more than 2048 correct matches (not evaluated).  CMC[r] = mean_q [first_rank <= r+1], mAP = mean_q ap. */
int edb_eval_rank(const float* dist, long long ldd, int q, int g, const long long* q_pid, const long long* g_pid,
                  const long long* q_key, const long long* g_key, double* ap, int* first_rank, int* overflow, void* stream);

/* ---- training-time input pipeline (the step before the path, data/datasets/make_dataloader.py) -------- */

/* Per-image random draws of the augmentation, made on the host (editor_b200/data.py): RandomHorizontalFlip,
 * RandomCrop offsets inside the padded image (0 .. 2*PADDING), RandomErasing rectangle (e_h == 0: no erasing) and the
 * Philox key of its noise. */
typedef struct EdbAugImage {
    int flip, top, left;
    int e_top, e_left, e_h, e_w;
    unsigned seed_lo, seed_hi;
} EdbAugImage;

/* T.Resize(SIZE_TRAIN, interpolation=3) -> RandomHorizontalFlip -> Pad(PADDING) -> RandomCrop(SIZE_TRAIN) -> ToTensor ->
 * Normalize(mean, std) -> RandomErasing(mode='pixel', max_count=1)  (make_dataloader.py:245-253; RandomErasing :55-140)
 * for the 3 x B modality images of a batch (bases.py:100-103).  src_*: uint8 [B][Hs][Ws][3] (decoded images of ONE size);
 * out_*: fp32 [B][3][H][W].  hb/hk (vb/vk): bounds [W][2] ([H][2]) and 22-bit fixed-point coefficients [W][ksh] ([H][ksv]) of
 * Pillow's bicubic resample for Ws -> W (Hs -> H), needed only for an axis that is resized.  params: [3*B] device array,
 * image m*B + b.  noise: optional fp32 [3*B][3][H][W] normal draws used inside the erase rectangles (tests inject
 * torch's); NULL -> Philox4x32-10 + Box-Muller keyed by params[].seed_*.  mean/std: host pointers to 3 floats. */
size_t edb_augment_workspace_bytes(int B, int Hs, int Ws, int W);
int edb_augment_u8(const unsigned char* src_rgb, const unsigned char* src_ni, const unsigned char* src_ti, int B, int Hs,
                   int Ws, int H, int W, int pad, const int* hb, const int* hk, int ksh, const int* vb, const int* vk,
                   int ksv, const float* mean, const float* std, const EdbAugImage* params, const float* noise,
                   float* out_rgb, float* out_ni, float* out_ti, void* workspace, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EDITOR_B200_H */

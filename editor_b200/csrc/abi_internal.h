// Internal declarations shared by the .cu translation units behind include/editor_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>
#include "../../include/editor_b200.h"

namespace edb {

int edb_set_error(int code, const char* msg);

#define EDB_CHECK_LAUNCH()                                                              \
    do {                                                                                \
        cudaError_t e__ = cudaGetLastError();                                           \
        if (e__ != cudaSuccess) return edb::edb_set_error(EDB_ERR_CUDA, cudaGetErrorString(e__)); \
    } while (0)

#define EDB_TRY(expr)                  \
    do {                               \
        int rc__ = (expr);             \
        if (rc__ != EDB_OK) return rc__; \
    } while (0)

int num_sms();
int make_tmap_bf16(CUtensorMap* map, const void* base, long long inner, long long outer, long long ld, int box_rows);
int gemm_bf16(const EdbGemmDesc& g, cudaStream_t stream);
int gemm_set_mode(int mode);

int layernorm_fwd(const float* x, long long ldx, const float* gamma, const float* beta, float eps, void* y,
                  long long ldy, int y_f32, float* mean, float* rstd, int rows, int dim, const int* rows_dev,
                  cudaStream_t st);
size_t layernorm_bwd_workspace_bytes();
int layernorm_bwd(const void* dy, long long lddy, int dy_f32, const float* x, long long ldx, const float* mean,
                  const float* rstd, const float* gamma, const float* g_in, float* g_out, long long ldg, void* g_bf16,
                  long long ldgb, float* dgamma, float* dbeta, float* dcol, void* workspace, size_t ws_bytes, int rows,
                  int dim, const float* row_scale, int scale_group, const int* rows_dev, cudaStream_t st);
int colsum(const void* src, long long ld, int src_f32, int rows, int N, float* out, cudaStream_t st);
int cast_f32_bf16(const float* src, void* dst, size_t n, cudaStream_t st);
int cast_rows_f32_bf16(const float* src, void* dst, int max_rows, int cols, const int* rows_dev, cudaStream_t st);
int zero_rows(void* base, long long row_bytes, const int* rows_dev, int nrows, cudaStream_t st);
int split_bf16x3(const float* src, long long ld, int rows, int K, void* dst, int role, cudaStream_t st);
int patch_im2col(const float* rgb, const float* ni, const float* ti, int B, int H, int W, void* out, long long ldo,
                 int out_f32, cudaStream_t st);
int embed_assemble(const float* patch_out, const float* cls, const float* pos, const float* sie, const long long* cam,
                   float coe, int S, int B, int P, float* x, cudaStream_t st);
int embed_assemble_bwd(const float* g, int S, int B, int P, const long long* cam, float coe, float* dpos, float* dsie,
                       void* dpatch, int dpatch_f32, cudaStream_t st);
int gelu_bwd_f32(const float* dh, const float* pre, float* out, size_t n, cudaStream_t st);
int sgd_step(float* p, const float* g, float* buf, void* p16, const unsigned char* flags, size_t n, float lr, float mu,
             float wd, float wd_bias, float bias_lr_factor, float gscale, int first, cudaStream_t st);
int attention_simple(const EdbAttnDesc& d, bool bwd, cudaStream_t st);
int attention_var(const EdbAttnDesc& d, bool bwd, cudaStream_t st);
int attention_tc_fwd(const EdbAttnDesc& d, cudaStream_t st);
int attention_tc_bwd(const EdbAttnDesc& d, cudaStream_t st);
int freq_counts(const float* rgb, const float* ni, const float* ti, int B, int H, int W, int* counts, cudaStream_t st);
int topk_mask(const void* vals, int vals_f32, long long ld, int rows, int n, int k, unsigned* mask, int accumulate,
              cudaStream_t st);
int rollout_topk(const void* const* maps, int layers, int maps_f32, int nseq, int B, int heads, long long p_rows,
                 long long ldp, int k, unsigned* index, unsigned* mod_mask, float* rows_out, cudaStream_t st);
int index_finalize(const unsigned* index, int B, int* seq_off, int* seq_off3, cudaStream_t st);
int sfts_pack_fwd(const float* tokens, const unsigned* index, const int* seq_off, int B, long long cap, float* packed,
                  float* loss_bcc, cudaStream_t st);
int sfts_pack_bwd(const float* tokens, const unsigned* index, const int* seq_off, int B, long long cap,
                  const float* d_packed, const float* g_loss, float* d_tokens, cudaStream_t st);
int joint_gather(float* mod, long long cap, float* joint, const int* seq_off, int B, int max_len, int dir, cudaStream_t st);
int pool_fwd(const float* x, const int* seq_off, int B, float* cls_out, float* patch_mean, int* num, cudaStream_t st);
int pool_bwd(const float* d_cls, const float* d_patch, const int* seq_off, const int* num, int B, int max_len, float* dx,
             cudaStream_t st);
int cls_rows(float* packed, long long cap, const int* seq_off, int B, float* rows, int dir, cudaStream_t st);

int bn1d_fwd(const float* x, long long ldx, int B, int F, const float* gamma, const float* beta, float* run_mean,
             float* run_var, float momentum, float eps, float* y, long long ldy, float* save_mean, float* save_invstd,
             cudaStream_t st);
int bn1d_bwd(const float* dy, long long lddy, const float* x, long long ldx, int B, int F, const float* gamma,
             const float* save_mean, const float* save_invstd, float* dx, long long lddx, float* dgamma, float* dbeta,
             cudaStream_t st);
int ocfr_fwd(const float* x, const long long* label, int B, int C, float* c0, float* c1, float* c2, float mom, float* fn,
             float* inv_norm, float* loss, cudaStream_t st);
int ocfr_bwd(const float* fn, const float* inv_norm, const long long* label, int B, float* c0, float* c1, float* c2,
             const float* g_loss, float* dx, cudaStream_t st);
int ce_smooth(const float* logits, long long ld, const long long* label, int B, int C, float eps, float* loss,
              float* dlogits, long long ldd, cudaStream_t st);
size_t triplet_workspace_bytes(int B);
int triplet_fwd(const float* x, long long ld, const long long* label, int B, int F, float* loss, void* workspace,
                size_t ws_bytes, cudaStream_t st);
int triplet_bwd(const float* x, long long ld, int B, int F, const void* workspace, const float* g, float* dx,
                long long ldd, int accumulate, cudaStream_t st);
int scale_by(const float* x, const float* a, float* y, size_t n, cudaStream_t st);

int eval_normalize(float* feats, long long ld, int N, int F, float eps, cudaStream_t st);
int eval_distmat(const float* qf, long long ldq, int Q, const float* gf, long long ldg, int G, int F, float* dist,
                 long long ldd, cudaStream_t st);
int eval_rank(const float* dist, long long ldd, int Q, int G, const long long* q_pid, const long long* g_pid,
              const long long* q_key, const long long* g_key, double* ap, int* first_rank, int* overflow, cudaStream_t st);

size_t augment_workspace_bytes(int B, int Hs, int Ws, int W);
int augment_u8(const uint8_t* s0, const uint8_t* s1, const uint8_t* s2, int B, int Hs, int Ws, int H, int W, int pad,
               const int* hb, const int* hk, int ksh, const int* vb, const int* vk, int ksv, const float* mean,
               const float* stdv, const EdbAugImage* params, const float* noise, float* o0, float* o1, float* o2,
               void* workspace, size_t ws_bytes, cudaStream_t st);

}  // namespace edb

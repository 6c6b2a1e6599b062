// extern "C" surface of libeditor_b200.so (see include/editor_b200.h).
#include "abi_internal.h"
#include <string.h>
#include <stdio.h>

namespace edb {
static thread_local char g_err[512] = "";
int edb_set_error(int code, const char* msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg ? msg : "");
    return code;
}
}  // namespace edb

extern "C" {

int edb_version(void) { return 100; }
const char* edb_last_error(void) { return edb::g_err; }

int edb_gemm_bf16(const EdbGemmDesc* d, void* stream) {
    if (d == nullptr) return edb::edb_set_error(EDB_ERR_SHAPE, "null descriptor");
    return edb::gemm_bf16(*d, static_cast<cudaStream_t>(stream));
}

int edb_gemm_set_mode(int mode) { return edb::gemm_set_mode(mode); }

#define ST static_cast<cudaStream_t>(stream)

int edb_layernorm_fwd(const float* x, long long ldx, const float* gamma, const float* beta, float eps, void* y,
                      long long ldy, int y_f32, float* mean, float* rstd, int rows, int dim, const int* rows_dev,
                      void* stream) {
    return edb::layernorm_fwd(x, ldx, gamma, beta, eps, y, ldy, y_f32, mean, rstd, rows, dim, rows_dev, ST);
}
size_t edb_layernorm_bwd_workspace_bytes(void) { return edb::layernorm_bwd_workspace_bytes(); }
int edb_layernorm_bwd(const void* dy, long long lddy, int dy_f32, const float* x, long long ldx, const float* mean,
                      const float* rstd, const float* gamma, const float* g_in, float* g_out, long long ldg,
                      void* g_bf16, long long ldgb, float* dgamma, float* dbeta, float* dcol, void* workspace,
                      size_t ws_bytes, int rows, int dim, const float* row_scale, int scale_group, const int* rows_dev,
                      void* stream) {
    return edb::layernorm_bwd(dy, lddy, dy_f32, x, ldx, mean, rstd, gamma, g_in, g_out, ldg, g_bf16, ldgb, dgamma, dbeta,
                              dcol, workspace, ws_bytes, rows, dim, row_scale, scale_group, rows_dev, ST);
}
int edb_colsum(const void* src, long long ld, int src_f32, int rows, int n, float* out, void* stream) {
    return edb::colsum(src, ld, src_f32, rows, n, out, ST);
}
int edb_cast_f32_bf16(const float* src, void* dst, size_t n, void* stream) { return edb::cast_f32_bf16(src, dst, n, ST); }
int edb_cast_rows_f32_bf16(const float* src, void* dst, int max_rows, int cols, const int* rows_dev, void* stream) {
    return edb::cast_rows_f32_bf16(src, dst, max_rows, cols, rows_dev, ST);
}
int edb_zero_rows(void* base, long long row_bytes, const int* rows_dev, int nrows, void* stream) {
    return edb::zero_rows(base, row_bytes, rows_dev, nrows, ST);
}
int edb_sgd_step(float* p, const float* g, float* buf, void* p16, const unsigned char* flags, size_t n, float lr,
                 float momentum, float wd, float wd_bias, float bias_lr_factor, float gscale, int first, void* stream) {
    return edb::sgd_step(p, g, buf, p16, flags, n, lr, momentum, wd, wd_bias, bias_lr_factor, gscale, first, ST);
}
int edb_split_bf16x3(const float* src, long long ld, int rows, int K, void* dst, int role, void* stream) {
    return edb::split_bf16x3(src, ld, rows, K, dst, role, ST);
}
int edb_patch_im2col(const float* rgb, const float* ni, const float* ti, int B, int H, int W, void* out, long long ldo,
                     int out_f32, void* stream) {
    return edb::patch_im2col(rgb, ni, ti, B, H, W, out, ldo, out_f32, ST);
}
int edb_embed_assemble(const float* patch_out, const float* cls, const float* pos, const float* sie,
                       const long long* cam, float coe, int S, int B, int P, float* x, void* stream) {
    return edb::embed_assemble(patch_out, cls, pos, sie, cam, coe, S, B, P, x, ST);
}
int edb_embed_assemble_bwd(const float* g, int S, int B, int P, const long long* cam, float coe, float* dpos,
                           float* dsie, void* dpatch, int dpatch_f32, void* stream) {
    return edb::embed_assemble_bwd(g, S, B, P, cam, coe, dpos, dsie, dpatch, dpatch_f32, ST);
}
int edb_gelu_bwd_f32(const float* dh, const float* pre, float* out, size_t n, void* stream) {
    return edb::gelu_bwd_f32(dh, pre, out, n, ST);
}
static bool tc_eligible(const EdbAttnDesc* d) {
    return d->impl == 0 && !d->f32 && d->seq_off == nullptr && d->fixed_len == 129 && d->heads == 12 && d->P != nullptr &&
           d->p_rows == 129 && d->ldp == 136;
}
int edb_attention_fwd(const EdbAttnDesc* d, void* stream) {
    if (d == nullptr) return edb::edb_set_error(EDB_ERR_SHAPE, "null descriptor");
    if (d->impl == 2) return edb::attention_var(*d, false, ST);
    if (tc_eligible(d)) return edb::attention_tc_fwd(*d, ST);
    return edb::attention_simple(*d, false, ST);
}
int edb_attention_bwd(const EdbAttnDesc* d, void* stream) {
    if (d == nullptr) return edb::edb_set_error(EDB_ERR_SHAPE, "null descriptor");
    if (d->impl == 2) return edb::attention_var(*d, true, ST);
    if (tc_eligible(d)) return edb::attention_tc_bwd(*d, ST);
    return edb::attention_simple(*d, true, ST);
}
int edb_freq_counts(const float* rgb, const float* ni, const float* ti, int B, int H, int W, int* counts, void* stream) {
    return edb::freq_counts(rgb, ni, ti, B, H, W, counts, ST);
}
int edb_topk_mask(const void* vals, int vals_f32, long long ld, int rows, int n, int k, unsigned* mask, int accumulate,
                  void* stream) {
    return edb::topk_mask(vals, vals_f32, ld, rows, n, k, mask, accumulate, ST);
}
int edb_rollout_topk(const void* const* maps_host, int layers, int maps_f32, int nseq, int B, int heads,
                     long long p_rows, long long ldp, int k, unsigned* index, unsigned* mod_mask, float* rows_out,
                     void* stream) {
    return edb::rollout_topk(maps_host, layers, maps_f32, nseq, B, heads, p_rows, ldp, k, index, mod_mask, rows_out, ST);
}
int edb_index_finalize(const unsigned* index, int B, int* seq_off, int* seq_off3, void* stream) {
    return edb::index_finalize(index, B, seq_off, seq_off3, ST);
}
int edb_sfts_pack_fwd(const float* tokens, const unsigned* index, const int* seq_off, int B, long long cap,
                      float* packed, float* loss_bcc, void* stream) {
    return edb::sfts_pack_fwd(tokens, index, seq_off, B, cap, packed, loss_bcc, ST);
}
int edb_sfts_pack_bwd(const float* tokens, const unsigned* index, const int* seq_off, int B, long long cap,
                      const float* d_packed, const float* g_loss, float* d_tokens, void* stream) {
    return edb::sfts_pack_bwd(tokens, index, seq_off, B, cap, d_packed, g_loss, d_tokens, ST);
}
int edb_joint_gather(float* mod, long long cap, float* joint, const int* seq_off, int B, int max_len, int dir,
                     void* stream) {
    return edb::joint_gather(mod, cap, joint, seq_off, B, max_len, dir, ST);
}
int edb_pool_fwd(const float* x, const int* seq_off, int B, float* cls_out, float* patch_mean, int* num, void* stream) {
    return edb::pool_fwd(x, seq_off, B, cls_out, patch_mean, num, ST);
}
int edb_pool_bwd(const float* d_cls, const float* d_patch, const int* seq_off, const int* num, int B, int max_len,
                 float* dx, void* stream) {
    return edb::pool_bwd(d_cls, d_patch, seq_off, num, B, max_len, dx, ST);
}
int edb_cls_rows(float* packed, long long cap, const int* seq_off, int B, float* rows, int dir, void* stream) {
    return edb::cls_rows(packed, cap, seq_off, B, rows, dir, ST);
}

int edb_bn1d_fwd(const float* x, long long ldx, int B, int F, const float* gamma, const float* beta, float* run_mean,
                 float* run_var, float momentum, float eps, float* y, long long ldy, float* save_mean,
                 float* save_invstd, void* stream) {
    return edb::bn1d_fwd(x, ldx, B, F, gamma, beta, run_mean, run_var, momentum, eps, y, ldy, save_mean, save_invstd, ST);
}
int edb_bn1d_bwd(const float* dy, long long lddy, const float* x, long long ldx, int B, int F, const float* gamma,
                 const float* save_mean, const float* save_invstd, float* dx, long long lddx, float* dgamma,
                 float* dbeta, void* stream) {
    return edb::bn1d_bwd(dy, lddy, x, ldx, B, F, gamma, save_mean, save_invstd, dx, lddx, dgamma, dbeta, ST);
}
int edb_ocfr_fwd(const float* x, const long long* label, int B, int C, float* c_rgb, float* c_nir, float* c_tir,
                 float momentum, float* fn, float* inv_norm, float* loss, void* stream) {
    return edb::ocfr_fwd(x, label, B, C, c_rgb, c_nir, c_tir, momentum, fn, inv_norm, loss, ST);
}
int edb_ocfr_bwd(const float* fn, const float* inv_norm, const long long* label, int B, float* c_rgb, float* c_nir,
                 float* c_tir, const float* g_loss, float* dx, void* stream) {
    return edb::ocfr_bwd(fn, inv_norm, label, B, c_rgb, c_nir, c_tir, g_loss, dx, ST);
}
int edb_ce_smooth(const float* logits, long long ld, const long long* label, int B, int C, float eps, float* loss,
                  float* dlogits, long long ldd, void* stream) {
    return edb::ce_smooth(logits, ld, label, B, C, eps, loss, dlogits, ldd, ST);
}
size_t edb_triplet_workspace_bytes(int B) { return edb::triplet_workspace_bytes(B); }
int edb_triplet_fwd(const float* x, long long ld, const long long* label, int B, int F, float* loss, void* workspace,
                    size_t ws_bytes, void* stream) {
    return edb::triplet_fwd(x, ld, label, B, F, loss, workspace, ws_bytes, ST);
}
int edb_triplet_bwd(const float* x, long long ld, int B, int F, const void* workspace, const float* g_loss, float* dx,
                    long long ldd, int accumulate, void* stream) {
    return edb::triplet_bwd(x, ld, B, F, workspace, g_loss, dx, ldd, accumulate, ST);
}
int edb_scale_by(const float* x, const float* a, float* y, size_t n, void* stream) { return edb::scale_by(x, a, y, n, ST); }


int edb_eval_normalize(float* feats, long long ld, int n, int f, float eps, void* stream) {
    return edb::eval_normalize(feats, ld, n, f, eps, ST);
}
int edb_eval_distmat(const float* qf, long long ldq, int q, const float* gf, long long ldg, int g, int f, float* dist,
                     long long ldd, void* stream) {
    return edb::eval_distmat(qf, ldq, q, gf, ldg, g, f, dist, ldd, ST);
}
int edb_eval_rank(const float* dist, long long ldd, int q, int g, const long long* q_pid, const long long* g_pid,
                  const long long* q_key, const long long* g_key, double* ap, int* first_rank, int* overflow, void* stream) {
    return edb::eval_rank(dist, ldd, q, g, q_pid, g_pid, q_key, g_key, ap, first_rank, overflow, ST);
}
size_t edb_augment_workspace_bytes(int B, int Hs, int Ws, int W) { return edb::augment_workspace_bytes(B, Hs, Ws, W); }
int edb_augment_u8(const unsigned char* src_rgb, const unsigned char* src_ni, const unsigned char* src_ti, int B, int Hs,
                   int Ws, int H, int W, int pad, const int* hb, const int* hk, int ksh, const int* vb, const int* vk,
                   int ksv, const float* mean, const float* std, const EdbAugImage* params, const float* noise,
                   float* out_rgb, float* out_ni, float* out_ti, void* workspace, size_t ws_bytes, void* stream) {
    return edb::augment_u8(src_rgb, src_ni, src_ti, B, Hs, Ws, H, W, pad, hb, hk, ksh, vb, vk, ksv, mean, std, params, noise,
                           out_rgb, out_ni, out_ti, workspace, ws_bytes, ST);
}
}  // extern "C"

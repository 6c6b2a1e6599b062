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
#define EPI_GELU 1     /* out2 = acc + bias (optional), D = gelu_erf(acc + bias)    (vit_pytorch.py:139-145) */
#define EPI_RESIDUAL 2 /* D(f32) = aux(f32) + acc + bias; D may alias aux           (vit_pytorch.py:217-219) */
#define EPI_GELU_BWD 3 /* D = acc * gelu'(aux)                                                      */
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
} EdbGemmDesc;

int edb_gemm_bf16(const EdbGemmDesc* desc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EDITOR_B200_H */

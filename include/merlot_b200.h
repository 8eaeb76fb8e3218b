/*
 * merlot_b200 -- C-ABI of the B200-native (sm_100a) implementation of MERLOT's dense forward/backward hot path.
 *
 * The reference (rowanz/merlot) has no FFI/operator layer: its boundary is the Python class `MerlotModel`
 * (model/modeling.py:47-668) plus `optimization.build_optimizer_from_config` (utils/optimization.py:11-30), both of
 * which only call stock TF ops.  This header is the operator layer a maintainer would bind instead of those TF ops;
 * every entry point cites the reference call site(s) whose arithmetic it replaces.  The Python mirror of the
 * reference surface lives in merlot_b200/modeling.py and binds this header with ctypes.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer owned by the caller (PyTorch allocates);
 *    the library allocates nothing persistent.
 *  - `stream` is a cudaStream_t passed as void*; all work is stream-ordered and asynchronous.
 *  - return value: MERLOT_OK (0) or a negative MERLOT_E* code; merlot_last_error() gives a thread-local message.
 *  - matrices are row-major; `ld*` are leading dimensions in ELEMENTS; bf16 = 16-bit brain float, f32 = IEEE float.
 *  - "tokens" M = batch*seq rows of the flattened [M, H] residual stream, as in utils/transformer.py:185.
 */
#ifndef MERLOT_B200_H_
#define MERLOT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MERLOT_OK 0
#define MERLOT_EINVAL (-1)   /* bad argument / null pointer / unsupported flag combination */
#define MERLOT_ESHAPE (-2)   /* shape or alignment constraint violated (mirrors the reference's ValueError/assert) */
#define MERLOT_ECUDA (-3)    /* CUDA runtime / driver error */
#define MERLOT_ENOTIMPL (-4) /* config key accepted by the reference but not yet provided here (raised loudly) */

const char* merlot_last_error(void);
int merlot_abi_version(void);
/* number of kernels this library has launched since the last reset (bench.py reports it as gpu_launches) */
long long merlot_launch_count(void);
void merlot_reset_launch_count(void);

/* ------------------------------------------------------------------------------------------------------------
 * K1: bf16 tensor-core GEMM (tcgen05.mma, TMA-fed, fp32 accumulation in TMEM) with fused epilogues.
 *     C[M,N] = epilogue( alpha * sum_k A(m,k) * B(n,k) )
 * Replaces every tf.layers.dense / tf.matmul on the hot path and their tf.gradients:
 *   utils/transformer.py:21-25 (q/k/v), :130-135 (context_projection_layer), :149-155 (intermediate + gelu),
 *   :157-161 (output); utils/vision_transformer.py:196-205 (patch-embed conv as im2col GEMM);
 *   model/modeling.py:28-42 (project_and_norm), :208-217 (lm_head + tied logits), :521 (contrastive logits),
 *   :582-595 (temporal head).
 * Operand storage:
 *   a_mn_major = 0 : A is stored [M][K] (lda >= K)     a_mn_major = 1 : A is stored [K][M] (lda >= M)
 *   b_mn_major = 0 : B is stored [N][K] (ldb >= K)     b_mn_major = 1 : B is stored [K][N] (ldb >= N)
 *   (a TF `kernel` [in,out] used in the forward pass is B with b_mn_major=1; the same buffer is the K-major B of the
 *    dgrad GEMM; wgrad uses both activations MN-major.)  lda/ldb must be multiples of 8 elements, bases 16B-aligned.
 * Epilogue, applied in this order per element (row m, column n):
 *   v = alpha*acc; v += bias[n]; if GELU: {pre=v; v=gelu_erf(v)}; if MUL_DGELU: v *= gelu_erf'(aux[m,n]);
 *   if DROPOUT: v = keep(m,n) ? v/(1-p) : 0; v += resid[m,n]; store.
 *   With GELU and out2 != NULL: out <- pre (bf16), out2 <- gelu(pre).
 *   ATOMIC: fp32 red.add into out (split-K wgrad accumulation; out must be pre-zeroed or hold the running sum).
 * ------------------------------------------------------------------------------------------------------------ */
#define MERLOT_GEMM_OUT_F32 1u
#define MERLOT_GEMM_ATOMIC 2u
#define MERLOT_GEMM_GELU 4u
#define MERLOT_GEMM_MUL_DGELU 8u
#define MERLOT_GEMM_DROPOUT 16u

typedef struct merlot_gemm {
  int M, N, K;
  const void* a; int lda; int a_mn_major;
  const void* b; int ldb; int b_mn_major;
  void* out; int ld_out;          /* bf16 unless MERLOT_GEMM_OUT_F32 */
  void* out2; int ld_out2;        /* optional second bf16 output (post-GELU) */
  const float* bias;              /* [N] fp32 or NULL */
  const void* resid; int ld_resid;/* bf16 [M,N] or NULL */
  const void* aux; int ld_aux;    /* bf16 [M,N] pre-activation for MUL_DGELU */
  float alpha;
  uint32_t flags;
  float dropout_p; uint64_t dropout_seed; uint32_t dropout_site;
  int splits;                     /* 0 = auto; >1 requires MERLOT_GEMM_ATOMIC */
  int block_n;                    /* 0 = auto; else 128 or 256 (tuning / tests) */
} merlot_gemm_t;

int merlot_gemm_bf16(const merlot_gemm_t* g, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MERLOT_B200_H_ */

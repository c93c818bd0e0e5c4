/*
 * multimae_b200 — C ABI of the B200-native (sm_100a) MultiMAE pre-training hot path.
 *
 * The reference (EPFL-VILAB/MultiMAE) is pure Python on top of PyTorch: it has no FFI layer of its own.
 * The "plugin boundary" of the hot path is the Python module API consumed by run_pretraining_multimae.py
 * (SURVEY.md §8b).  This header is what the Python host side (multimae_b200/*.py, loaded with ctypes)
 * binds; every entry point names the reference code it replaces (path:line in /root/reference).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; no PyTorch types cross this ABI;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - functions return 0 on success, non-zero on error; mmae_last_error() returns a message for the
 *     calling thread's last failure;  nothing here falls back to a CPU path;
 *   - matrices are row-major; "bf16" is __nv_bfloat16; leading dimensions are in elements;
 *   - all leading dimensions must be multiples of 8 elements and base pointers 16-byte aligned.
 */
#ifndef MULTIMAE_B200_H_
#define MULTIMAE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMAE_ABI_VERSION 1

int mmae_abi_version(void);
const char* mmae_last_error(void);
/* number of kernel launches issued through this library by the calling process so far */
int64_t mmae_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * GEMM on the tcgen05 tensor cores (bf16 x bf16 -> fp32 accumulate in TMEM), TMA-fed.
 * Replaces every nn.Linear / nn.Conv2d(k=s=P) matmul of the path:
 *   multimae/multimae_utils.py:149,153,172,180,203-204,212 ; multimae/input_adapters.py:110,232 ;
 *   multimae/output_adapters.py:258,274 and their autograd backward (dgrad / wgrad).
 *
 *   C[M,N] = epilogue( alpha * sum_k A(m,k) * B(n,k) )
 *   a_mn_major = 0 : A is stored [M,K] (K contiguous),  lda = row pitch
 *   a_mn_major = 1 : A is stored [K,M] (M contiguous),  lda = row pitch      (used by wgrad)
 *   b_mn_major = 0 : B is stored [N,K] (K contiguous)   (nn.Linear weight as-is: forward)
 *   b_mn_major = 1 : B is stored [K,N] (N contiguous)   (nn.Linear weight as-is: dgrad; wgrad input)
 *
 * Epilogue, in this order (all optional):
 *   v = alpha*acc ; v += bias[n] ; preact_bf16[m,n] = v ; v = act(v) ; v *= gelu'(dgelu_z[m,n]) ;
 *   v += residual[m,n] ; out_f32[m,n] (=|+=) v ; out_bf16[m,n] = v
 * split_k > 1 (or accumulate != 0) accumulates into out_f32 with fp32 atomics; the destination must have
 * been zeroed (or hold the value to accumulate onto); bias/residual are applied by split 0 only; act,
 * dgelu_z, preact_bf16 and out_bf16 are not allowed with split_k > 1.
 * ---------------------------------------------------------------------------------------------- */
typedef struct mmae_gemm_epilogue {
  float alpha;
  int act;                 /* 0 = identity, 1 = exact-erf GELU (nn.GELU default) */
  int accumulate;          /* 1: out_f32 += v (atomic) */
  int reserved;
  const float* bias;       /* [N] fp32 or NULL */
  const float* residual;   /* [M, ld_residual] fp32 or NULL */
  const void* dgelu_z;     /* bf16 [M, ld_dgelu_z] or NULL */
  void* preact_bf16;       /* bf16 [M, ld_preact] or NULL */
  float* out_f32;          /* [M, ld_out_f32] or NULL */
  void* out_bf16;          /* bf16 [M, ld_out_bf16] or NULL */
  int64_t ld_residual, ld_dgelu_z, ld_preact, ld_out_f32, ld_out_bf16;
} mmae_gemm_epilogue;

int mmae_gemm_bf16(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb, int b_mn_major,
                   int M, int N, int K, int split_k, const mmae_gemm_epilogue* ep, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Element-wise / layout helpers (HBM-bound)
 * ---------------------------------------------------------------------------------------------- */
/* dst_bf16[i] = bf16(src[i])            (autocast's fp32->half cast of weights/activations) */
int mmae_cast_f32_to_bf16(const float* src, void* dst_bf16, int64_t n, void* stream);
/* dst[M,N] bf16 = bf16(src[M,N] fp32) and colsum[n] += sum_m src[m,n]  (bias gradient: autograd of
 * nn.Linear bias, multimae/multimae_utils.py:143-145,165-167); colsum may be NULL; dst may be NULL */
int mmae_cast_colsum_f32(const float* src, int64_t ld_src, void* dst_bf16, int64_t ld_dst, float* colsum,
                         int M, int N, void* stream);
/* colsum[n] += sum_m src_bf16[m,n] */
int mmae_colsum_bf16(const void* src_bf16, int64_t ld_src, float* colsum, int M, int N, void* stream);
/* dst[N,M] = src[M,N]^T (bf16) */
int mmae_transpose_bf16(const void* src, int64_t ld_src, void* dst, int64_t ld_dst, int M, int N, void* stream);

/* ------------------------------------------------------------------------------------------------
 * LayerNorm (fp32 statistics; eps given).  nn.LayerNorm(eps=1e-6): multimae/multimae_utils.py:222,225 and
 * multimae/output_adapters.py:120-122.  D must be a multiple of 128, <= 1024.
 *   forward : y = (x-mean)*rstd*gamma+beta, written as bf16 (GEMM operand) and/or fp32; saves mean/rstd.
 *   backward: dx = dx_resid(optional) + LN'(dy); dgamma/dbeta are ACCUMULATED (+=) with fp32 atomics.
 * ---------------------------------------------------------------------------------------------- */
int mmae_layernorm_forward(const float* x, int64_t ldx, const float* gamma, const float* beta, void* y_bf16,
                           int64_t ldy, float* y_f32, int64_t ldyf, float* mean, float* rstd, int M, int D,
                           float eps, void* stream);
int mmae_layernorm_backward(const void* dy, int dy_is_bf16, int64_t lddy, const float* x, int64_t ldx,
                            const float* mean, const float* rstd, const float* gamma, const float* dx_resid,
                            int64_t ldr, float* dx, int64_t lddx, float* dgamma, float* dbeta, int M, int D,
                            void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused multi-head attention (scores stay on chip).  Attention / CrossAttention:
 * multimae/multimae_utils.py:172-179, 203-211.  q/k/v/o are bf16 views into row-major [B*N, ld] projection
 * buffers: head h = columns [h*head_dim, (h+1)*head_dim) from the given base pointer, batch b = rows
 * [b*N, (b+1)*N).  head_dim in {32, 64}.  lse[B,H,Nq] = log-sum-exp of the scaled scores (saved for backward).
 * backward: delta_ws is a [B,H,Nq] fp32 scratch; dq/dk/dv are written (not accumulated).
 * ---------------------------------------------------------------------------------------------- */
int mmae_attention_forward(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                           void* o, int64_t ldo, float* lse, int B, int H, int Nq, int Nk, int head_dim,
                           float scale, void* stream);
int mmae_attention_backward(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                            const void* o, int64_t ldo, const void* d_o, int64_t lddo, const float* lse,
                            float* delta_ws, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv,
                            int64_t lddv, int B, int H, int Nq, int Nk, int head_dim, float scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MULTIMAE_B200_H_ */

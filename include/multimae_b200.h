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

#define MMAE_ABI_VERSION 4

int mmae_abi_version(void);
const char* mmae_last_error(void);
/* number of kernel launches issued through this library by the calling process so far */
int64_t mmae_launch_count(void);

/* Optional device timing of every GEMM launch (cudaEvent pair on the launch stream); used by bench.py to report the
 * achieved TFLOP/s of the dominant kernel.  mmae_profile_gemm(1) starts/resets, (0) stops; _read synchronises the
 * recorded events and returns the summed algorithmic FLOPs (2*M*N*K), summed kernel milliseconds and launch count. */
int mmae_profile_gemm(int enable);
int mmae_profile_gemm_read(double* flops, double* ms, int64_t* launches);
/* text dump, one line per recorded launch: "M N K flags ms" (flags: bit0 A MN-major, bit1 B MN-major, bits 8+ split_k) */
int64_t mmae_profile_gemm_dump(char* buf_host, int64_t capacity);

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
 * split_k <= 0: automatic - split count and tile width are chosen together so that one wave of ~SM-count work items
 * covers the problem (fp32-only linear epilogues; anything else runs unsplit).  The accumulation then uses TMA
 * reduce-add tiles (cp.reduce.async.bulk.tensor) instead of per-lane atomics.
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

/* Kernel variant selection for measurements: -1 = heuristic (default), 0 = one-tile-per-CTA 128x128,
 * 1 = persistent 128x128 with double-buffered TMEM, 2 = persistent 128x256, 3 = persistent 128x192,
 * 4 / 5 / 6 = CTA-pair kernels (cluster of 2, tcgen05 cta_group::2) with 256x256 / 256x192 / 256x128 tiles.
 * Env MMAE_GEMM_VARIANT sets the initial value; MMAE_GEMM_PAIR=0 keeps the heuristic off the pair kernels. */
int mmae_gemm_set_variant(int variant);
/* 1 (default): bf16-only epilogues of the persistent kernels leave through shared memory + TMA tile stores;
 * 0: per-lane global stores (kept for A/B measurements and for epilogues with extra operands).  Env MMAE_GEMM_TMA_STORE. */
int mmae_gemm_set_tma_store(int enable);
/* 1: kernels are launched with programmatic stream serialization and start with griddepcontrol (launch_dependents +
 * wait), so the next kernel's blocks are scheduled while the previous one drains; 0 (default): plain stream order - a
 * programmatically launched dependent keeps stale L1 lines for non-coherent loads (see runtime.cu).  Env MMAE_PDL. */
int mmae_set_pdl(int enable);
/* SM budget of the persistent kernels (GEMM, warp-specialised attention, one-wave element-wise grids): 0 = every SM
 * (default), n = at most n.  Data-parallel runs leave NCCL's all-reduce CTAs their SMs.  Env MMAE_SM_BUDGET. */
int mmae_set_sm_budget(int sms);
/* 1: mmae_block_backward runs its four weight-gradient GEMMs on a library-owned side stream, forked behind the kernel
 * that produces their dY operand and joined before the call returns (the caller's stream order is unchanged);
 * 0 (default; no gain measured at the MultiMAE-B shapes): everything on the caller's stream.  Env MMAE_WGRAD_STREAM. */
int mmae_set_wgrad_stream(int enable);

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
/* exact-erf GELU over a bf16 stream (nn.GELU, multimae/multimae_utils.py:139,150): backward = 0: io[i] = gelu(z[i]);
 * backward = 1: io[i] *= gelu'(z[i]) in place.  n must be a multiple of 8. */
int mmae_gelu_bf16(const void* z, void* io, int64_t n, int backward, void* stream);
/* dz[m,n] *= gelu'(z[m,n]) in place and colsum[n] += sum_m dz[m,n] (fc1 bias gradient) in one pass */
int mmae_dgelu_colsum_bf16(const void* z, void* dz, int64_t ld, float* colsum, int M, int N, void* stream);
/* 1: the module-level entry points fuse GELU / GELU' into the GEMM epilogue; 0 (default): streaming kernel after the
 * GEMM (measured faster: the epilogue is instruction-issue bound).  Env MMAE_FUSE_GELU sets the initial value. */
int mmae_set_fuse_gelu(int enable);
/* out[i] = x[i] + float(y_bf16[i]); n must be a multiple of 8 */
int mmae_add_bf16_f32(const float* x, const void* y_bf16, float* out, int64_t n, void* stream);
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
/* x_sum = x + addend_bf16 (fp32, written when non-NULL) followed by LayerNorm(x_sum) -> y_bf16: the residual add
 * `x = x + branch(...)` (multimae/multimae_utils.py:230-231) fused in front of the next norm. */
int mmae_add_layernorm_forward(const float* x, int64_t ldx, const void* addend_bf16, int64_t ldadd, float* x_sum,
                               int64_t ldsum, const float* gamma, const float* beta, void* y_bf16, int64_t ldy,
                               float* mean, float* rstd, int M, int D, float eps, void* stream);
int mmae_layernorm_backward(const void* dy, int dy_is_bf16, int64_t lddy, const float* x, int64_t ldx,
                            const float* mean, const float* rstd, const float* gamma, const float* dx_resid,
                            int64_t ldr, float* dx, int64_t lddx, float* dgamma, float* dbeta, int M, int D,
                            void* stream);

/* as mmae_layernorm_backward, additionally writing bf16(dx) and accumulating dx_colsum[d] += sum_rows dx[:,d] */
int mmae_layernorm_backward_ex(const void* dy, int dy_is_bf16, int64_t lddy, const float* x, int64_t ldx,
                               const float* mean, const float* rstd, const float* gamma, const float* dx_resid,
                               int64_t ldr, float* dx, int64_t lddx, float* dgamma, float* dbeta, void* dx_bf16,
                               int64_t lddxb, float* dx_colsum, int M, int D, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused multi-head attention (scores stay on chip).  Attention / CrossAttention:
 * multimae/multimae_utils.py:172-179, 203-211.  q/k/v/o are bf16 views into row-major [B*N, ld] projection
 * buffers: head h = columns [h*head_dim, (h+1)*head_dim) from the given base pointer, batch b = rows
 * [b*N, (b+1)*N).  head_dim in {32, 64}.  lse[B,H,Nq] = log-sum-exp of the scaled scores (saved for backward).
 * backward: delta_ws is a [B,H,Nq] fp32 scratch; dq/dk/dv are written (not accumulated).
 * ---------------------------------------------------------------------------------------------- */
/* bit mask: 1 = fused single-tile tcgen05 kernels (Nq, Nk <= 128, head_dim 64), 2 = general tcgen05 forward (head_dim
 * 32/64; used for <= 128 keys, with 8 also for <= 256 keys), 4 = general tcgen05 backward; 0 = warp-MMA (mma.sync +
 * ldmatrix + cp.async) kernels everywhere; 32 / 64 = warp-specialised persistent tcgen05 forward / backward (TMA warp, MMA
 * warp, 8 compute warps; attention_ws.cu) for every shape they support (forward: <= 256 keys; backward: <= 256 queries).
 * 128 = the warp-specialised forward only where the alternative is the mma.sync kernel (129..256 keys).
 * Default 3 | 64 | 128 (env MMAE_ATTN_TC); a negative value restores the start-up default. */
int mmae_attention_set_tc(int enable);
/* diagnostics for the warp-specialised kernels (attention_ws.cu): a device buffer of 64 x 16 int64 that CTA 0 of the next
 * launches fills with clock64 stamps of its pipeline phases (slot 16 * item + phase); NULL switches it off. */
int mmae_attention_ws_set_trace(long long* device_buffer);
int mmae_attention_forward(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                           void* o, int64_t ldo, float* lse, int B, int H, int Nq, int Nk, int head_dim,
                           float scale, void* stream);
int mmae_attention_backward(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                            const void* o, int64_t ldo, const void* d_o, int64_t lddo, const float* lse,
                            float* delta_ws, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv,
                            int64_t lddv, int B, int H, int Nq, int Nk, int head_dim, float scale, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Dirichlet token-mask sampler.  MultiMAE.generate_random_masks, multimae/multimae.py:189-216, as a pure
 * function of the random draws: shares[B,T] (Dirichlet sample), noise_task[B,total] (the per-task
 * torch.rand draws concatenated in task order), noise_all[B,total].  Outputs are int64 like the reference:
 * task_masks[B,total] (0 = visible; split per task by the caller), ids_keep[B,num_encoded],
 * ids_restore[B,total].  Ties break towards the lower index.
 * ---------------------------------------------------------------------------------------------- */
int mmae_sample_masks(const float* shares, const float* noise_task, const float* noise_all, int B, int num_tasks,
                      const int* tokens_per_task_host, int num_encoded, int64_t* task_masks, int64_t* ids_keep,
                      int64_t* ids_restore, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Gather-first patch embedding (PatchedInputAdapter.forward multimae/input_adapters.py:97-119,
 * SemSegInputAdapter.forward :215-241, and the cat/gather/global-token cat of multimae/multimae.py:340-347).
 * Only the num_encoded visible patches per sample are embedded: one K-concatenated GEMM
 * [B*T, sum_t K_t] x [D, sum_t K_t]^T whose A rows are zero outside the token's own task segment.
 * ---------------------------------------------------------------------------------------------- */
#define MMAE_MAX_TASKS 8

typedef struct mmae_embed_layout {
  int num_tasks;
  int grid_h[MMAE_MAX_TASKS], grid_w[MMAE_MAX_TASKS]; /* patch grid of each task at the current image size */
  int tok_offset[MMAE_MAX_TASKS + 1];                 /* token range of each task in the concatenated sequence */
  int k_offset[MMAE_MAX_TASKS + 1];                   /* K segment of each task (K_t = channels * patch^2) */
  int patch[MMAE_MAX_TASKS];                          /* P_H = P_W in input pixels (16, 16, 4) */
  int channels[MMAE_MAX_TASKS];                       /* image channels, or dim_class_emb for semseg */
  int is_semseg[MMAE_MAX_TASKS];
  int num_classes[MMAE_MAX_TASKS];
} mmae_embed_layout;

typedef struct mmae_embed_inputs {
  const void* data[MMAE_MAX_TASKS];       /* fp32 [B,C,H,W]  or  int64 [B,H,W] class ids */
  const float* class_emb[MMAE_MAX_TASKS]; /* [num_classes, dim_class_emb] (semseg) or NULL */
} mmae_embed_inputs;

typedef struct mmae_embed_params {
  const float* weight[MMAE_MAX_TASKS]; /* proj.weight viewed as [D, C*P*P] */
  const float* bias[MMAE_MAX_TASKS];   /* proj.bias [D] */
  const float* pos[MMAE_MAX_TASKS];    /* [N_t, D] rows of the (resized) positional table */
  const float* global_tokens;          /* [G, D] */
} mmae_embed_params;

typedef struct mmae_embed_grads {      /* accumulated (+=) */
  float* weight[MMAE_MAX_TASKS];
  float* bias[MMAE_MAX_TASKS];
  float* class_emb[MMAE_MAX_TASKS];
  float* global_tokens;
} mmae_embed_grads;

int64_t mmae_embed_saved_bytes(const mmae_embed_layout* layout, int B, int T, int D);
int64_t mmae_embed_workspace_bytes(const mmae_embed_layout* layout, int B, int T, int D);
/* x_out: [B, T+G, D] fp32 packed encoder input (global tokens last) */
int mmae_embed_forward(const mmae_embed_layout* layout, const mmae_embed_inputs* in, const mmae_embed_params* prm,
                       const int64_t* ids_keep, int B, int T, int G, int D, float* x_out, void* saved, void* ws,
                       void* stream);
int mmae_embed_backward(const mmae_embed_layout* layout, const mmae_embed_inputs* in, const mmae_embed_params* prm,
                        const mmae_embed_grads* grads, const int64_t* ids_keep, int B, int T, int G, int D,
                        const float* dx, const void* saved, void* ws, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Pre-LN transformer block (Block / Attention / Mlp, multimae/multimae_utils.py:138-182, 217-232); used by the
 * encoder (D=768/1024) and the decoder transformer layers (D=256).  x: [B, N, D] fp32 residual stream.
 * `saved` keeps the bf16 operands needed by backward; sizes from the *_bytes queries.  Gradients of the
 * parameters are ACCUMULATED (+=) into `grads` (zero them once per step).
 * ---------------------------------------------------------------------------------------------- */
typedef struct mmae_block_params {
  const float *norm1_w, *norm1_b, *qkv_w, *qkv_b, *proj_w, *proj_b, *norm2_w, *norm2_b, *fc1_w, *fc1_b, *fc2_w, *fc2_b;
} mmae_block_params;
typedef struct mmae_block_grads {
  float *norm1_w, *norm1_b, *qkv_w, *qkv_b, *proj_w, *proj_b, *norm2_w, *norm2_b, *fc1_w, *fc1_b, *fc2_w, *fc2_b;
} mmae_block_grads;

int64_t mmae_block_saved_bytes(int B, int N, int D, int H, int hidden);
int64_t mmae_block_workspace_bytes(int B, int N, int D, int H, int hidden);
int mmae_block_forward(const float* x_in, float* x_out, int B, int N, int D, int H, int hidden, float eps,
                       const mmae_block_params* prm, void* saved, void* ws, void* stream);
int mmae_block_backward(const float* x_in, const float* dx_out, float* dx_in, int B, int N, int D, int H, int hidden,
                        const mmae_block_params* prm, const mmae_block_grads* grads, const void* saved, void* ws,
                        void* stream);
/* Chained blocks (nn.Sequential of Blocks: the encoder, multimae/multimae.py:349, and each decoder_transformer,
 * multimae/output_adapters.py:271).  The residual add that ends a block, `x = x + mlp(norm2(x))`
 * (multimae/multimae_utils.py:231), is handed to the next block instead of running as a pass of its own:
 *   forward  - y_out_bf16 != NULL: the MLP branch output is written there, x_out is not written, and the block's output is
 *              x_mid + y_out, x_mid = mmae_block_saved_x_mid(saved);  x_add_bf16 != NULL: the block input is
 *              x_in + x_add_bf16, formed inside the first LayerNorm kernel and written to x_sum (fp32) - the x_in to give
 *              the backward call;
 *   backward - dx_in_bf16 != NULL: the first LayerNorm's backward also writes bf16(dx_in) there and adds colsum(dx_in) to
 *              dx_in_colsum (the PREVIOUS block's fc2 bias gradient);  dx_out_bf16 != NULL: that copy, made by the next
 *              block's backward - this block's cast + bias-gradient pass over dx_out is skipped.
 * With all optional pointers NULL the calls equal mmae_block_forward / mmae_block_backward. */
int mmae_block_forward_chain(const float* x_in, const void* x_add_bf16, float* x_sum, float* x_out, void* y_out_bf16,
                             int B, int N, int D, int H, int hidden, float eps, const mmae_block_params* prm, void* saved,
                             void* ws, void* stream);
float* mmae_block_saved_x_mid(void* saved, int B, int N, int D, int H, int hidden);
int mmae_block_backward_chain(const float* x_in, const float* dx_out, const void* dx_out_bf16, float* dx_in,
                              void* dx_in_bf16, float* dx_in_colsum, int B, int N, int D, int H, int hidden,
                              const mmae_block_params* prm, const mmae_block_grads* grads, const void* saved, void* ws,
                              void* stream);

/* ------------------------------------------------------------------------------------------------
 * SpatialOutputAdapter, split at its decoder_transformer (multimae/output_adapters.py:236-282):
 *   head: proj_context -> queries/context construction (get_queries_and_context :183-234) -> LayerNorms ->
 *         CrossAttention (no residual) -> x + mlp(out_norm(x))                                   (:258-266)
 *   tail: out_proj + un-patchify to [B, C, H, W]                                                (:274-280)
 * ---------------------------------------------------------------------------------------------- */
typedef struct mmae_decoder_index {
  int batch, dim, num_visible, num_global, num_queries, total_tokens, num_tasks, own_task;
  /* query_mode 0 (output_adapters.py:209-213): the queries are the own_task rows of the restored context, own_task in
   * [0, num_tasks).  query_mode 1 (:214-221, use_task_queries=False or a task that is not among the inputs): every query
   * is mask_token + pos (+ task_emb[own_task] when own_task >= 0; the slot may be num_tasks, an embedding that belongs to
   * no input task). */
  int query_mode;
  int tok_offset[MMAE_MAX_TASKS + 1];
  const int64_t* ids_keep;    /* [B, num_visible] */
  const int64_t* ids_restore; /* [B, total_tokens] */
} mmae_decoder_index;

typedef struct mmae_dechead_params {
  const float *proj_context_w, *proj_context_b, *mask_token, *pos;  /* pos: [num_queries, Dd] resized table rows */
  const float* task_emb[MMAE_MAX_TASKS];                           /* [Dd] per context task or NULL */
  const float *context_norm_w, *context_norm_b, *query_norm_w, *query_norm_b, *out_norm_w, *out_norm_b;
  const float *q_w, *q_b, *kv_w, *kv_b, *proj_w, *proj_b, *fc1_w, *fc1_b, *fc2_w, *fc2_b;
} mmae_dechead_params;
typedef struct mmae_dechead_grads {
  float *proj_context_w, *proj_context_b, *mask_token;
  float* task_emb[MMAE_MAX_TASKS];
  float *context_norm_w, *context_norm_b, *query_norm_w, *query_norm_b, *out_norm_w, *out_norm_b;
  float *q_w, *q_b, *kv_w, *kv_b, *proj_w, *proj_b, *fc1_w, *fc1_b, *fc2_w, *fc2_b;
} mmae_dechead_grads;

int64_t mmae_dechead_saved_bytes(const mmae_decoder_index* ix, int D_enc, int H, int hidden);
int64_t mmae_dechead_workspace_bytes(const mmae_decoder_index* ix, int D_enc, int H, int hidden);
/* enc: [B, T+G, D_enc] fp32 encoder output; x_out: [B, num_queries, Dd] fp32 */
int mmae_dechead_forward(const float* enc, int D_enc, const mmae_decoder_index* ix, int H, int hidden, float eps,
                         const mmae_dechead_params* prm, float* x_out, void* saved, void* ws, void* stream);
/* denc is ACCUMULATED (+=): the four adapters share one encoder-output gradient */
int mmae_dechead_backward(const float* enc, int D_enc, const mmae_decoder_index* ix, int H, int hidden,
                          const mmae_dechead_params* prm, const mmae_dechead_grads* grads, const float* dx_out,
                          float* denc, const void* saved, void* ws, void* stream);

/* Shared context projection.  MultiMAE.forward passes the SAME encoder output to every output adapter
 * (multimae/multimae.py:357-366) and each adapter begins with its own proj_context Linear
 * (multimae/output_adapters.py:258): n Linears [rows, D_enc] -> [rows, Dd_i] on one input.  mmae_ctxproj_forward runs them
 * as ONE GEMM with N = sum_i Dd_i (one bf16 cast of enc, ctx = enc W_cat^T + b_cat, fp32 [rows, dim_total], adapter i owns
 * the column segment starting at sum(dim[0..i))); the heads then run through mmae_dechead_forward_ctx /
 * mmae_dechead_backward_ctx, which read their segment, write their bf16 context gradient into the matching segment of one
 * [rows, dim_total] matrix and produce the proj_context BIAS gradient; mmae_ctxproj_backward finishes with one weight
 * gradient GEMM (K = rows) and one input-gradient GEMM (K = dim_total) that WRITES denc.  Parameters, their bf16 mirror and
 * their gradient slots are used in place when the n tensors lie back to back in memory, gathered otherwise.
 * For the *_ctx heads, mmae_dechead_saved_bytes / _workspace_bytes are queried with D_enc = 0. */
typedef struct mmae_ctxproj_params {
  int num;                              /* adapters sharing the projection, 1..MMAE_MAX_TASKS */
  int dim[MMAE_MAX_TASKS];              /* dim_tokens of each adapter (multiples of 8) */
  const float* weight[MMAE_MAX_TASKS];  /* proj_context.weight [dim_i, D_enc] */
  const float* bias[MMAE_MAX_TASKS];    /* proj_context.bias [dim_i] */
} mmae_ctxproj_params;
typedef struct mmae_ctxproj_grads {     /* accumulated (+=) */
  float* weight[MMAE_MAX_TASKS];
} mmae_ctxproj_grads;
int64_t mmae_ctxproj_saved_bytes(int rows, int D_enc, int dim_total);
/* enc: [rows, D_enc] fp32 -> ctx: [rows, dim_total] fp32 */
int mmae_ctxproj_forward(const float* enc, int rows, int D_enc, const mmae_ctxproj_params* prm, float* ctx, void* saved,
                         void* stream);
/* dctx_bf16: [rows, dim_total] bf16 (every segment written by its head); denc: [rows, D_enc] fp32, WRITTEN */
int mmae_ctxproj_backward(int rows, int D_enc, const mmae_ctxproj_params* prm, const mmae_ctxproj_grads* grads,
                          const void* dctx_bf16, float* denc, const void* saved, void* stream);
/* ctx: this adapter's segment of the shared projection (row stride ld_ctx floats); prm->proj_context_w / _b are unused */
int mmae_dechead_forward_ctx(const float* ctx, int64_t ld_ctx, const mmae_decoder_index* ix, int H, int hidden, float eps,
                             const mmae_dechead_params* prm, float* x_out, void* saved, void* ws, void* stream);
/* dctx_bf16: this adapter's segment of the shared gradient matrix (row stride ld_dctx bf16 elements, WRITTEN);
 * grads->proj_context_b is accumulated, grads->proj_context_w is left to mmae_ctxproj_backward */
int mmae_dechead_backward_ctx(const mmae_decoder_index* ix, int H, int hidden, const mmae_dechead_params* prm,
                              const mmae_dechead_grads* grads, const float* dx_out, void* dctx_bf16, int64_t ld_dctx,
                              const void* saved, void* ws, void* stream);

int64_t mmae_dectail_saved_bytes(int B, int nh, int nw, int Dd, int C, int P);
int64_t mmae_dectail_workspace_bytes(int B, int nh, int nw, int Dd, int C, int P);
/* x: [B, nh*nw, Dd] fp32 -> pred [B, C, nh*P, nw*P] fp32 */
int mmae_dectail_forward(const float* x, int B, int nh, int nw, int Dd, int C, int P, const float* out_w,
                         const float* out_b, float* pred, void* saved, void* ws, void* stream);
int mmae_dectail_backward(const float* dpred, int B, int nh, int nw, int Dd, int C, int P, const float* out_w,
                          float* d_out_w, float* d_out_b, float* dx, const void* saved, void* ws, void* stream);

/* ------------------------------------------------------------------------------------------------
 * fp32 tier for `fp32_output_adapters` (multimae/multimae.py:367-377: the listed output adapters run outside autocast;
 * the shipped pre-training config lists ['semseg']).  Every activation stays fp32; Linear layers run as ONE bf16 tcgen05
 * GEMM over 3-way split, K-concatenated operands (x_hi W_hi + x_hi W_lo + x_lo W_hi, fp32 accumulation: ~2^-16 relative),
 * attention and GELU as fp32 CUDA-core kernels.  Same argument meaning as the bf16-tier entry points of the same name;
 * the parameter / gradient structs are shared.  head_dim 32 only (the decoders').
 * ---------------------------------------------------------------------------------------------- */
int64_t mmae_linear_f32_workspace_bytes(int M, int N, int K);
/* y[M,N] = x[M,K] W[N,K]^T (+ bias[N]) (+ residual[M,N]); N, K multiples of 8 */
int mmae_linear_f32_forward(const float* x, const float* W, const float* bias, const float* residual, float* y, int M, int N,
                            int K, void* ws, void* stream);
/* dx[M,K] = dy W (when dx != NULL); dW[N,K] += dy^T x, db[N] += colsum(dy) (when dW / db != NULL); M, N, K multiples of 8 */
int mmae_linear_f32_backward(const float* x, const float* W, const float* dy, float* dx, float* dW, float* db, int M, int N, int K,
                             void* ws, void* stream);
/* io = gelu(z) (backward = 0) or io *= gelu'(z) (backward = 1); exact erf form, n multiple of 4 */
int mmae_gelu_f32(const float* z, float* io, int64_t n, int backward, void* stream);
int mmae_attention_f32_forward(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv, float* o,
                               int64_t ldo, float* lse, int B, int H, int Nq, int Nk, int head_dim, float scale, void* stream);
int mmae_attention_f32_backward(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                                const float* o, int64_t ldo, const float* d_o, int64_t lddo, const float* lse, float* delta_ws,
                                float* dq, int64_t lddq, float* dk, int64_t lddk, float* dv, int64_t lddv, int B, int H, int Nq,
                                int Nk, int head_dim, float scale, void* stream);
int64_t mmae_block_f32_saved_bytes(int B, int N, int D, int H, int hidden);
int64_t mmae_block_f32_workspace_bytes(int B, int N, int D, int H, int hidden);
int mmae_block_f32_forward(const float* x_in, float* x_out, int B, int N, int D, int H, int hidden, float eps,
                           const mmae_block_params* prm, void* saved, void* ws, void* stream);
int mmae_block_f32_backward(const float* x_in, const float* dx_out, float* dx_in, int B, int N, int D, int H, int hidden,
                            const mmae_block_params* prm, const mmae_block_grads* grads, const void* saved, void* ws,
                            void* stream);
int64_t mmae_dechead_f32_saved_bytes(const mmae_decoder_index* ix, int D_enc, int H, int hidden);
int64_t mmae_dechead_f32_workspace_bytes(const mmae_decoder_index* ix, int D_enc, int H, int hidden);
int mmae_dechead_f32_forward(const float* enc, int D_enc, const mmae_decoder_index* ix, int H, int hidden, float eps,
                             const mmae_dechead_params* prm, float* x_out, void* saved, void* ws, void* stream);
int mmae_dechead_f32_backward(const float* enc, int D_enc, const mmae_decoder_index* ix, int H, int hidden,
                              const mmae_dechead_params* prm, const mmae_dechead_grads* grads, const float* dx_out,
                              float* denc, const void* saved, void* ws, void* stream);
int64_t mmae_dectail_f32_workspace_bytes(int B, int nh, int nw, int Dd, int C, int P);
int mmae_dectail_f32_forward(const float* x, int B, int nh, int nw, int Dd, int C, int P, const float* out_w,
                             const float* out_b, float* pred, void* ws, void* stream);
int mmae_dectail_f32_backward(const float* x, const float* dpred, int B, int nh, int nw, int Dd, int C, int P,
                              const float* out_w, float* d_out_w, float* d_out_b, float* dx, void* ws, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Masked reconstruction losses (multimae/criterion.py:37-57, 84-114, 141-171).
 * kind: 0 = MSE, 1 = L1, 2 = cross-entropy (pred = logits [B,C,H,W], target = int64 [B,H,W]).
 * mask: int64 [B, (H/scale)*(W/scale)] (non-zero = contributes) or NULL (plain mean).  ws: 2*B floats kept until
 * backward.  loss_out / grad_out are device scalars.  No host synchronisation (mask.sum()==0 -> 0 on device).
 * ---------------------------------------------------------------------------------------------- */
int mmae_masked_loss_forward(int kind, int norm_pix, float label_smoothing, const float* pred, const void* target,
                             const int64_t* mask, int B, int C, int H, int W, int scale, float* ws, float* loss_out,
                             void* stream);
int mmae_masked_loss_backward(int kind, int norm_pix, float label_smoothing, const float* pred, const void* target,
                              const int64_t* mask, int B, int C, int H, int W, int scale, const float* ws,
                              const float* grad_out, float* dpred, void* stream);

/* bf16 weight mirror.  Registers a bf16 twin (same element count and layout) of a flat fp32 parameter buffer; pass
 * mirror_bf16 = NULL to unregister.  While registered: (1) the block / decoder / tail orchestrators take the bf16 GEMM
 * operand of any weight that lies inside `params_f32` (at an offset that is a multiple of 8 elements) from the twin
 * instead of casting it per call, (2) mmae_adamw_step on `params_f32` also writes the updated values to the twin.  The
 * caller keeps the twin in sync after any other modification of the parameters (mmae_cast_f32_to_bf16 over the buffer).
 * Replaces the per-Linear autocast weight casts of the reference (torch.cuda.amp.autocast, run_pretraining_multimae.py:452). */
int mmae_weight_mirror_register(const float* params_f32, void* mirror_bf16, int64_t n);

/* ------------------------------------------------------------------------------------------------
 * Flat-buffer gradient post-processing and AdamW (utils/native_scaler.py:34-36, 49-62; torch.optim.AdamW as built
 * by utils/optim_factory.py:155-174).  out2[0] = sum of squares AFTER unscaling, out2[1] = 1 if a non-finite
 * gradient was seen; grads *= inv_scale * post_scale (inv_scale read from inv_scale_dev when non-NULL).
 * ---------------------------------------------------------------------------------------------- */
int mmae_grad_unscale_norm(float* grads, int64_t n, const float* inv_scale_dev, float inv_scale, float post_scale,
                           float* out2, float* norm_out, void* stream);
/* dyn_lr_step_dev (optional): device float[2] = {learning rate, step count}; when non-NULL it overrides the host
 * `lr` / `step` arguments so that a CUDA-graph replay picks up the current schedule values, and the call itself advances
 * the step count by one on the device - unless found_inf_dev[0] != 0, in which case nothing is updated (GradScaler.step
 * does not call optimizer.step() for a non-finite gradient, so skipped steps do not advance the bias correction). */
int mmae_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                    float beta1, float beta2, float eps, float weight_decay, int step, const float* found_inf_dev,
                    const float* dyn_lr_step_dev, void* stream);

/* image <-> token layout helpers ('b (nh nw) (c ph pw) <-> b c (nh ph) (nw pw)') */
int mmae_unpatchify(const float* tokens, int64_t ld_tok, float* image, int B, int C, int nh, int nw, int P,
                    void* stream);
int mmae_patchify(const float* image, float* tokens, int64_t ld_tok, int B, int C, int nh, int nw, int P, void* stream);
int mmae_unpatchify_bf16(const void* tokens_bf16, int64_t ld_tok, float* image, int B, int C, int nh, int nw, int P,
                         void* stream);
int mmae_patchify_bf16(const float* image, void* tokens_bf16, int64_t ld_tok, int B, int C, int nh, int nw, int P,
                       void* stream);

/* ------------------------------------------------------------------------------------------------
 * Truncated depth standardisation — the caller-side step of train_one_epoch, run_pretraining_multimae.py:487-492
 * (torch.sort of every depth map, slice [int(0.1 n), int(0.9 n)), mean / unbiased var of the slice, standardise
 * the whole map).  depth, out: [B, n] fp32 (out may alias depth); lo / hi: the slice bounds as the reference
 * computes them on the host; eps: 1e-6 in the reference; stats (optional): [B, 2] = {mean, var} of the kept
 * values.  One launch: radix select of the two order statistics instead of a sort.
 * ---------------------------------------------------------------------------------------------- */
int mmae_standardize_depth(const float* depth, float* out, int B, int n, int lo, int hi, float eps, float* stats,
                           void* stream);
/* 1 (default): the kernel validated in round 1; 2: experimental second version (histogram copies, cluster split of
 * large maps over distributed shared memory) - also MMAE_DEPTH_STD_VARIANT=2 */
int mmae_standardize_depth_set_variant(int variant);

#ifdef __cplusplus
}
#endif
#endif /* MULTIMAE_B200_H_ */

// Flat-buffer gradient post-processing and optimizer update (HBM-bound, single pass each).
//
// mmae_grad_unscale_norm replaces GradScaler.unscale_ + get_grad_norm_ (utils/native_scaler.py:34-35, 49-62): the
// reference launches ~344 per-tensor norm kernels + a stack + a norm; the L2 norm of per-tensor L2 norms equals the
// L2 norm of the concatenation, so one pass over the flat gradient buffer suffices.
// mmae_adamw_step is torch.optim.AdamW's update (utils/optim_factory.py:155-174 builds it) over flat buffers.
#include "common.cuh"
#include "../../include/multimae_b200.h"
#include "internal.h"

namespace mmae {
void count_launch();
namespace {

// out[0] += sum(g^2) after unscaling, out[1] = 1.0 if any non-finite value was seen (left untouched otherwise)
__global__ void __launch_bounds__(256) unscale_norm_kernel(float* __restrict__ g, int64_t n, const float* __restrict__ inv_scale_ptr,
                                                           float inv_scale_imm, float post_scale, float* __restrict__ out) {
  pdl_prologue();
  __shared__ float red[8];
  const float inv = (inv_scale_ptr ? inv_scale_ptr[0] : inv_scale_imm) * post_scale;
  float acc = 0.f;
  bool bad = false;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x * 4;
  constexpr int U = 4;   // independent 16-byte loads in flight per thread
  for (int64_t i0 = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i0 < n; i0 += stride * U) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i + 4 <= n) v[u] = *reinterpret_cast<const float4*>(g + i);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i + 4 <= n) {
        float4 w = v[u];
        w.x *= inv; w.y *= inv; w.z *= inv; w.w *= inv;
        if (inv != 1.0f) *reinterpret_cast<float4*>(g + i) = w;
        acc += w.x * w.x + w.y * w.y + w.z * w.z + w.w * w.w;
        bad |= !(isfinite(w.x) && isfinite(w.y) && isfinite(w.z) && isfinite(w.w));
      } else if (i < n) {
        for (int64_t j = i; j < n; ++j) {
          const float w = g[j] * inv;
          if (inv != 1.0f) g[j] = w;
          acc += w * w;
          bad |= !isfinite(w);
        }
      }
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float t = red[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffu, t, o);
    if (threadIdx.x == 0) atomicAdd(out, t);
  }
  if (__syncthreads_or(bad ? 1 : 0) && threadIdx.x == 0) out[1] = 1.0f;
}

__global__ void sqrt_kernel(const float* __restrict__ in, float* __restrict__ norm_out) {
  pdl_prologue(); norm_out[0] = sqrtf(in[0]); }

// device-side step counter: dyn[1] advances by one unless the step is skipped (non-finite gradients), like GradScaler.step,
// which does not call optimizer.step() at all in that case - the bias correction must not see skipped steps
__global__ void adamw_advance_kernel(float* __restrict__ dyn, const float* __restrict__ found_inf) {
  pdl_prologue();
  if (found_inf == nullptr || found_inf[0] == 0.f) dyn[1] += 1.0f;
}

// decoupled weight decay AdamW, skipping the whole update when found_inf[0] != 0 (GradScaler.step semantics)
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, int64_t n, float lr, float beta1, float beta2,
                                                    float eps, float wd, float bc1, float bc2_sqrt,
                                                    const float* __restrict__ found_inf, const float* __restrict__ dyn,
                                                    bf16* __restrict__ mirror) {
  pdl_prologue();
  if (found_inf != nullptr && found_inf[0] != 0.f) return;
  if (dyn != nullptr) {   // learning rate and step count live on the device (CUDA-graph replay): dyn = {lr, step}
    lr = dyn[0];
    const float step = dyn[1];
    bc1 = 1.0f - powf(beta1, step);
    bc2_sqrt = sqrtf(1.0f - powf(beta2, step));
  }
  const int64_t stride = int64_t(gridDim.x) * blockDim.x * 4;
  for (int64_t i = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 4 <= n) {
      float4 pp = *reinterpret_cast<float4*>(p + i);
      const float4 gg = *reinterpret_cast<const float4*>(g + i);
      float4 mm = *reinterpret_cast<float4*>(m + i);
      float4 vv = *reinterpret_cast<float4*>(v + i);
#define MMAE_ADAM1(P, G, M, V)                                     \
  P *= (1.f - lr * wd);                                            \
  M = beta1 * M + (1.f - beta1) * G;                               \
  V = beta2 * V + (1.f - beta2) * G * G;                           \
  P -= (lr / bc1) * M / (sqrtf(V) / bc2_sqrt + eps);
      MMAE_ADAM1(pp.x, gg.x, mm.x, vv.x)
      MMAE_ADAM1(pp.y, gg.y, mm.y, vv.y)
      MMAE_ADAM1(pp.z, gg.z, mm.z, vv.z)
      MMAE_ADAM1(pp.w, gg.w, mm.w, vv.w)
      *reinterpret_cast<float4*>(p + i) = pp;
      *reinterpret_cast<float4*>(m + i) = mm;
      *reinterpret_cast<float4*>(v + i) = vv;
      if (mirror != nullptr) {   // bf16 twin of the parameters: the GEMM weight operands of the next step
        uint2 pk;
        pk.x = pack_bf16x2(pp.x, pp.y);
        pk.y = pack_bf16x2(pp.z, pp.w);
        *reinterpret_cast<uint2*>(mirror + i) = pk;
      }
    } else {
      for (int64_t j = i; j < n; ++j) {
        float P = p[j], M = m[j], V = v[j];
        const float G = g[j];
        MMAE_ADAM1(P, G, M, V)
        p[j] = P; m[j] = M; v[j] = V;
        if (mirror != nullptr) mirror[j] = __float2bfloat16_rn(P);
      }
    }
  }
#undef MMAE_ADAM1
}

}  // namespace
}  // namespace mmae

using namespace mmae;

extern "C" int mmae_grad_unscale_norm(float* grads, int64_t n, const float* inv_scale_dev, float inv_scale,
                                      float post_scale, float* out2, float* norm_out, void* stream) {
  MMAE_CHECK(grads && out2 && n >= 0 && (reinterpret_cast<uintptr_t>(grads) & 15) == 0, MMAE_ERR_ARG,
             "mmae_grad_unscale_norm: bad args");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  MMAE_CUDA_OK(cudaMemsetAsync(out2, 0, 2 * sizeof(float), st));
  if (n > 0) {
    int64_t blocks = (n / 4 + 255) / 256;
    const int64_t cap = int64_t(sm_count()) * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    launch_k(unscale_norm_kernel, (unsigned)blocks, 256, 0, st, grads, n, inv_scale_dev, inv_scale, post_scale, out2);
    count_launch();
    MMAE_LAUNCH_OK();
  }
  if (norm_out) {
    launch_k(sqrt_kernel, 1, 1, 0, st, out2, norm_out);
    count_launch();
    MMAE_LAUNCH_OK();
  }
  return MMAE_OK;
}

extern "C" int mmae_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                               float beta1, float beta2, float eps, float weight_decay, int step,
                               const float* found_inf_dev, const float* dyn_lr_step_dev, void* stream) {
  MMAE_CHECK(params && grads && exp_avg && exp_avg_sq && n >= 0 && step >= 1, MMAE_ERR_ARG, "mmae_adamw_step: bad args");
  if (n == 0) return MMAE_OK;
  const float bc1 = 1.0f - powf(beta1, (float)step);
  const float bc2_sqrt = sqrtf(1.0f - powf(beta2, (float)step));
  int64_t blocks = (n / 4 + 255) / 256;
  const int64_t cap = int64_t(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (dyn_lr_step_dev != nullptr) {
    launch_k(adamw_advance_kernel, 1, 1, 0, reinterpret_cast<cudaStream_t>(stream), const_cast<float*>(dyn_lr_step_dev), found_inf_dev);
    count_launch();
    MMAE_LAUNCH_OK();
  }
  launch_k(adamw_kernel, (unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream), params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2_sqrt, found_inf_dev, dyn_lr_step_dev,
      const_cast<bf16*>(mirror_lookup(params)));
  count_launch();
  MMAE_LAUNCH_OK();
  return MMAE_OK;
}

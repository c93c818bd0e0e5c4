// Dirichlet token-mask sampler as ONE kernel: one CTA per sample, bitonic sorts in shared memory.
//
// Pure function of the random draws (task shares + uniform noise), so a CPU oracle fed the same draws agrees
// bit for bit.  Replaces the ~25 launches (5 argsorts, gathers, where, cat, split) of
// MultiMAE.generate_random_masks, multimae/multimae.py:189-216:
//   k_t        = round_half_even(share_t * num_encoded)                                    (:189)
//   order_t    = argsort(noise_t)              ; mask_t[j] = order_t[j] < k_t ? 0 : 1      (:195-200)
//   ids_shuffle= argsort(float(mask_all) + noise_all) ; ids_restore = argsort(ids_shuffle) (:203-205)
//   ids_keep   = ids_shuffle[:, :num_encoded]                                              (:206)
//   task_masks = (ids_restore >= num_encoded)                                              (:209-214)
// Ties are broken by lower index first (the reference's argsort is unstable).
#include "common.cuh"
#include "../../include/multimae_b200.h"

namespace mmae {
void count_launch();
namespace {

constexpr int MS_THREADS = 256;

// sort `n_pad` (power of two) 64-bit keys ascending
__device__ void bitonic_sort(unsigned long long* keys, int n_pad) {
  for (int k = 2; k <= n_pad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < n_pad; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = keys[i], b = keys[ixj];
          const bool up = (i & k) == 0;
          if ((a > b) == up) {
            keys[i] = b;
            keys[ixj] = a;
          }
        }
      }
    }
  }
  __syncthreads();
}

__device__ __forceinline__ int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

struct SamplerTasks {
  int num_tasks;
  int offset[9];  // offset[t] .. offset[t+1]: token range of task t in the concatenated sequence
};

__global__ void __launch_bounds__(MS_THREADS) mask_sampler_kernel(const float* __restrict__ shares,
                                                                  const float* __restrict__ noise_task,
                                                                  const float* __restrict__ noise_all,
                                                                  SamplerTasks tasks, int num_encoded,
                                                                  int64_t* __restrict__ task_masks,
                                                                  int64_t* __restrict__ ids_keep,
                                                                  int64_t* __restrict__ ids_restore) {
  pdl_prologue();
  extern __shared__ unsigned long long keys[];  // n_pad entries
  const int b = blockIdx.x;
  const int total = tasks.offset[tasks.num_tasks];
  unsigned char* vis = reinterpret_cast<unsigned char*>(keys + next_pow2(total));  // per-token first-stage mask

  // ---- stage 1: per-task keep sets
  for (int t = 0; t < tasks.num_tasks; ++t) {
    const int off = tasks.offset[t], n = tasks.offset[t + 1] - off;
    const int n_pad = next_pow2(n);
    const long long k_t = (long long)rintf(shares[b * tasks.num_tasks + t] * (float)num_encoded);
    __syncthreads();
    for (int i = threadIdx.x; i < n_pad; i += blockDim.x) {
      unsigned long long key = ~0ull;
      if (i < n) key = ((unsigned long long)__float_as_uint(noise_task[int64_t(b) * total + off + i]) << 32) | (unsigned)i;
      keys[i] = key;
    }
    bitonic_sort(keys, n_pad);
    // position j is visible iff the index of the j-th smallest noise value is < k_t
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
      const long long idx = (long long)(keys[j] & 0xffffffffull);
      vis[off + j] = idx < k_t ? 0 : 1;
    }
  }
  __syncthreads();

  // ---- stage 2: global shuffle, visible tokens first
  const int n_pad = next_pow2(total);
  for (int i = threadIdx.x; i < n_pad; i += blockDim.x) {
    unsigned long long key = ~0ull;
    if (i < total) {
      const float k2 = __fadd_rn((float)vis[i], noise_all[int64_t(b) * total + i]);
      key = ((unsigned long long)__float_as_uint(k2) << 32) | (unsigned)i;
    }
    keys[i] = key;
  }
  bitonic_sort(keys, n_pad);
  for (int r = threadIdx.x; r < total; r += blockDim.x) {
    const int idx = (int)(keys[r] & 0xffffffffull);
    ids_restore[int64_t(b) * total + idx] = r;
    task_masks[int64_t(b) * total + idx] = r < num_encoded ? 0 : 1;
    if (r < num_encoded) ids_keep[int64_t(b) * num_encoded + r] = idx;
  }
}

}  // namespace
}  // namespace mmae

using namespace mmae;

extern "C" int mmae_sample_masks(const float* shares, const float* noise_task, const float* noise_all, int B,
                                 int num_tasks, const int* tokens_per_task_host, int num_encoded,
                                 int64_t* task_masks, int64_t* ids_keep, int64_t* ids_restore, void* stream) {
  MMAE_CHECK(shares && noise_task && noise_all && tokens_per_task_host && task_masks && ids_keep && ids_restore,
             MMAE_ERR_ARG, "mmae_sample_masks: null argument");
  MMAE_CHECK(B > 0 && num_tasks > 0 && num_tasks <= 8, MMAE_ERR_ARG, "mmae_sample_masks: 1..8 tasks supported");
  SamplerTasks tasks;
  tasks.num_tasks = num_tasks;
  tasks.offset[0] = 0;
  for (int t = 0; t < num_tasks; ++t) {
    MMAE_CHECK(tokens_per_task_host[t] > 0, MMAE_ERR_ARG, "mmae_sample_masks: empty task");
    tasks.offset[t + 1] = tasks.offset[t] + tokens_per_task_host[t];
  }
  const int total = tasks.offset[num_tasks];
  MMAE_CHECK(num_encoded > 0 && num_encoded <= total, MMAE_ERR_ARG, "mmae_sample_masks: num_encoded out of range");
  int n_pad = 1;
  while (n_pad < total) n_pad <<= 1;
  const size_t smem = size_t(n_pad) * 8 + align_up(total, 16);
  MMAE_CHECK(smem <= 200 * 1024, MMAE_ERR_UNSUPPORTED, "mmae_sample_masks: %d tokens exceed the shared-memory sort", total);
  static bool configured = false;
  if (!configured) {
    MMAE_CUDA_OK(cudaFuncSetAttribute(mask_sampler_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  launch_k(mask_sampler_kernel, B, MS_THREADS, smem, reinterpret_cast<cudaStream_t>(stream), shares, noise_task, noise_all, tasks, num_encoded, task_masks, ids_keep, ids_restore);
  count_launch();
  MMAE_LAUNCH_OK();
  return MMAE_OK;
}

// tcgen05 / TMA GEMM for sm_100a:  C[M,N] = epilogue(alpha * A * B^T), bf16 operands, fp32 accumulate in TMEM.
//
// One CTA computes one 128 x BN output tile over a K range (split-K along gridDim.z).
//   warp 0     : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier complete_tx)
//   warp 1     : TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BN x 16), tcgen05.commit
//   warps 2..5 : epilogue       (tcgen05.ld TMEM -> registers -> fused bias/GELU/dGELU/residual -> global)
// Two CTAs are resident per SM (3-stage ring, 96 KB each), so one CTA's epilogue overlaps the other's mainloop.
//
// Replaces the cuBLAS/cuDNN calls behind nn.Linear / nn.Conv2d(k=s=P) on the reference path
// (multimae/multimae_utils.py:149-153,172,180,203-212; multimae/input_adapters.py:110,232;
//  multimae/output_adapters.py:258,274) and their autograd dgrad/wgrad.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "../../include/multimae_b200.h"

namespace mmae {

void count_launch();
bool gemm_profile_begin(cudaStream_t st, double flops, int M, int N, int K, int flags);
void gemm_profile_end(cudaStream_t st);

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 192;

struct GemmParams {
  int M, N, K;
  int num_kb;        // total k-blocks (ceil(K / BK))
  int kb_per_split;  // k-blocks handled by one split
  int tma_store;     // persistent kernel, output through shared memory + TMA: 1 = bf16 store, 2 = fp32 store, 3 = fp32 add
  mmae_gemm_epilogue ep;
};

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int BIAS_OFFSET = BAR_OFFSET + 256;
  static constexpr int TOTAL = BIAS_OFFSET + BN * 4 + 1024;  // + barriers + bias tile + alignment slack
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}


// Fused epilogue for 32 consecutive accumulator columns of one output row (held by one thread).
// `sbias`: the 32 bias values of this chunk staged in shared memory (or nullptr).  Reading the bias with per-group
// global loads inside this loop cost 30-40 % of the kernel (measured: 962 -> 673 TF/s at 12672x3072x768).
__device__ __forceinline__ void epilogue_chunk32(const uint32_t (&r)[32], int row, bool row_ok, int nb, const GemmParams& p,
                                                 bool first_split, bool atomic_out, const float* sbias) {
  const mmae_gemm_epilogue& ep = p.ep;
  const float* bias = sbias;
  const float* resid = first_split ? ep.residual : nullptr;
  const bf16* zptr = reinterpret_cast<const bf16*>(ep.dgelu_z);
  bf16* preact = reinterpret_cast<bf16*>(ep.preact_bf16);
  bf16* out_b = reinterpret_cast<bf16*>(ep.out_bf16);
  float* out_f = ep.out_f32;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int n = nb + g * 8;
    if (!row_ok || n >= p.N) continue;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[g * 8 + i]) * ep.alpha;
    if (bias) {
      const float4 b0 = *reinterpret_cast<const float4*>(bias + g * 8);
      const float4 b1 = *reinterpret_cast<const float4*>(bias + g * 8 + 4);
      v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
      v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
    }
    if (preact) {
      uint4 o;
      o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
      o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
      *reinterpret_cast<uint4*>(preact + int64_t(row) * ep.ld_preact + n) = o;
    }
    if (ep.act == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = gelu_erf(v[i]);
    }
    if (zptr) {
      const uint4 z = __ldg(reinterpret_cast<const uint4*>(zptr + int64_t(row) * ep.ld_dgelu_z + n));
      const float2 z0 = unpack_bf16x2(z.x), z1 = unpack_bf16x2(z.y), z2 = unpack_bf16x2(z.z), z3 = unpack_bf16x2(z.w);
      v[0] *= dgelu_erf(z0.x); v[1] *= dgelu_erf(z0.y); v[2] *= dgelu_erf(z1.x); v[3] *= dgelu_erf(z1.y);
      v[4] *= dgelu_erf(z2.x); v[5] *= dgelu_erf(z2.y); v[6] *= dgelu_erf(z3.x); v[7] *= dgelu_erf(z3.y);
    }
    if (resid) {
      const float4 r0 = __ldg(reinterpret_cast<const float4*>(resid + int64_t(row) * ep.ld_residual + n));
      const float4 r1 = __ldg(reinterpret_cast<const float4*>(resid + int64_t(row) * ep.ld_residual + n + 4));
      v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
      v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
    }
    if (out_f) {
      float* dst = out_f + int64_t(row) * ep.ld_out_f32 + n;
      if (atomic_out) {
        red_add_v4(dst, v[0], v[1], v[2], v[3]);
        red_add_v4(dst + 4, v[4], v[5], v[6], v[7]);
      } else {
        *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
      }
    }
    if (out_b) {
      uint4 o;
      o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
      o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
      *reinterpret_cast<uint4*>(out_b + int64_t(row) * ep.ld_out_bf16 + n) = o;
    }
  }
}


template <int BN, int STAGES, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, (BN <= 128 ? 2 : 1))
    gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const GemmParams p) {
  pdl_launch_dependents();   // the wait follows the barrier / TMEM setup below
  using L = GemmSmem<BN, STAGES>;
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;  // power of two >= 32 (BN in {64,128,256})

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * BM;
  const int kb_begin = blockIdx.z * p.kb_per_split;
  int nkb = p.num_kb - kb_begin;
  if (nkb > p.kb_per_split) nkb = p.kb_per_split;
  if (nkb < 0) nkb = 0;

  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
    }
  } else if (warp == 1) {
    if (elect_one()) {
#pragma unroll
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      mbar_init(tmem_full_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr_smem, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();   // the previous kernel's outputs (our operands) are complete and visible from here on

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        mbar_expect_tx(&full_bar[s], L::STAGE_BYTES);
        uint8_t* sA = smem + s * L::STAGE_BYTES;
        uint8_t* sB = sA + L::A_BYTES;
        const int k0 = (kb_begin + kb) * BK;
        if constexpr (!A_MN) {
          tma_load_2d(sA, &tmA, &full_bar[s], k0, m0);
        } else {
#pragma unroll
          for (int c = 0; c < BM / 64; ++c) tma_load_2d(sA + c * (64 * BK * 2), &tmA, &full_bar[s], m0 + c * 64, k0);
        }
        if constexpr (!B_MN) {
          tma_load_2d(sB, &tmB, &full_bar[s], k0, n0);
        } else {
#pragma unroll
          for (int c = 0; c < BN / 64; ++c) tma_load_2d(sB + c * (64 * BK * 2), &tmB, &full_bar[s], n0 + c * 64, k0);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * L::STAGE_BYTES);
        const uint32_t b_addr = a_addr + L::A_BYTES;
#pragma unroll
        for (int j = 0; j < BK / UMMA_K; ++j) {
          // K-major : advance 16 elements (32 B) inside the 128 B swizzle row
          // MN-major: advance 16 k-rows of 128 B; 64-element M/N chunks are (64*BK*2) bytes apart
          const uint64_t da = A_MN ? umma_smem_desc_sw128(a_addr + j * (UMMA_K * 128), 64 * BK * 2, 1024)
                                   : umma_smem_desc_sw128(a_addr + j * (UMMA_K * 2), 16, 1024);
          const uint64_t db = B_MN ? umma_smem_desc_sw128(b_addr + j * (UMMA_K * 128), 64 * BK * 2, 1024)
                                   : umma_smem_desc_sw128(b_addr + j * (UMMA_K * 2), 16, 1024);
          tc_mma_f16_ss(tmem_base, da, db, idesc, (kb | j) != 0 ? 1u : 0u);
        }
        tc_commit(&empty_bar[s]);  // frees this smem stage once the MMAs above retire
      }
      tc_commit(tmem_full_bar);  // accumulator complete
    }
  } else if (nkb > 0) {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    float* sbias = reinterpret_cast<float*>(smem + L::BIAS_OFFSET);
    const bool use_bias = p.ep.bias != nullptr && blockIdx.z == 0;
    if (use_bias) {
      for (int c = (warp - 2) * 32 + lane; c < BN; c += 128) sbias[c] = n0 + c < p.N ? __ldg(p.ep.bias + n0 + c) : 0.f;
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = m0 + q * 32 + lane;
    const bool row_ok = row < p.M;
    const bool first_split = blockIdx.z == 0;
    const bool atomic_out = p.ep.accumulate != 0 || gridDim.z > 1;

#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      const int nb = n0 + c * 32;
      if (nb >= p.N) break;  // warp-uniform
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(c * 32), r);
      tc_wait_ld();
      epilogue_chunk32(r, row, row_ok, nb, p, first_split, atomic_out, use_bias ? sbias + c * 32 : nullptr);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int BN, bool A_MN, bool B_MN>
int launch_gemm1(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap&, const GemmParams& p, int split_k,
                 cudaStream_t stream) {
  constexpr int STAGES = 3;
  using L = GemmSmem<BN, STAGES>;
  auto kern = gemm_bf16_kernel<BN, STAGES, A_MN, B_MN>;
  static bool configured = false;
  if (!configured) {
    MMAE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  dim3 grid(ceil_div(p.N, BN), ceil_div(p.M, BM), split_k);
  const bool prof = gemm_profile_begin(stream, 2.0 * p.M * p.N * p.K, p.M, p.N, p.K, (A_MN ? 1 : 0) | (B_MN ? 2 : 0) | (split_k << 8));
  launch_k(kern, grid, GEMM_THREADS, L::TOTAL, stream, tmA, tmB, p);
  if (prof) gemm_profile_end(stream);
  count_launch();
  MMAE_LAUNCH_OK();
  return MMAE_OK;
}


// TMA-store epilogue for 32 accumulator columns: alpha, bias, optional GELU, bf16 pack, then one 64-byte row per lane
// into this warp's 32x32 staging tile (SWIZZLE_64B: 16-byte chunk j of row r lives at chunk j ^ ((r >> 1) & 3), which
// makes the warp's st.shared.v4 conflict-free) and one bulk tensor store per warp.  Per-lane global stores touch 32
// different 128-byte lines per instruction, i.e. ~4096 L1 tag cycles per 128x256 tile - more than the MMA time of a
// K <= 512 tile; the TMA unit writes full rows instead and clips the M / N edges itself.
__device__ __forceinline__ void epilogue_chunk32_tma(const uint32_t (&r)[32], const GemmParams& p, const float* sbias,
                                                     uint8_t* stage, int lane, const CUtensorMap* tmC, int col, int row0) {
  const mmae_gemm_epilogue& ep = p.ep;
  uint32_t packed[16];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[g * 8 + i]) * ep.alpha;
    if (sbias) {
      const float4 b0 = *reinterpret_cast<const float4*>(sbias + g * 8);
      const float4 b1 = *reinterpret_cast<const float4*>(sbias + g * 8 + 4);
      v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
      v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
    }
    if (ep.act == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = gelu_erf(v[i]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) packed[g * 4 + i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
  }
  // the previous store of this warp must have finished reading the staging tile
  if (lane == 0) bulk_wait_read_all();
  __syncwarp();
  const int sw = (lane >> 1) & 3;
#pragma unroll
  for (int g = 0; g < 4; ++g)
    *reinterpret_cast<uint4*>(stage + lane * 64 + ((g ^ sw) << 4)) =
        make_uint4(packed[g * 4], packed[g * 4 + 1], packed[g * 4 + 2], packed[g * 4 + 3]);
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    tma_store_2d(tmC, stage, col, row0);
    bulk_commit_group();
  }
}

// fp32 flavour: two 32x16 boxes (64-byte rows) per 32-column chunk, stored or ADDED (cp.reduce ... .add: split-K partial
// sums and gradient accumulation without per-lane red.global instructions).
__device__ __forceinline__ void epilogue_chunk32_tma_f32(const uint32_t (&r)[32], const GemmParams& p, const float* sbias,
                                                         uint8_t* stage, int lane, const CUtensorMap* tmC, int col, int row0,
                                                         bool add) {
  const mmae_gemm_epilogue& ep = p.ep;
  const int sw = (lane >> 1) & 3;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    if (col + h * 16 >= p.N) break;   // warp-uniform
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[h * 16 + i]) * ep.alpha;
    if (sbias) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float4 b = *reinterpret_cast<const float4*>(sbias + h * 16 + g * 4);
        v[g * 4] += b.x; v[g * 4 + 1] += b.y; v[g * 4 + 2] += b.z; v[g * 4 + 3] += b.w;
      }
    }
    if (lane == 0) bulk_wait_read_all();
    __syncwarp();
#pragma unroll
    for (int g = 0; g < 4; ++g)
      *reinterpret_cast<float4*>(stage + lane * 64 + ((g ^ sw) << 4)) = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      if (add) tma_reduce_add_2d(tmC, stage, col + h * 16, row0);
      else tma_store_2d(tmC, stage, col + h * 16, row0);
      bulk_commit_group();
    }
  }
}

// =====================================================================================================================
// v2: persistent CTAs (one per SM), double-buffered TMEM accumulators: the epilogue of work item i overlaps the
// TMA/MMA mainloop of item i+1.  BN = 256 halves the shared-memory operand bandwidth per MMA relative to BN = 128
// (a 128x128x16 UMMA consumes 128 B/clk of smem reads, the SM's limit; 128x256x16 needs 96 B/clk).
//   warp 0 : TMA producer      warp 1 : TMEM alloc + MMA issuer      warps 2..2+EPI-1 : epilogue
// Work item = (m tile, n tile, k split); items are dealt round-robin: item = blockIdx.x + i * gridDim.x, n fastest.
// =====================================================================================================================
template <int BN>
struct Gemm2Cfg {
  static constexpr int STAGES = BN == 256 ? 4 : (BN == 192 ? 5 : 6);
  static constexpr int EPI_WARPS = BN >= 192 ? 8 : 4;
  static constexpr int THREADS = 64 + EPI_WARPS * 32;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STORE_OFFSET = STAGES * STAGE_BYTES;       // one 32x32 bf16 staging tile (2 KB) per epilogue warp
  static constexpr int BAR_OFFSET = STORE_OFFSET + EPI_WARPS * 2048;
  static constexpr int BIAS_OFFSET = BAR_OFFSET + 512;             // 2 x BN floats (double-buffered per work item)
  static constexpr int TOTAL = BIAS_OFFSET + 2 * BN * 4 + 1024;
  static_assert(TOTAL <= 232448, "persistent GEMM exceeds the 227 KB shared-memory limit");
  static constexpr uint32_t TMEM_COLS = BN > 128 ? 512 : 256;  // two accumulators, power-of-two allocation
};

struct Gemm2Sched {
  int tiles_m, tiles_n, splits, total;
};

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(Gemm2Cfg<BN>::THREADS, 1)
    gemm_bf16_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                const __grid_constant__ CUtensorMap tmC, const GemmParams p, const Gemm2Sched sc) {
  pdl_launch_dependents();   // the wait follows the barrier / TMEM setup below
  using C = Gemm2Cfg<BN>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      if (p.tma_store) tma_prefetch_desc(&tmC);
    }
  } else if (warp == 1) {
    if (elect_one()) {
#pragma unroll
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        mbar_init(&tmem_full_bar[b], 1);
        mbar_init(&tmem_empty_bar[b], C::EPI_WARPS);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr_smem, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();   // the previous kernel's outputs (our operands) are complete and visible from here on

  auto decode = [&](int item, int& m0, int& n0, int& kb_begin, int& nkb, int& z) {
    z = item % sc.splits;
    const int tile = item / sc.splits;
    n0 = (tile % sc.tiles_n) * BN;
    m0 = (tile / sc.tiles_n) * BM;
    kb_begin = z * p.kb_per_split;
    nkb = min(p.kb_per_split, p.num_kb - kb_begin);
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < sc.total; item += gridDim.x) {
        int m0, n0, kb_begin, nkb, z;
        decode(item, m0, n0, kb_begin, nkb, z);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          mbar_expect_tx(&full_bar[stage], C::STAGE_BYTES);
          uint8_t* sA = smem + stage * C::STAGE_BYTES;
          uint8_t* sB = sA + C::A_BYTES;
          const int k0 = (kb_begin + kb) * BK;
          if constexpr (!A_MN) {
            tma_load_2d(sA, &tmA, &full_bar[stage], k0, m0);
          } else {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) tma_load_2d(sA + c * (64 * BK * 2), &tmA, &full_bar[stage], m0 + c * 64, k0);
          }
          if constexpr (!B_MN) {
            tma_load_2d(sB, &tmB, &full_bar[stage], k0, n0);
          } else {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c) tma_load_2d(sB + c * (64 * BK * 2), &tmB, &full_bar[stage], n0 + c * 64, k0);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int item = blockIdx.x; item < sc.total; item += gridDim.x, ++it) {
        int m0, n0, kb_begin, nkb, z;
        decode(item, m0, n0, kb_begin, nkb, z);
        const int buf = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty_bar[buf], acc_phase ^ 1u);   // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + uint32_t(buf * BN);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t b_addr = a_addr + C::A_BYTES;
#pragma unroll
          for (int j = 0; j < BK / UMMA_K; ++j) {
            const uint64_t da = A_MN ? umma_smem_desc_sw128(a_addr + j * (UMMA_K * 128), 64 * BK * 2, 1024)
                                     : umma_smem_desc_sw128(a_addr + j * (UMMA_K * 2), 16, 1024);
            const uint64_t db = B_MN ? umma_smem_desc_sw128(b_addr + j * (UMMA_K * 128), 64 * BK * 2, 1024)
                                     : umma_smem_desc_sw128(b_addr + j * (UMMA_K * 2), 16, 1024);
            tc_mma_f16_ss(tmem_d, da, db, idesc, (kb | j) != 0 ? 1u : 0u);
          }
          tc_commit(&empty_bar[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        tc_commit(&tmem_full_bar[buf]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int e = warp - 2;                       // 0 .. EPI_WARPS-1
    const int q = warp & 3;                       // TMEM lane quarter accessible to this warp
    constexpr int COLS_PER_WARP = BN / (C::EPI_WARPS / 4);   // 128 (BN 128 / 256) or 96 (BN 192)
    const int col0 = (e >> 2) * COLS_PER_WARP;
    int it = 0;
    for (int item = blockIdx.x; item < sc.total; item += gridDim.x, ++it) {
      int m0, n0, kb_begin, nkb, z;
      decode(item, m0, n0, kb_begin, nkb, z);
      const int buf = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      // stage this tile's bias in shared memory while the MMAs run; buffer (it & 1) was last read two items ago and
      // every epilogue warp has passed the previous item's barrier since
      float* sbias = reinterpret_cast<float*>(smem + C::BIAS_OFFSET) + buf * BN;
      const bool use_bias = p.ep.bias != nullptr && z == 0;
      if (p.ep.bias != nullptr) {
        if (use_bias)
          for (int c = e * 32 + lane; c < BN; c += C::EPI_WARPS * 32) sbias[c] = n0 + c < p.N ? __ldg(p.ep.bias + n0 + c) : 0.f;
        asm volatile("bar.sync 1, %0;" ::"n"(C::EPI_WARPS * 32) : "memory");
      }
      mbar_wait(&tmem_full_bar[buf], acc_phase);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.M;
      const bool first_split = z == 0;
      const bool atomic_out = p.ep.accumulate != 0 || sc.splits > 1;
      // NOTE (measured, round 1): transposing the chunk through shared memory for row-contiguous global accesses made
      // this epilogue SLOWER (it is issue/latency bound, not sector bound: QKV 900 -> 697 TF/s); one row per lane stays.
#pragma unroll 1
      for (int c = 0; c < COLS_PER_WARP / 32; ++c) {
        const int nb = n0 + col0 + c * 32;
        if (nb >= p.N) break;
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(buf * BN + col0 + c * 32), r);
        tc_wait_ld();
        if (p.tma_store == 1)
          epilogue_chunk32_tma(r, p, use_bias ? sbias + col0 + c * 32 : nullptr, smem + C::STORE_OFFSET + e * 2048, lane,
                               &tmC, nb, m0 + q * 32);
        else if (p.tma_store != 0)
          epilogue_chunk32_tma_f32(r, p, use_bias ? sbias + col0 + c * 32 : nullptr, smem + C::STORE_OFFSET + e * 2048, lane,
                                   &tmC, nb, m0 + q * 32, p.tma_store == 3);
        else
          epilogue_chunk32(r, row, row_ok, nb, p, first_split, atomic_out, use_bias ? sbias + col0 + c * 32 : nullptr);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
    }
    if (p.tma_store && lane == 0) bulk_wait_all();   // stores must have drained before the CTA's shared memory goes away
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// Persistent grid: min(items, slots) CTAs.  (Launching ceil(items / rounds) CTAs instead - 98 for a 196-tile decoder GEMM -
// to leave SMs to the other task decoders' streams was measured on B200 and removed: 16.69 vs 16.67-16.77 ms/step.)
static int persistent_grid(int items, int slots) { return std::min(items, slots); }

template <int BN, bool A_MN, bool B_MN>
int launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const GemmParams& p, int split_k,
                 cudaStream_t stream) {
  using C = Gemm2Cfg<BN>;
  auto kern = gemm_bf16_persistent_kernel<BN, A_MN, B_MN>;
  static bool configured = false;
  if (!configured) {
    MMAE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::TOTAL));
    configured = true;
  }
  Gemm2Sched sc;
  sc.tiles_m = ceil_div(p.M, BM);
  sc.tiles_n = ceil_div(p.N, BN);
  sc.splits = split_k;
  sc.total = sc.tiles_m * sc.tiles_n * split_k;
  const int grid = persistent_grid(sc.total, sm_count());
  const bool prof = gemm_profile_begin(stream, 2.0 * p.M * p.N * p.K, p.M, p.N, p.K, (A_MN ? 1 : 0) | (B_MN ? 2 : 0) | (split_k << 8));
  launch_k(kern, grid, C::THREADS, C::TOTAL, stream, tmA, tmB, tmC, p, sc);
  if (prof) gemm_profile_end(stream);
  count_launch();
  MMAE_LAUNCH_OK();
  return MMAE_OK;
}

// =====================================================================================================================
// v3: CTA pairs.  A cluster of two CTAs (the two SMs of a TPC) owns a 256 x BN output tile: each CTA stages its own 128
// rows of A and HALF of the B tile; the leader CTA (cluster rank 0) issues tcgen05.mma.cta_group::2 (M = 256), which reads
// both halves and accumulates each CTA's 128 rows in that CTA's TMEM.  Per SM and 16-deep K step the shared-memory fill +
// read traffic drops from 2 x 12 KB to 2 x 8 KB at BN = 256: the single-CTA kernels sit at ~64 % tensor-pipe activity
// because TMA fills and UMMA reads share the 128 B/clk shared-memory port (ncu: profiles/r01_ncu_full_hot_kernels.txt).
//   both CTAs : warp 0 TMA producer (own A rows, own B half; bytes counted on the LEADER's "stage full" barrier)
//               warps 2.. epilogue of the CTA's own 128 rows (same code as v2)
//   leader    : warp 1 issues the pair MMAs; its commits are multicast to the "stage empty" / "accumulator full"
//               barriers of both CTAs; the peers' epilogue warps arrive remotely on the leader's "accumulator empty".
// =====================================================================================================================
template <int BN>
struct Gemm3Cfg {
  static constexpr int BNH = BN / 2;                                // B rows (output columns) staged by one CTA
  static constexpr int STAGES = 6;
  static constexpr int EPI_WARPS = BN >= 192 ? 8 : 4;
  static constexpr int THREADS = 64 + EPI_WARPS * 32;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BNH * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STORE_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int BAR_OFFSET = STORE_OFFSET + EPI_WARPS * 2048;
  static constexpr int BIAS_OFFSET = BAR_OFFSET + 512;
  static constexpr int TOTAL = BIAS_OFFSET + 2 * BN * 4 + 1024;
  static constexpr uint32_t TMEM_COLS = BN > 128 ? 512 : 256;
  static_assert(STAGE_BYTES % 1024 == 0, "stages must keep the 1024-byte swizzle alignment");
  static_assert(TOTAL <= 232448, "pair GEMM exceeds the 227 KB shared-memory limit");
};

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(Gemm3Cfg<BN>::THREADS, 1)
    gemm_bf16_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                          const __grid_constant__ CUtensorMap tmC, const GemmParams p, const Gemm2Sched sc) {
  pdl_launch_dependents();
  using C = Gemm3Cfg<BN>;
  constexpr int STAGES = C::STAGES;
  static_assert(!B_MN || C::BNH % 64 == 0, "MN-major B halves are made of 64-column chunks");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]  (the leader's copy is the one in use)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      if (p.tma_store) tma_prefetch_desc(&tmC);
    }
  } else if (warp == 1) {
    if (elect_one()) {
#pragma unroll
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full_bar[s], 1);      // leader: its own expect_tx arrival; transaction bytes of both CTAs
        mbar_init(&empty_bar[s], 1);     // one multicast commit per use
      }
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        mbar_init(&tmem_full_bar[b], 1);
        mbar_init(&tmem_empty_bar[b], 2 * C::EPI_WARPS);   // epilogue warps of both CTAs
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_2cta(tmem_ptr_smem, C::TMEM_COLS);   // collective of the pair: one warp in each CTA
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();     // the peer's barriers exist before any remote arrival / transaction lands on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();

  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  auto decode = [&](int item, int& m0, int& n0, int& kb_begin, int& nkb, int& z) {
    z = item % sc.splits;
    const int tile = item / sc.splits;
    n0 = (tile % sc.tiles_n) * BN;
    m0 = (tile / sc.tiles_n) * (2 * BM) + int(rank) * BM;     // this CTA's 128 rows of the 256-row pair tile
    kb_begin = z * p.kb_per_split;
    nkb = min(p.kb_per_split, p.num_kb - kb_begin);
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = cluster_id; item < sc.total; item += num_clusters) {
        int m0, n0, kb_begin, nkb, z;
        decode(item, m0, n0, kb_begin, nkb, z);
        const int nb0 = n0 + int(rank) * C::BNH;             // this CTA's half of the B tile
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * C::STAGE_BYTES);
          const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[stage]), 0);
          uint8_t* sA = smem + stage * C::STAGE_BYTES;
          uint8_t* sB = sA + C::A_BYTES;
          const int k0 = (kb_begin + kb) * BK;
          if constexpr (!A_MN) {
            tma_load_2d_2cta(sA, &tmA, full_leader, k0, m0);
          } else {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) tma_load_2d_2cta(sA + c * (64 * BK * 2), &tmA, full_leader, m0 + c * 64, k0);
          }
          if constexpr (!B_MN) {
            tma_load_2d_2cta(sB, &tmB, full_leader, k0, nb0);
          } else {
#pragma unroll
            for (int c = 0; c < C::BNH / 64; ++c) tma_load_2d_2cta(sB + c * (64 * BK * 2), &tmB, full_leader, nb0 + c * 64, k0);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader && elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(2 * BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int item = cluster_id; item < sc.total; item += num_clusters, ++it) {
        int m0, n0, kb_begin, nkb, z;
        decode(item, m0, n0, kb_begin, nkb, z);
        const int buf = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty_bar[buf], acc_phase ^ 1u);   // both CTAs' epilogues released this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + uint32_t(buf * BN);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t b_addr = a_addr + C::A_BYTES;
#pragma unroll
          for (int j = 0; j < BK / UMMA_K; ++j) {
            const uint64_t da = A_MN ? umma_smem_desc_sw128(a_addr + j * (UMMA_K * 128), 64 * BK * 2, 1024)
                                     : umma_smem_desc_sw128(a_addr + j * (UMMA_K * 2), 16, 1024);
            const uint64_t db = B_MN ? umma_smem_desc_sw128(b_addr + j * (UMMA_K * 128), 64 * BK * 2, 1024)
                                     : umma_smem_desc_sw128(b_addr + j * (UMMA_K * 2), 16, 1024);
            tc_mma_f16_ss_2cta(tmem_d, da, db, idesc, (kb | j) != 0 ? 1u : 0u);
          }
          tc_commit_2cta(&empty_bar[stage], 0x3);          // both producers may refill this stage
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        tc_commit_2cta(&tmem_full_bar[buf], 0x3);          // both epilogues may read their 128 rows
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (both CTAs, own 128 rows)
    const int e = warp - 2;
    const int q = warp & 3;
    constexpr int COLS_PER_WARP = BN / (C::EPI_WARPS / 4);
    const int col0 = (e >> 2) * COLS_PER_WARP;
    int it = 0;
    for (int item = cluster_id; item < sc.total; item += num_clusters, ++it) {
      int m0, n0, kb_begin, nkb, z;
      decode(item, m0, n0, kb_begin, nkb, z);
      const int buf = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      float* sbias = reinterpret_cast<float*>(smem + C::BIAS_OFFSET) + buf * BN;
      const bool use_bias = p.ep.bias != nullptr && z == 0;
      if (p.ep.bias != nullptr) {
        if (use_bias)
          for (int c = e * 32 + lane; c < BN; c += C::EPI_WARPS * 32) sbias[c] = n0 + c < p.N ? __ldg(p.ep.bias + n0 + c) : 0.f;
        asm volatile("bar.sync 1, %0;" ::"n"(C::EPI_WARPS * 32) : "memory");
      }
      mbar_wait(&tmem_full_bar[buf], acc_phase);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.M;
      const bool first_split = z == 0;
      const bool atomic_out = p.ep.accumulate != 0 || sc.splits > 1;
#pragma unroll 1
      for (int c = 0; c < COLS_PER_WARP / 32; ++c) {
        const int nb = n0 + col0 + c * 32;
        if (nb >= p.N) break;
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(buf * BN + col0 + c * 32), r);
        tc_wait_ld();
        if (p.tma_store == 1)
          epilogue_chunk32_tma(r, p, use_bias ? sbias + col0 + c * 32 : nullptr, smem + C::STORE_OFFSET + e * 2048, lane,
                               &tmC, nb, m0 + q * 32);
        else if (p.tma_store != 0)
          epilogue_chunk32_tma_f32(r, p, use_bias ? sbias + col0 + c * 32 : nullptr, smem + C::STORE_OFFSET + e * 2048, lane,
                                   &tmC, nb, m0 + q * 32, p.tma_store == 3);
        else
          epilogue_chunk32(r, row, row_ok, nb, p, first_split, atomic_out, use_bias ? sbias + col0 + c * 32 : nullptr);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty_bar[buf]), 0));   // the leader's barrier
    }
    if (p.tma_store && lane == 0) bulk_wait_all();
  }

  // no CTA of the pair may leave (or free TMEM) while the other can still address its shared memory / TMEM
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2cta(tmem_base, C::TMEM_COLS);
}

template <int BN, bool A_MN, bool B_MN>
int launch_gemm3(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const GemmParams& p, int split_k,
                 cudaStream_t stream) {
  using C = Gemm3Cfg<BN>;
  auto kern = gemm_bf16_pair_kernel<BN, A_MN, B_MN>;
  static bool configured = false;
  if (!configured) {
    MMAE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::TOTAL));
    MMAE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 0));
    configured = true;
  }
  Gemm2Sched sc;
  sc.tiles_m = ceil_div(p.M, 2 * BM);
  sc.tiles_n = ceil_div(p.N, BN);
  sc.splits = split_k;
  sc.total = sc.tiles_m * sc.tiles_n * split_k;
  const int clusters = persistent_grid(sc.total, sm_count() / 2);
  const bool prof = gemm_profile_begin(stream, 2.0 * p.M * p.N * p.K, p.M, p.N, p.K, (A_MN ? 1 : 0) | (B_MN ? 2 : 0) | (split_k << 8));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(C::THREADS);
  cfg.dynamicSmemBytes = C::TOTAL;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  MMAE_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, p, sc));
  if (prof) gemm_profile_end(stream);
  count_launch();
  MMAE_LAUNCH_OK();
  return MMAE_OK;
}

}  // namespace
}  // namespace mmae

using namespace mmae;

static int g_gemm_variant = []() {
  const char* e = getenv("MMAE_GEMM_VARIANT");
  return e ? atoi(e) : -1;
}();

// MMAE_GEMM_TMA_STORE=0 (or mmae_gemm_set_variant(v | 0x100)) falls back to per-lane global stores (A/B measurements)
static int g_gemm_tma_store = []() {
  const char* e = getenv("MMAE_GEMM_TMA_STORE");
  return e ? atoi(e) : 1;
}();

extern "C" int mmae_gemm_set_tma_store(int enable) {
  g_gemm_tma_store = enable != 0;
  return MMAE_OK;
}

// MMAE_GEMM_PAIR=0 keeps the heuristic on the single-CTA kernels
static int g_gemm_pair = []() {
  const char* e = getenv("MMAE_GEMM_PAIR");
  return e ? atoi(e) : 1;
}();

extern "C" int mmae_gemm_set_variant(int variant) {
  g_gemm_variant = variant;
  return MMAE_OK;
}

extern "C" int mmae_gemm_bf16(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb,
                              int b_mn_major, int M, int N, int K, int split_k, const mmae_gemm_epilogue* ep,
                              void* stream) {
  MMAE_CHECK(A && B && ep, MMAE_ERR_ARG, "mmae_gemm_bf16: null operand");
  MMAE_CHECK(M > 0 && N > 0 && K > 0, MMAE_ERR_ARG, "mmae_gemm_bf16: bad shape M=%d N=%d K=%d", M, N, K);
  MMAE_CHECK(N % 8 == 0, MMAE_ERR_ARG, "mmae_gemm_bf16: N=%d must be a multiple of 8", N);
  MMAE_CHECK(lda % 8 == 0 && ldb % 8 == 0, MMAE_ERR_ARG, "mmae_gemm_bf16: lda/ldb must be multiples of 8");
  MMAE_CHECK((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0,
             MMAE_ERR_ARG, "mmae_gemm_bf16: operands must be 16-byte aligned");
  MMAE_CHECK(ep->out_f32 || ep->out_bf16, MMAE_ERR_ARG, "mmae_gemm_bf16: no output");
  const int num_kb = ceil_div(K, BK);
  int variant = g_gemm_variant;
  if (split_k <= 0 && !(ep->out_f32 && !ep->out_bf16 && !ep->preact_bf16 && !ep->dgelu_z && ep->act == 0)) split_k = 1;
  if (split_k <= 0) {
    // auto split-K (weight gradients: small M x N, long K).  Joint choice of tile width and split count: one wave of
    // ~SM-count work items; an item costs ~BN * (k-blocks + 6), the 6 standing for the fp32 reduce-add epilogue.
    // Measured (B200, K = 25088): 256x256 -> BN 128 x 37 splits 11.8 us vs BN 256 x 66 splits 22 us.
    const int sms = sm_count();
    const int tm = ceil_div(M, BM);
    const int cand_bn[3] = {128, 192, 256};
    const int cand_var[3] = {1, 3, 2};
    long best_cost = -1;
    int best_split = 1, best_var = 1;
    for (int i = 0; i < 3; ++i) {
      if (variant > 0 && cand_var[i] != variant) continue;
      if (cand_bn[i] > 128 && N < cand_bn[i]) continue;
      const int tiles = tm * ceil_div(N, cand_bn[i]);
      int s_ = std::max(1, sms / tiles);
      s_ = std::min(s_, std::max(1, num_kb / 4));
      const int kbps = ceil_div(num_kb, s_);
      s_ = ceil_div(num_kb, kbps);
      const long rounds = (long(tiles) * s_ + sms - 1) / sms;
      const long cost = rounds * cand_bn[i] * (kbps + 6);
      if (best_cost < 0 || cost < best_cost) {
        best_cost = cost;
        best_split = s_;
        best_var = cand_var[i];
      }
    }
    split_k = best_split;
    if (variant < 0) variant = best_var;
  }
  if (split_k > num_kb) split_k = num_kb;
  const int kb_per_split = ceil_div(num_kb, split_k);
  split_k = ceil_div(num_kb, kb_per_split);  // no empty splits
  if (split_k > 1 || ep->accumulate) {
    MMAE_CHECK(ep->out_f32 && !ep->out_bf16 && !ep->preact_bf16 && !ep->dgelu_z && ep->act == 0, MMAE_ERR_ARG,
               "mmae_gemm_bf16: split-K / accumulate supports only a linear fp32 epilogue");
  }
#define MMAE_LD_OK(ptr, ld) (!(ptr) || ((ld) % 8 == 0 && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0))
  MMAE_CHECK(MMAE_LD_OK(ep->residual, ep->ld_residual) && MMAE_LD_OK(ep->dgelu_z, ep->ld_dgelu_z) &&
                 MMAE_LD_OK(ep->preact_bf16, ep->ld_preact) && MMAE_LD_OK(ep->out_f32, ep->ld_out_f32) &&
                 MMAE_LD_OK(ep->out_bf16, ep->ld_out_bf16) &&
                 (!ep->bias || (reinterpret_cast<uintptr_t>(ep->bias) & 15) == 0),
             MMAE_ERR_ARG, "mmae_gemm_bf16: epilogue tensors need 16-byte alignment and ld %% 8 == 0");
#undef MMAE_LD_OK

  // kernel variant: 0 = v1 (one tile per CTA, BN=128), 1 / 3 / 2 = persistent BN = 128 / 192 / 256, 6 / 5 / 4 = CTA-pair
  // kernels with 256 x (128 / 192 / 256) tiles.
  // MMAE_GEMM_VARIANT overrides the heuristic (for A/B measurements).
  if (variant < 0) {
    // tile-quantisation model: persistent CTAs process ceil(tiles / SMs) rounds; a round costs ~(BN + c) where c
    // stands for the per-tile fixed work (pipeline fill, epilogue tail).  Pick the cheapest of BN = 128 / 192 / 256.
    const int sms = sm_count();
    const int tm = ceil_div(M, BM);
    int best = 1;
    long best_cost = -1;
    const int cand_bn[3] = {128, 192, 256};
    const int cand_var[3] = {1, 3, 2};
    for (int i = 0; i < 3; ++i) {
      if (cand_bn[i] > 128 && N < cand_bn[i]) continue;
      const long items = long(tm) * ceil_div(N, cand_bn[i]) * split_k;
      const long rounds = (items + sms - 1) / sms;
      const long cost = rounds * (cand_bn[i] + 40);
      if (best_cost < 0 || cost < best_cost) {
        best_cost = cost;
        best = cand_var[i];
      }
    }
    // CTA-pair kernels (256-row tiles, rounds counted in pairs): measured +3..6 % on the big encoder GEMMs (N*K >= 2 M
    // elements: fc1 / fc2 forward, fc2 dgrad), slower on short-K or narrow problems - only those shapes are candidates.
    static const long pair_min_nk = []() {
      const char* e = getenv("MMAE_GEMM_PAIR_NK");
      return e ? atol(e) : 1700000L;   // QKV forward (N*K = 1.77 M): 870 -> 900 TF/s with the pair kernel
    }();
    if (split_k == 1 && g_gemm_pair && M >= 2048 && long(N) * K >= pair_min_nk && sms >= 2) {
      const int tm2 = ceil_div(M, 2 * BM);
      const int pair_bn[2] = {192, 256};
      const int pair_var[2] = {5, 4};
      for (int i = 0; i < 2; ++i) {
        if (N < pair_bn[i] || (pair_bn[i] == 192 && b_mn_major)) continue;
        const long items = long(tm2) * ceil_div(N, pair_bn[i]);
        const long rounds = (items + sms / 2 - 1) / (sms / 2);
        const long cost = rounds * (pair_bn[i] + 40) * 95 / 100;
        if (cost < best_cost) {
          best_cost = cost;
          best = pair_var[i];
        }
      }
    }
    variant = best;
  }
  // variants 4 / 5 / 6: CTA-pair kernels (256 x BN tiles, BN = 256 / 192 / 128); an MN-major B needs 64-column halves
  if (variant == 5 && b_mn_major) variant = 4;
  const bool pair = variant >= 4;
  const int BNsel = (variant == 2 || variant == 4) ? 256 : ((variant == 3 || variant == 5) ? 192 : 128);

  CUtensorMap tmA, tmB;
  int rc;
  if (!a_mn_major) {
    rc = make_tmap_2d_bf16(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BK, BM);
  } else {
    MMAE_CHECK(M % 8 == 0, MMAE_ERR_ARG, "mmae_gemm_bf16: MN-major A needs M %% 8 == 0");
    rc = make_tmap_2d_bf16(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, 64, BK);
  }
  if (rc) return rc;
  if (!b_mn_major) {
    rc = make_tmap_2d_bf16(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, BK, pair ? BNsel / 2 : BNsel);
  } else {
    rc = make_tmap_2d_bf16(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, 64, BK);
  }
  if (rc) return rc;

  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.num_kb = num_kb;
  p.kb_per_split = kb_per_split;
  p.ep = *ep;
  // bf16-only linear / GELU epilogues of the persistent kernels leave through TMA tile stores
  // and so do fp32-only linear epilogues: plain store, or reduce-add for split-K / accumulate
  p.tma_store = 0;
  CUtensorMap tmC = tmA;
  if (g_gemm_tma_store && variant != 0 && !ep->preact_bf16 && !ep->dgelu_z && !ep->residual) {
    if (ep->out_bf16 && !ep->out_f32 && split_k == 1 && !ep->accumulate) {
      p.tma_store = 1;
      rc = make_tmap_2d_store(&tmC, ep->out_bf16, 2, (uint64_t)M, (uint64_t)N, (uint64_t)ep->ld_out_bf16);
      if (rc) return rc;
    } else if (ep->out_f32 && !ep->out_bf16 && ep->act == 0 && ep->ld_out_f32 % 4 == 0) {
      p.tma_store = (split_k > 1 || ep->accumulate) ? 3 : 2;
      rc = make_tmap_2d_store(&tmC, ep->out_f32, 4, (uint64_t)M, (uint64_t)N, (uint64_t)ep->ld_out_f32);
      if (rc) return rc;
    }
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define MMAE_DISPATCH(FN, ...)                                                                   \
  do {                                                                                           \
    if (!a_mn_major && !b_mn_major) return FN<__VA_ARGS__, false, false>(tmA, tmB, tmC, p, split_k, st); \
    if (!a_mn_major && b_mn_major) return FN<__VA_ARGS__, false, true>(tmA, tmB, tmC, p, split_k, st);   \
    if (a_mn_major && !b_mn_major) return FN<__VA_ARGS__, true, false>(tmA, tmB, tmC, p, split_k, st);   \
    return FN<__VA_ARGS__, true, true>(tmA, tmB, tmC, p, split_k, st);                                  \
  } while (0)
  if (variant == 4) MMAE_DISPATCH(launch_gemm3, 256);
  if (variant == 6) MMAE_DISPATCH(launch_gemm3, 128);
  if (variant == 5) {
    if (!a_mn_major) return launch_gemm3<192, false, false>(tmA, tmB, tmC, p, split_k, st);
    return launch_gemm3<192, true, false>(tmA, tmB, tmC, p, split_k, st);
  }
  if (variant == 2) MMAE_DISPATCH(launch_gemm2, 256);
  if (variant == 3) MMAE_DISPATCH(launch_gemm2, 192);
  if (variant == 1) MMAE_DISPATCH(launch_gemm2, 128);
  MMAE_DISPATCH(launch_gemm1, 128);
#undef MMAE_DISPATCH
}

// tcgen05 fused attention for short sequences (Nq, Nk <= 128, head_dim 64): the encoder MHSA of MultiMAE-B/L
// (99 x 99 x 64 per (batch, head) problem), forward and backward.  One CTA per (b, h); the whole problem lives on chip:
//
//   forward : TMA (3D map, zero-filled past the sequence end) -> Q, K, V tiles in 128B-swizzled smem
//             S = Q K^T      tcgen05.mma 128 x Nk16 x 64, fp32 accumulator in TMEM columns [0,128)
//             softmax        one thread per query row: tcgen05.ld, exp2, P -> bf16 K-major swizzled smem (aliases Q|K)
//             O = P V        tcgen05.mma 128 x 64 x Nk16 (V consumed as an MN-major operand straight from its row-major
//                            tile), accumulator aliases the S columns; epilogue scales by 1/l and stores bf16 rows
//   backward: S and dP = dO V^T recomputed on the tensor cores, P/dS written once to smem and consumed by three more
//             UMMAs in two roles: K-major A (dQ = dS K) and MN-major A (dV = P^T dO, dK = dS^T Q) -- the same bytes.
//
// Replaces multimae/multimae_utils.py:175-179 (q@k^T*scale -> softmax -> @v) and its autograd backward for the
// encoder; other shapes (decoder head_dim 32, longer sequences) run on the mma.sync kernels in attention.cu.
#include "common.cuh"
#include "../../include/multimae_b200.h"

namespace mmae {
void count_launch();
namespace {

constexpr int TQ = 128;          // query rows per CTA (one TMEM lane each)
constexpr int DH = 64;           // head dim: 64 bf16 = one 128-byte swizzle row
constexpr int TILE_BYTES = TQ * DH * 2;   // 16 KB: a [128 x 64] bf16 tile
constexpr float LOG2E_F = 1.4426950408889634f;

// byte offset of element (row r, 16-byte chunk j) inside a [rows x 128 B] SWIZZLE_128B tile
__device__ __forceinline__ uint32_t swz128(int r, int j) { return uint32_t(r) * 128u + uint32_t((j ^ (r & 7)) << 4); }

// write 32 consecutive bf16 values (already packed 2 per u32) of row r, columns [c0, c0+32) of a K-major tile made of
// 64-column panels of TILE_BYTES each
__device__ __forceinline__ void store_row32(uint8_t* tile, int r, int c0, const uint32_t (&pk)[16]) {
  uint8_t* panel = tile + (c0 >> 6) * TILE_BYTES;
  const int j0 = (c0 & 63) >> 3;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 v = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
    *reinterpret_cast<uint4*>(panel + swz128(r, j0 + j)) = v;
  }
}

struct AttnTcParams {
  int Nq, Nk, H;
  float scale;
  bf16* O;
  int64_t ldo;
  float* lse;
};

__global__ void __launch_bounds__(128) attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                          const __grid_constant__ CUtensorMap tmK,
                                                          const __grid_constant__ CUtensorMap tmV, const AttnTcParams p) {
  pdl_prologue();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQ = smem;                      // 16 KB   } P (two 64-key panels, 32 KB) aliases Q|K once S is complete
  uint8_t* sK = smem + TILE_BYTES;         // 16 KB   }
  uint8_t* sV = smem + 2 * TILE_BYTES;     // 16 KB
  uint8_t* sP = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 3 * TILE_BYTES);   // [0] loads, [1] S ready, [2] O ready
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 3);

  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk16 = (p.Nk + 15) & ~15;      // UMMA N of S / K extent of P V

  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      mbar_init(&bars[0], 1);
      mbar_init(&bars[1], 1);
      mbar_init(&bars[2], 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr_smem, 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_smem;

  if (threadIdx.x == 0) {
    mbar_expect_tx(&bars[0], 3 * TILE_BYTES);
    tma_load_3d(sQ, &tmQ, &bars[0], h * DH, 0, b);
    tma_load_3d(sK, &tmK, &bars[0], h * DH, 0, b);
    tma_load_3d(sV, &tmV, &bars[0], h * DH, 0, b);
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    // S[128 x nk16] = Q K^T : both operands K-major, K = 64 -> 4 UMMA steps
    const uint32_t idesc = umma_idesc_bf16(128, nk16, 0, 0);
#pragma unroll
    for (int j = 0; j < DH / 16; ++j) {
      const uint64_t da = umma_smem_desc_sw128(smem_u32(sQ) + j * 32, 16, 1024);
      const uint64_t db = umma_smem_desc_sw128(smem_u32(sK) + j * 32, 16, 1024);
      tc_mma_f16_ss(tmem, da, db, idesc, j != 0 ? 1u : 0u);
    }
    tc_commit(&bars[1]);
  }
  __syncwarp();

  // ---------------------------------------------------------------- softmax: thread = query row
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  const uint32_t trow = tmem + (uint32_t(warp * 32) << 16);
  const float sl2 = p.scale * LOG2E_F;
  float mx = -INFINITY;
  for (int c0 = 0; c0 < nk16; c0 += 32) {
    uint32_t r[32];
    tmem_ld_32x32(trow + c0, r);
    tc_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (c0 + i < p.Nk) mx = fmaxf(mx, __uint_as_float(r[i]));
  }
  const float moff = mx * sl2;
  float sum = 0.f;
  for (int c0 = 0; c0 < 128; c0 += 32) {
    uint32_t pk[16];
    if (c0 < nk16) {
      uint32_t r[32];
      tmem_ld_32x32(trow + c0, r);
      tc_wait_ld();
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float a = c0 + 2 * i < p.Nk ? fast_exp2(__uint_as_float(r[2 * i]) * sl2 - moff) : 0.f;
        const float c = c0 + 2 * i + 1 < p.Nk ? fast_exp2(__uint_as_float(r[2 * i + 1]) * sl2 - moff) : 0.f;
        sum += a + c;
        pk[i] = pack_bf16x2(a, c);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) pk[i] = 0u;
    }
    // the S reads of every row must be finished before P overwrites the Q|K tiles that MMA1 consumed: MMA1 has
    // completed (bars[1]); P rows are private to this thread, so no further hazard
    store_row32(sP, row, c0, pk);
  }
  fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the tensor-core (async) proxy
  tc_fence_before();
  __syncthreads();

  if (threadIdx.x == 0) {
    tc_fence_after();
    // O[128 x 64] = P V : A = P (K-major, 64-key panels), B = V (MN-major: row-major [key][dh] tile as is)
    const uint32_t idesc = umma_idesc_bf16(128, DH, 0, 1);
    const int ksteps = nk16 / 16;
    for (int j = 0; j < ksteps; ++j) {
      const uint32_t a_addr = smem_u32(sP) + (j >> 2) * TILE_BYTES + (j & 3) * 32;
      const uint64_t da = umma_smem_desc_sw128(a_addr, 16, 1024);
      const uint64_t db = umma_smem_desc_sw128(smem_u32(sV) + j * (16 * 128), TILE_BYTES, 1024);
      tc_mma_f16_ss(tmem, da, db, idesc, j != 0 ? 1u : 0u);
    }
    tc_commit(&bars[2]);
  }
  __syncwarp();

  // ---------------------------------------------------------------- epilogue
  mbar_wait(&bars[2], 0);
  tc_fence_after();
  const float inv = 1.0f / sum;
#pragma unroll
  for (int c0 = 0; c0 < DH; c0 += 32) {
    uint32_t r[32];
    tmem_ld_32x32(trow + c0, r);
    tc_wait_ld();
    if (row < p.Nq) {
      bf16* dst = p.O + (int64_t(b) * p.Nq + row) * p.ldo + h * DH + c0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 v;
        v.x = pack_bf16x2(__uint_as_float(r[8 * j + 0]) * inv, __uint_as_float(r[8 * j + 1]) * inv);
        v.y = pack_bf16x2(__uint_as_float(r[8 * j + 2]) * inv, __uint_as_float(r[8 * j + 3]) * inv);
        v.z = pack_bf16x2(__uint_as_float(r[8 * j + 4]) * inv, __uint_as_float(r[8 * j + 5]) * inv);
        v.w = pack_bf16x2(__uint_as_float(r[8 * j + 6]) * inv, __uint_as_float(r[8 * j + 7]) * inv);
        *reinterpret_cast<uint4*>(dst + 8 * j) = v;
      }
    }
  }
  if (p.lse != nullptr && row < p.Nq) p.lse[(int64_t(b) * p.H + h) * p.Nq + row] = mx * p.scale + logf(sum);

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

// =====================================================================================================================
// forward, persistent variant - EXPERIMENTAL, NOT YET RUN ON HARDWARE (written after round 1's GPU budget was spent).
// Selected only by bit 4 of the switch (mmae_attention_set_tc(3 | 16) / MMAE_ATTN_TC=19); the default stays
// attn_tc_fwd_kernel.  scripts/gpu_check_attention_v2.py compares and times the two.
//
// Same arithmetic as attn_tc_fwd_kernel.  What changes is the schedule: 2 CTAs per SM each loop over (batch, head) items
// with a 2-stage Q/K/V ring, so the TMA of item i+1 is in flight while item i runs S -> softmax -> P V -> store, and the
// per-CTA set-up (TMEM allocation, barrier init, tensor-map fetch) is paid once per CTA instead of once per item; 1536 items
// are dealt over 296 CTAs instead of running as 2.6 waves of 592.  S / O share one 128-column TMEM accumulator, as before.
// =====================================================================================================================
constexpr int FWD2_STAGE_BYTES = 3 * TILE_BYTES;
constexpr int FWD2_SMEM = 2 * FWD2_STAGE_BYTES + 128 + 1024;

__global__ void __launch_bounds__(128) attn_tc_fwd_persistent_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                     const __grid_constant__ CUtensorMap tmK,
                                                                     const __grid_constant__ CUtensorMap tmV,
                                                                     const AttnTcParams p, const int items) {
  pdl_prologue();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * FWD2_STAGE_BYTES);   // [0],[1] stage loaded, [2] S ready, [3] O ready
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk16 = (p.Nk + 15) & ~15;

  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
#pragma unroll
      for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr_smem, 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_smem;

  auto issue_loads = [&](int item, int stage) {
    const int h = item % p.H, b = item / p.H;
    uint8_t* base = smem + stage * FWD2_STAGE_BYTES;
    mbar_expect_tx(&bars[stage], 3 * TILE_BYTES);
    tma_load_3d(base, &tmQ, &bars[stage], h * DH, 0, b);
    tma_load_3d(base + TILE_BYTES, &tmK, &bars[stage], h * DH, 0, b);
    tma_load_3d(base + 2 * TILE_BYTES, &tmV, &bars[stage], h * DH, 0, b);
  };
  if (threadIdx.x == 0 && int(blockIdx.x) < items) issue_loads(blockIdx.x, 0);

  const int row = warp * 32 + lane;
  const uint32_t trow = tmem + (uint32_t(warp * 32) << 16);
  const float sl2 = p.scale * LOG2E_F;

  int it = 0;
  for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
    const int stage = it & 1;
    const uint32_t ld_phase = uint32_t(it >> 1) & 1u;     // each stage barrier completes once per use of the stage
    const uint32_t ph = uint32_t(it) & 1u;                // bars[2] / bars[3] complete once per item
    const int h = item % p.H, b = item / p.H;
    uint8_t* sQ = smem + stage * FWD2_STAGE_BYTES;
    uint8_t* sK = sQ + TILE_BYTES;
    uint8_t* sV = sQ + 2 * TILE_BYTES;
    uint8_t* sP = sQ;                                     // P (two 64-key panels) aliases Q|K once S is complete

    if (threadIdx.x == 0) {
      // prefetch the next item into the other stage: its last readers (the P V MMAs of item it-1) completed before every
      // thread passed the barrier that ended iteration it-1
      const int next = item + int(gridDim.x);
      if (next < items) issue_loads(next, stage ^ 1);
      mbar_wait(&bars[stage], ld_phase);
      tc_fence_after();
      const uint32_t idesc = umma_idesc_bf16(128, nk16, 0, 0);
#pragma unroll
      for (int j = 0; j < DH / 16; ++j) {
        const uint64_t da = umma_smem_desc_sw128(smem_u32(sQ) + j * 32, 16, 1024);
        const uint64_t db = umma_smem_desc_sw128(smem_u32(sK) + j * 32, 16, 1024);
        tc_mma_f16_ss(tmem, da, db, idesc, j != 0 ? 1u : 0u);
      }
      tc_commit(&bars[2]);
    }
    __syncwarp();

    // ---------------------------------------------------------------- softmax: thread = query row
    mbar_wait(&bars[2], ph);
    tc_fence_after();
    float mx = -INFINITY;
    for (int c0 = 0; c0 < nk16; c0 += 32) {
      uint32_t r[32];
      tmem_ld_32x32(trow + c0, r);
      tc_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (c0 + i < p.Nk) mx = fmaxf(mx, __uint_as_float(r[i]));
    }
    const float moff = mx * sl2;
    float sum = 0.f;
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t pk[16];
      if (c0 < nk16) {
        uint32_t r[32];
        tmem_ld_32x32(trow + c0, r);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float a = c0 + 2 * i < p.Nk ? fast_exp2(__uint_as_float(r[2 * i]) * sl2 - moff) : 0.f;
          const float c = c0 + 2 * i + 1 < p.Nk ? fast_exp2(__uint_as_float(r[2 * i + 1]) * sl2 - moff) : 0.f;
          sum += a + c;
          pk[i] = pack_bf16x2(a, c);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = 0u;
      }
      store_row32(sP, row, c0, pk);                       // Q|K are dead: the S MMAs completed (bars[2])
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();

    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t idesc = umma_idesc_bf16(128, DH, 0, 1);
      const int ksteps = nk16 / 16;
      for (int j = 0; j < ksteps; ++j) {
        const uint32_t a_addr = smem_u32(sP) + (j >> 2) * TILE_BYTES + (j & 3) * 32;
        const uint64_t da = umma_smem_desc_sw128(a_addr, 16, 1024);
        const uint64_t db = umma_smem_desc_sw128(smem_u32(sV) + j * (16 * 128), TILE_BYTES, 1024);
        tc_mma_f16_ss(tmem, da, db, idesc, j != 0 ? 1u : 0u);
      }
      tc_commit(&bars[3]);
    }
    __syncwarp();

    // ---------------------------------------------------------------- epilogue
    mbar_wait(&bars[3], ph);
    tc_fence_after();
    const float inv = 1.0f / sum;
#pragma unroll
    for (int c0 = 0; c0 < DH; c0 += 32) {
      uint32_t r[32];
      tmem_ld_32x32(trow + c0, r);
      tc_wait_ld();
      if (row < p.Nq) {
        bf16* dst = p.O + (int64_t(b) * p.Nq + row) * p.ldo + h * DH + c0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(r[8 * j + 0]) * inv, __uint_as_float(r[8 * j + 1]) * inv);
          v.y = pack_bf16x2(__uint_as_float(r[8 * j + 2]) * inv, __uint_as_float(r[8 * j + 3]) * inv);
          v.z = pack_bf16x2(__uint_as_float(r[8 * j + 4]) * inv, __uint_as_float(r[8 * j + 5]) * inv);
          v.w = pack_bf16x2(__uint_as_float(r[8 * j + 6]) * inv, __uint_as_float(r[8 * j + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + 8 * j) = v;
        }
      }
    }
    if (p.lse != nullptr && row < p.Nq) p.lse[(int64_t(b) * p.H + h) * p.Nq + row] = mx * p.scale + logf(sum);
    // every TMEM read of this item is done before the next item's S MMAs overwrite the accumulator, and this stage's
    // shared memory may be refilled by the TMA issued at the top of the next-but-one iteration
    tc_fence_before();
    __syncthreads();
  }

  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 128);
  }
}

// =====================================================================================================================
// backward
// =====================================================================================================================
struct AttnTcBwdParams {
  int Nq, Nk, H;
  float scale;
  const float* lse;
  const float* delta;
  bf16 *dQ, *dK, *dV;
  int64_t lddq, lddk, lddv;
};

// Two CTAs per SM: 256 TMEM columns and 7 smem tiles each.
// TMEM columns: S [0,128)  dP [128,256); once P and dS are in shared memory the gradient accumulators reuse them:
//               dV [0,64)  dK [64,128)  dQ [128,192).
// smem tiles  : Q | K | dO | P panel 0 | V, later P panel 1 (V is dead once dP is complete) | dS panel 0 | dS panel 1
constexpr int ATTN_BWD_SMEM = 7 * TILE_BYTES + 64 + 960;   // 2 x (this + 1 KB system reserve) = the SM's 228 KB

__global__ void __launch_bounds__(128, 2) attn_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                             const __grid_constant__ CUtensorMap tmK,
                                                             const __grid_constant__ CUtensorMap tmV,
                                                             const __grid_constant__ CUtensorMap tmdO,
                                                             const AttnTcBwdParams p) {
  pdl_prologue();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  if (pad > 960u) __trap();                // the launch reserves 960 bytes of alignment slack
  uint8_t* smem = smem_raw + pad;
  uint8_t* sQ = smem;
  uint8_t* sK = smem + 1 * TILE_BYTES;
  uint8_t* sdO = smem + 2 * TILE_BYTES;
  uint8_t* sP = smem + 3 * TILE_BYTES;     // 32 KB: [q][key] bf16, two 64-key panels; panel 1 overlays V
  uint8_t* sV = smem + 4 * TILE_BYTES;
  uint8_t* sdS = smem + 5 * TILE_BYTES;    // 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 7 * TILE_BYTES);   // [0] loads, [1] S & dP ready, [2] grads ready
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 3);

  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk16 = (p.Nk + 15) & ~15;
  const int nq16 = (p.Nq + 15) & ~15;

  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      tma_prefetch_desc(&tmdO);
      mbar_init(&bars[0], 1);
      mbar_init(&bars[1], 1);
      mbar_init(&bars[2], 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr_smem, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_smem;

  if (threadIdx.x == 0) {
    mbar_expect_tx(&bars[0], 4 * TILE_BYTES);
    tma_load_3d(sQ, &tmQ, &bars[0], h * DH, 0, b);
    tma_load_3d(sK, &tmK, &bars[0], h * DH, 0, b);
    tma_load_3d(sV, &tmV, &bars[0], h * DH, 0, b);
    tma_load_3d(sdO, &tmdO, &bars[0], h * DH, 0, b);
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16(128, nk16, 0, 0);
#pragma unroll
    for (int j = 0; j < DH / 16; ++j) {   // S = Q K^T
      tc_mma_f16_ss(tmem, umma_smem_desc_sw128(smem_u32(sQ) + j * 32, 16, 1024),
                    umma_smem_desc_sw128(smem_u32(sK) + j * 32, 16, 1024), idesc, j != 0 ? 1u : 0u);
    }
#pragma unroll
    for (int j = 0; j < DH / 16; ++j) {   // dP = dO V^T
      tc_mma_f16_ss(tmem + 128, umma_smem_desc_sw128(smem_u32(sdO) + j * 32, 16, 1024),
                    umma_smem_desc_sw128(smem_u32(sV) + j * 32, 16, 1024), idesc, j != 0 ? 1u : 0u);
    }
    tc_commit(&bars[1]);
  }
  __syncwarp();

  // ---------------------------------------------------------------- P and dS: thread = query row
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  const uint32_t trow = tmem + (uint32_t(warp * 32) << 16);
  const float sl2 = p.scale * LOG2E_F;
  const bool row_ok = row < p.Nq;
  const float lse2 = row_ok ? p.lse[(int64_t(b) * p.H + h) * p.Nq + row] * LOG2E_F : INFINITY;   // +inf -> P = 0
  const float del = row_ok ? p.delta[(int64_t(b) * p.H + h) * p.Nq + row] : 0.f;
  for (int c0 = 0; c0 < 128; c0 += 32) {
    uint32_t pp[16], ds[16];
    if (c0 < nk16) {
      uint32_t s[32], d[32];
      tmem_ld_32x32(trow + c0, s);
      tmem_ld_32x32(trow + 128 + c0, d);
      tc_wait_ld();
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float p0 = c0 + 2 * i < p.Nk ? fast_exp2(__uint_as_float(s[2 * i]) * sl2 - lse2) : 0.f;
        const float p1 = c0 + 2 * i + 1 < p.Nk ? fast_exp2(__uint_as_float(s[2 * i + 1]) * sl2 - lse2) : 0.f;
        pp[i] = pack_bf16x2(p0, p1);
        ds[i] = pack_bf16x2(p0 * (__uint_as_float(d[2 * i]) - del), p1 * (__uint_as_float(d[2 * i + 1]) - del));
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) pp[i] = ds[i] = 0u;
    }
    store_row32(sP, row, c0, pp);
    store_row32(sdS, row, c0, ds);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();

  if (threadIdx.x == 0) {
    tc_fence_after();
    // dV[key x dh] = P^T dO : A = P as MN-major (rows = keys; 64-key panels are TILE_BYTES apart), B = dO MN-major
    // dK[key x dh] = dS^T Q : same with dS / Q.    K extent = queries (nq16), 16 query rows (2048 B) per UMMA step
    const uint32_t idesc_t = umma_idesc_bf16(128, DH, 1, 1);
    for (int j = 0; j < nq16 / 16; ++j) {
      const uint64_t db_do = umma_smem_desc_sw128(smem_u32(sdO) + j * 2048, TILE_BYTES, 1024);
      const uint64_t db_q = umma_smem_desc_sw128(smem_u32(sQ) + j * 2048, TILE_BYTES, 1024);
      tc_mma_f16_ss(tmem + 0, umma_smem_desc_sw128(smem_u32(sP) + j * 2048, TILE_BYTES, 1024), db_do, idesc_t,
                    j != 0 ? 1u : 0u);
      tc_mma_f16_ss(tmem + 64, umma_smem_desc_sw128(smem_u32(sdS) + j * 2048, TILE_BYTES, 1024), db_q, idesc_t,
                    j != 0 ? 1u : 0u);
    }
    // dQ[q x dh] = dS K : A = dS K-major, B = K MN-major; K extent = keys
    const uint32_t idesc_q = umma_idesc_bf16(128, DH, 0, 1);
    for (int j = 0; j < nk16 / 16; ++j) {
      const uint32_t a_addr = smem_u32(sdS) + (j >> 2) * TILE_BYTES + (j & 3) * 32;
      tc_mma_f16_ss(tmem + 128, umma_smem_desc_sw128(a_addr, 16, 1024),
                    umma_smem_desc_sw128(smem_u32(sK) + j * 2048, TILE_BYTES, 1024), idesc_q, j != 0 ? 1u : 0u);
    }
    tc_commit(&bars[2]);
  }
  __syncwarp();

  // ---------------------------------------------------------------- epilogue: rows = keys for dV/dK, queries for dQ
  mbar_wait(&bars[2], 0);
  tc_fence_after();
#pragma unroll
  for (int which = 0; which < 3; ++which) {
    const int limit = which == 2 ? p.Nq : p.Nk;
    bf16* base = which == 0 ? p.dV : (which == 1 ? p.dK : p.dQ);
    const int64_t ld = which == 0 ? p.lddv : (which == 1 ? p.lddk : p.lddq);
    const float mul = which == 0 ? 1.0f : p.scale;
#pragma unroll
    for (int c0 = 0; c0 < DH; c0 += 32) {
      uint32_t r[32];
      tmem_ld_32x32(trow + which * 64 + c0, r);
      tc_wait_ld();
      if (row < limit) {
        bf16* dst = base + (int64_t(b) * limit + row) * ld + h * DH + c0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(r[8 * j + 0]) * mul, __uint_as_float(r[8 * j + 1]) * mul);
          v.y = pack_bf16x2(__uint_as_float(r[8 * j + 2]) * mul, __uint_as_float(r[8 * j + 3]) * mul);
          v.z = pack_bf16x2(__uint_as_float(r[8 * j + 4]) * mul, __uint_as_float(r[8 * j + 5]) * mul);
          v.w = pack_bf16x2(__uint_as_float(r[8 * j + 6]) * mul, __uint_as_float(r[8 * j + 7]) * mul);
          *reinterpret_cast<uint4*>(dst + 8 * j) = v;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}


// =====================================================================================================================
// general forward: any Nq (128-row query tiles), Nk <= 256, head_dim 64 or 32.
// head_dim 32: smem tiles are still 64 columns wide (one 128-byte swizzle row) and hold a PAIR of heads; S uses the
// 32-column half of the pair as its K extent (descriptor start offset +64 B inside the swizzled row); P V is computed
// for the whole 64-column tile (16 extra UMMA columns-worth, negligible) and only the head's half is stored.
// Keys are processed as S[128 x nk16] in one UMMA (N <= 256); P is produced and consumed 128 keys at a time through a
// 32 KB buffer that aliases the (dead) Q|K tiles.
// =====================================================================================================================
template <int HD, int KBOX>   // KBOX: number of 128-row K/V boxes (1: Nk <= 128, 2: Nk <= 256)
__global__ void __launch_bounds__(128) attn_tc_fwd_gen_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                              const __grid_constant__ CUtensorMap tmK,
                                                              const __grid_constant__ CUtensorMap tmV,
                                                              const AttnTcParams p) {
  pdl_prologue();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQ = smem;                                  // 16 KB
  uint8_t* sK = smem + TILE_BYTES;                     // KBOX * 16 KB
  uint8_t* sV = smem + (1 + KBOX) * TILE_BYTES;        // KBOX * 16 KB
  uint8_t* sP = smem;                                  // 32 KB, aliases Q | K[0] once S is complete
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (1 + 2 * KBOX) * TILE_BYTES);   // [0] loads [1] S [2],[3] PV halves
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 4);
  constexpr uint32_t TMEM_COLS = KBOX == 1 ? 128 : 256;
  static_assert(KBOX == 1 || KBOX == 2, "KBOX");
  // with KBOX == 1 the P buffer (32 KB) would overrun Q|K (32 KB) exactly; fine.  With KBOX == 2: Q|K = 48 KB.

  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk16 = (p.Nk + 15) & ~15;
  const int tile_col = HD == 64 ? h * 64 : (h >> 1) * 64;     // first column of the 64-wide smem tile
  const int sub_off = HD == 64 ? 0 : (h & 1) * 64;            // byte offset of this head inside the tile row

  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
#pragma unroll
      for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr_smem, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_smem;

  if (threadIdx.x == 0) {
    mbar_expect_tx(&bars[0], (1 + 2 * KBOX) * TILE_BYTES);
    tma_load_3d(sQ, &tmQ, &bars[0], tile_col, qt * TQ, b);
#pragma unroll
    for (int kb = 0; kb < KBOX; ++kb) {
      tma_load_3d(sK + kb * TILE_BYTES, &tmK, &bars[0], tile_col, kb * TQ, b);
      tma_load_3d(sV + kb * TILE_BYTES, &tmV, &bars[0], tile_col, kb * TQ, b);
    }
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16(128, nk16, 0, 0);
#pragma unroll
    for (int j = 0; j < HD / 16; ++j) {
      const uint64_t da = umma_smem_desc_sw128(smem_u32(sQ) + sub_off + j * 32, 16, 1024);
      const uint64_t db = umma_smem_desc_sw128(smem_u32(sK) + sub_off + j * 32, 16, 1024);
      tc_mma_f16_ss(tmem, da, db, idesc, j != 0 ? 1u : 0u);
    }
    tc_commit(&bars[1]);
  }
  __syncwarp();

  mbar_wait(&bars[1], 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  const int qrow = qt * TQ + row;
  const uint32_t trow = tmem + (uint32_t(warp * 32) << 16);
  const float sl2 = p.scale * LOG2E_F;
  float mx = -INFINITY;
  for (int c0 = 0; c0 < nk16; c0 += 32) {
    uint32_t r[32];
    tmem_ld_32x32(trow + c0, r);
    tc_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (c0 + i < p.Nk) mx = fmaxf(mx, __uint_as_float(r[i]));
  }
  const float moff = mx * sl2;
  float sum = 0.f;
#pragma unroll 1
  for (int half = 0; half < KBOX; ++half) {
    const int kbase = half * 128;
    if (kbase >= nk16) break;                         // block-uniform
    if (half > 0) {                                   // previous P V must have consumed sP before it is rewritten
      mbar_wait(&bars[2], 0);
      tc_fence_after();
    }
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t pk[16];
      if (kbase + c0 < nk16) {
        uint32_t r[32];
        tmem_ld_32x32(trow + kbase + c0, r);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int k0 = kbase + c0 + 2 * i;
          const float a = k0 < p.Nk ? fast_exp2(__uint_as_float(r[2 * i]) * sl2 - moff) : 0.f;
          const float c = k0 + 1 < p.Nk ? fast_exp2(__uint_as_float(r[2 * i + 1]) * sl2 - moff) : 0.f;
          sum += a + c;
          pk[i] = pack_bf16x2(a, c);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = 0u;
      }
      store_row32(sP, row, c0, pk);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after();
      // O[128 x 64] (+)= P[:, kbase : kbase+128] V[kbase : kbase+128, tile]; accumulator aliases S columns [0, 64),
      // which every thread has finished reading (first half) / which the second half never reads
      const uint32_t idesc = umma_idesc_bf16(128, 64, 0, 1);
      const int ksteps = min(nk16 - kbase, 128) / 16;
      for (int j = 0; j < ksteps; ++j) {
        const uint32_t a_addr = smem_u32(sP) + (j >> 2) * TILE_BYTES + (j & 3) * 32;
        const uint64_t da = umma_smem_desc_sw128(a_addr, 16, 1024);
        const uint64_t db = umma_smem_desc_sw128(smem_u32(sV) + half * TILE_BYTES + j * (16 * 128), TILE_BYTES, 1024);
        tc_mma_f16_ss(tmem, da, db, idesc, (half | j) != 0 ? 1u : 0u);
      }
      tc_commit(&bars[2 + half]);
    }
    __syncwarp();
  }
  const int last = (nk16 > 128 && KBOX == 2) ? 1 : 0;
  mbar_wait(&bars[2 + last], 0);
  tc_fence_after();
  const float inv = 1.0f / sum;
  constexpr int OUT_COLS = HD;                         // columns of the 64-wide accumulator that belong to this head
  const int ocol0 = HD == 64 ? 0 : (h & 1) * 32;
#pragma unroll
  for (int c0 = 0; c0 < OUT_COLS; c0 += 32) {
    uint32_t r[32];
    tmem_ld_32x32(trow + ocol0 + c0, r);
    tc_wait_ld();
    if (qrow < p.Nq) {
      bf16* dst = p.O + (int64_t(b) * p.Nq + qrow) * p.ldo + h * HD + c0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 v;
        v.x = pack_bf16x2(__uint_as_float(r[8 * j + 0]) * inv, __uint_as_float(r[8 * j + 1]) * inv);
        v.y = pack_bf16x2(__uint_as_float(r[8 * j + 2]) * inv, __uint_as_float(r[8 * j + 3]) * inv);
        v.z = pack_bf16x2(__uint_as_float(r[8 * j + 4]) * inv, __uint_as_float(r[8 * j + 5]) * inv);
        v.w = pack_bf16x2(__uint_as_float(r[8 * j + 6]) * inv, __uint_as_float(r[8 * j + 7]) * inv);
        *reinterpret_cast<uint4*>(dst + 8 * j) = v;
      }
    }
  }
  if (p.lse != nullptr && qrow < p.Nq) p.lse[(int64_t(b) * p.H + h) * p.Nq + qrow] = mx * p.scale + logf(sum);

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

template <int HD, int KBOX>
int launch_fwd_gen(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnTcParams& p, int B,
                   cudaStream_t st) {
  constexpr int SMEM = (1 + 2 * KBOX) * TILE_BYTES + 64 + 1024;
  auto kern = attn_tc_fwd_gen_kernel<HD, KBOX>;
  static bool configured = false;
  if (!configured) {
    MMAE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  launch_k(kern, dim3(ceil_div(p.Nq, TQ), p.H, B), 128, SMEM, st, tq, tk, tv, p);
  count_launch();
  MMAE_LAUNCH_OK();
  return MMAE_OK;
}

// =====================================================================================================================
// general backward (any Nq / Nk, head_dim 64 or 32 via head pairs): two kernels, no atomics.
//   dKV kernel: CTA = (128-key tile, head, batch), loops over 128-query tiles:   dV += P^T dO,  dK += dS^T Q
//   dQ  kernel: CTA = (128-query tile, head, batch), loops over 128-key tiles:    dQ += dS K
// S = Q K^T and dP = dO V^T are recomputed on the tensor cores in both (cheap next to the HBM round trip they avoid).
// =====================================================================================================================
struct AttnTcBwdGenParams {
  int Nq, Nk, H;
  float scale;
  const float* lse;
  const float* delta;
  bf16 *dQ, *dK, *dV;
  int64_t lddq, lddk, lddv;
};

// P / dS for the 128 x 128 tile held in TMEM (S at column 0, dP at column 128); thread = query row
__device__ __forceinline__ void bwd_tile_elementwise(uint32_t trow, int row, int nk_valid, int nk16, float sl2, float lse2,
                                                     float del, uint8_t* sP, uint8_t* sdS) {
  for (int c0 = 0; c0 < 128; c0 += 32) {
    uint32_t pp[16], ds[16];
    if (c0 < nk16) {
      uint32_t s[32], d[32];
      tmem_ld_32x32(trow + c0, s);
      tmem_ld_32x32(trow + 128 + c0, d);
      tc_wait_ld();
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float p0 = c0 + 2 * i < nk_valid ? fast_exp2(__uint_as_float(s[2 * i]) * sl2 - lse2) : 0.f;
        const float p1 = c0 + 2 * i + 1 < nk_valid ? fast_exp2(__uint_as_float(s[2 * i + 1]) * sl2 - lse2) : 0.f;
        pp[i] = pack_bf16x2(p0, p1);
        ds[i] = pack_bf16x2(p0 * (__uint_as_float(d[2 * i]) - del), p1 * (__uint_as_float(d[2 * i + 1]) - del));
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) pp[i] = ds[i] = 0u;
    }
    if (sP != nullptr) store_row32(sP, row, c0, pp);
    store_row32(sdS, row, c0, ds);
  }
}

template <int HD>
__global__ void __launch_bounds__(128) attn_tc_bwd_dkv_gen_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                  const __grid_constant__ CUtensorMap tmK,
                                                                  const __grid_constant__ CUtensorMap tmV,
                                                                  const __grid_constant__ CUtensorMap tmdO,
                                                                  const AttnTcBwdGenParams p) {
  pdl_prologue();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sK = smem;
  uint8_t* sV = smem + 1 * TILE_BYTES;
  uint8_t* sQ = smem + 2 * TILE_BYTES;
  uint8_t* sdO = smem + 3 * TILE_BYTES;
  uint8_t* sP = smem + 4 * TILE_BYTES;     // 32 KB
  uint8_t* sdS = smem + 6 * TILE_BYTES;    // 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 8 * TILE_BYTES);   // [0] K,V  [1] Q,dO  [2] S,dP  [3] dV,dK step
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 4);

  const int kt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_col = HD == 64 ? h * 64 : (h >> 1) * 64;
  const int sub_off = HD == 64 ? 0 : (h & 1) * 64;
  const int nk_valid = min(p.Nk - kt * TQ, TQ);       // keys of this tile
  const int nk16 = (nk_valid + 15) & ~15;
  const int q_tiles = (p.Nq + TQ - 1) / TQ;

  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      tma_prefetch_desc(&tmdO);
#pragma unroll
      for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_smem;
  const int row = warp * 32 + lane;
  const uint32_t trow = tmem + (uint32_t(warp * 32) << 16);
  const float sl2 = p.scale * LOG2E_F;

  if (threadIdx.x == 0) {
    mbar_expect_tx(&bars[0], 2 * TILE_BYTES);
    tma_load_3d(sK, &tmK, &bars[0], tile_col, kt * TQ, b);
    tma_load_3d(sV, &tmV, &bars[0], tile_col, kt * TQ, b);
  }

  for (int qt = 0; qt < q_tiles; ++qt) {
    const uint32_t par = qt & 1;
    const int nq_valid = min(p.Nq - qt * TQ, TQ);
    const int nq16 = (nq_valid + 15) & ~15;
    if (threadIdx.x == 0) {
      if (qt > 0) {                          // previous dV/dK UMMAs have finished reading sQ / sdO / sP / sdS
        mbar_wait(&bars[3], par ^ 1u);
        tc_fence_after();
      }
      mbar_expect_tx(&bars[1], 2 * TILE_BYTES);
      tma_load_3d(sQ, &tmQ, &bars[1], tile_col, qt * TQ, b);
      tma_load_3d(sdO, &tmdO, &bars[1], tile_col, qt * TQ, b);
      if (qt == 0) mbar_wait(&bars[0], 0);
      mbar_wait(&bars[1], par);
      tc_fence_after();
      const uint32_t idesc = umma_idesc_bf16(128, nk16, 0, 0);
#pragma unroll
      for (int j = 0; j < HD / 16; ++j)      // S = Q K^T
        tc_mma_f16_ss(tmem, umma_smem_desc_sw128(smem_u32(sQ) + sub_off + j * 32, 16, 1024),
                      umma_smem_desc_sw128(smem_u32(sK) + sub_off + j * 32, 16, 1024), idesc, j != 0 ? 1u : 0u);
#pragma unroll
      for (int j = 0; j < HD / 16; ++j)      // dP = dO V^T
        tc_mma_f16_ss(tmem + 128, umma_smem_desc_sw128(smem_u32(sdO) + sub_off + j * 32, 16, 1024),
                      umma_smem_desc_sw128(smem_u32(sV) + sub_off + j * 32, 16, 1024), idesc, j != 0 ? 1u : 0u);
      tc_commit(&bars[2]);
    }
    __syncwarp();
    if (qt > 0) {                            // every thread: P / dS buffers are free again
      mbar_wait(&bars[3], par ^ 1u);
      tc_fence_after();
    }
    mbar_wait(&bars[2], par);
    tc_fence_after();
    const int qrow = qt * TQ + row;
    const bool row_ok = qrow < p.Nq;
    const float lse2 = row_ok ? p.lse[(int64_t(b) * p.H + h) * p.Nq + qrow] * LOG2E_F : INFINITY;
    const float del = row_ok ? p.delta[(int64_t(b) * p.H + h) * p.Nq + qrow] : 0.f;
    bwd_tile_elementwise(trow, row, nk_valid, nk16, sl2, lse2, del, sP, sdS);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t idesc_t = umma_idesc_bf16(128, 64, 1, 1);
      for (int j = 0; j < nq16 / 16; ++j) {
        const uint32_t acc = (qt | j) != 0 ? 1u : 0u;
        tc_mma_f16_ss(tmem + 256, umma_smem_desc_sw128(smem_u32(sP) + j * 2048, TILE_BYTES, 1024),
                      umma_smem_desc_sw128(smem_u32(sdO) + j * 2048, TILE_BYTES, 1024), idesc_t, acc);
        tc_mma_f16_ss(tmem + 320, umma_smem_desc_sw128(smem_u32(sdS) + j * 2048, TILE_BYTES, 1024),
                      umma_smem_desc_sw128(smem_u32(sQ) + j * 2048, TILE_BYTES, 1024), idesc_t, acc);
      }
      tc_commit(&bars[3]);
    }
    __syncwarp();
  }
  mbar_wait(&bars[3], (q_tiles - 1) & 1);
  tc_fence_after();
  const int krow = kt * TQ + row;
  const int ocol0 = HD == 64 ? 0 : (h & 1) * 32;
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    bf16* base = which == 0 ? p.dV : p.dK;
    const int64_t ld = which == 0 ? p.lddv : p.lddk;
    const float mul = which == 0 ? 1.0f : p.scale;
#pragma unroll
    for (int c0 = 0; c0 < HD; c0 += 32) {
      uint32_t r[32];
      tmem_ld_32x32(trow + 256 + which * 64 + ocol0 + c0, r);
      tc_wait_ld();
      if (krow < p.Nk) {
        bf16* dst = base + (int64_t(b) * p.Nk + krow) * ld + h * HD + c0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(r[8 * j + 0]) * mul, __uint_as_float(r[8 * j + 1]) * mul);
          v.y = pack_bf16x2(__uint_as_float(r[8 * j + 2]) * mul, __uint_as_float(r[8 * j + 3]) * mul);
          v.z = pack_bf16x2(__uint_as_float(r[8 * j + 4]) * mul, __uint_as_float(r[8 * j + 5]) * mul);
          v.w = pack_bf16x2(__uint_as_float(r[8 * j + 6]) * mul, __uint_as_float(r[8 * j + 7]) * mul);
          *reinterpret_cast<uint4*>(dst + 8 * j) = v;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int HD>
__global__ void __launch_bounds__(128) attn_tc_bwd_dq_gen_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                 const __grid_constant__ CUtensorMap tmK,
                                                                 const __grid_constant__ CUtensorMap tmV,
                                                                 const __grid_constant__ CUtensorMap tmdO,
                                                                 const AttnTcBwdGenParams p) {
  pdl_prologue();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sdO = smem + 1 * TILE_BYTES;
  uint8_t* sK = smem + 2 * TILE_BYTES;
  uint8_t* sV = smem + 3 * TILE_BYTES;
  uint8_t* sdS = smem + 4 * TILE_BYTES;    // 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * TILE_BYTES);   // [0] Q,dO  [1] K,V  [2] S,dP  [3] dQ step
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 4);

  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_col = HD == 64 ? h * 64 : (h >> 1) * 64;
  const int sub_off = HD == 64 ? 0 : (h & 1) * 64;
  const int k_tiles = (p.Nk + TQ - 1) / TQ;

  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      tma_prefetch_desc(&tmdO);
#pragma unroll
      for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_smem;
  const int row = warp * 32 + lane;
  const int qrow = qt * TQ + row;
  const bool row_ok = qrow < p.Nq;
  const uint32_t trow = tmem + (uint32_t(warp * 32) << 16);
  const float sl2 = p.scale * LOG2E_F;
  const float lse2 = row_ok ? p.lse[(int64_t(b) * p.H + h) * p.Nq + qrow] * LOG2E_F : INFINITY;
  const float del = row_ok ? p.delta[(int64_t(b) * p.H + h) * p.Nq + qrow] : 0.f;

  if (threadIdx.x == 0) {
    mbar_expect_tx(&bars[0], 2 * TILE_BYTES);
    tma_load_3d(sQ, &tmQ, &bars[0], tile_col, qt * TQ, b);
    tma_load_3d(sdO, &tmdO, &bars[0], tile_col, qt * TQ, b);
  }
  for (int kt = 0; kt < k_tiles; ++kt) {
    const uint32_t par = kt & 1;
    const int nk_valid = min(p.Nk - kt * TQ, TQ);
    const int nk16 = (nk_valid + 15) & ~15;
    if (threadIdx.x == 0) {
      if (kt > 0) {
        mbar_wait(&bars[3], par ^ 1u);
        tc_fence_after();
      }
      mbar_expect_tx(&bars[1], 2 * TILE_BYTES);
      tma_load_3d(sK, &tmK, &bars[1], tile_col, kt * TQ, b);
      tma_load_3d(sV, &tmV, &bars[1], tile_col, kt * TQ, b);
      if (kt == 0) mbar_wait(&bars[0], 0);
      mbar_wait(&bars[1], par);
      tc_fence_after();
      const uint32_t idesc = umma_idesc_bf16(128, nk16, 0, 0);
#pragma unroll
      for (int j = 0; j < HD / 16; ++j)
        tc_mma_f16_ss(tmem, umma_smem_desc_sw128(smem_u32(sQ) + sub_off + j * 32, 16, 1024),
                      umma_smem_desc_sw128(smem_u32(sK) + sub_off + j * 32, 16, 1024), idesc, j != 0 ? 1u : 0u);
#pragma unroll
      for (int j = 0; j < HD / 16; ++j)
        tc_mma_f16_ss(tmem + 128, umma_smem_desc_sw128(smem_u32(sdO) + sub_off + j * 32, 16, 1024),
                      umma_smem_desc_sw128(smem_u32(sV) + sub_off + j * 32, 16, 1024), idesc, j != 0 ? 1u : 0u);
      tc_commit(&bars[2]);
    }
    __syncwarp();
    if (kt > 0) {
      mbar_wait(&bars[3], par ^ 1u);
      tc_fence_after();
    }
    mbar_wait(&bars[2], par);
    tc_fence_after();
    bwd_tile_elementwise(trow, row, nk_valid, nk16, sl2, lse2, del, nullptr, sdS);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t idesc_q = umma_idesc_bf16(128, 64, 0, 1);   // dQ += dS K : A = dS K-major, B = K tile MN-major
      for (int j = 0; j < nk16 / 16; ++j) {
        const uint32_t a_addr = smem_u32(sdS) + (j >> 2) * TILE_BYTES + (j & 3) * 32;
        tc_mma_f16_ss(tmem + 256, umma_smem_desc_sw128(a_addr, 16, 1024),
                      umma_smem_desc_sw128(smem_u32(sK) + j * 2048, TILE_BYTES, 1024), idesc_q, (kt | j) != 0 ? 1u : 0u);
      }
      tc_commit(&bars[3]);
    }
    __syncwarp();
  }
  mbar_wait(&bars[3], (k_tiles - 1) & 1);
  tc_fence_after();
  const int ocol0 = HD == 64 ? 0 : (h & 1) * 32;
#pragma unroll
  for (int c0 = 0; c0 < HD; c0 += 32) {
    uint32_t r[32];
    tmem_ld_32x32(trow + 256 + ocol0 + c0, r);
    tc_wait_ld();
    if (row_ok) {
      bf16* dst = p.dQ + (int64_t(b) * p.Nq + qrow) * p.lddq + h * HD + c0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 v;
        v.x = pack_bf16x2(__uint_as_float(r[8 * j + 0]) * p.scale, __uint_as_float(r[8 * j + 1]) * p.scale);
        v.y = pack_bf16x2(__uint_as_float(r[8 * j + 2]) * p.scale, __uint_as_float(r[8 * j + 3]) * p.scale);
        v.z = pack_bf16x2(__uint_as_float(r[8 * j + 4]) * p.scale, __uint_as_float(r[8 * j + 5]) * p.scale);
        v.w = pack_bf16x2(__uint_as_float(r[8 * j + 6]) * p.scale, __uint_as_float(r[8 * j + 7]) * p.scale);
        *reinterpret_cast<uint4*>(dst + 8 * j) = v;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int HD>
int launch_bwd_gen(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& tdo,
                   const AttnTcBwdGenParams& p, int B, cudaStream_t st) {
  constexpr int SMEM_KV = 8 * TILE_BYTES + 64 + 1024, SMEM_Q = 6 * TILE_BYTES + 64 + 1024;
  auto kkv = attn_tc_bwd_dkv_gen_kernel<HD>;
  auto kq = attn_tc_bwd_dq_gen_kernel<HD>;
  static bool configured = false;
  if (!configured) {
    MMAE_CUDA_OK(cudaFuncSetAttribute(kkv, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_KV));
    MMAE_CUDA_OK(cudaFuncSetAttribute(kq, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_Q));
    configured = true;
  }
  launch_k(kkv, dim3(ceil_div(p.Nk, TQ), p.H, B), 128, SMEM_KV, st, tq, tk, tv, tdo, p);
  count_launch();
  MMAE_LAUNCH_OK();
  launch_k(kq, dim3(ceil_div(p.Nq, TQ), p.H, B), 128, SMEM_Q, st, tq, tk, tv, tdo, p);
  count_launch();
  MMAE_LAUNCH_OK();
  return MMAE_OK;
}

int make_maps(CUtensorMap* out, const void* ptr, int64_t ld, int B, int N, int width_cols) {
  // [B][N][width] view with row pitch ld; box = 64 columns x 128 rows x 1 sample; rows past N are zero-filled
  return make_tmap_3d_bf16(out, ptr, (uint64_t)width_cols, (uint64_t)N, (uint64_t)B, (uint64_t)ld, (uint64_t)N * ld, DH, TQ, 1);
}

}  // namespace

// fused single-CTA kernels (forward + backward): the whole (b, h) problem in one tile
bool attn_tc_supported(int Nq, int Nk, int head_dim) { return head_dim == DH && Nq <= TQ && Nk <= TQ && Nq >= 1 && Nk >= 1; }
// general forward kernel: query tiles, up to 256 keys, head_dim 64 or 32 (even head count for 32)
bool attn_tc_fwd_gen_supported(int H, int Nq, int Nk, int head_dim) {
  return Nq >= 1 && Nk >= 1 && Nk <= 256 && (head_dim == 64 || (head_dim == 32 && H % 2 == 0));
}

bool attn_tc_bwd_gen_supported(int H, int Nq, int Nk, int head_dim) {
  return Nq >= 1 && Nk >= 1 && (head_dim == 64 || (head_dim == 32 && H % 2 == 0));
}

int attn_tc_backward_gen(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* d_o,
                         int64_t lddo, const float* lse, const float* delta, void* dq, int64_t lddq, void* dk, int64_t lddk,
                         void* dv, int64_t lddv, int B, int H, int Nq, int Nk, int head_dim, float scale, cudaStream_t st) {
  CUtensorMap tq, tk, tv, tdo;
  int rc;
  const int width = H * head_dim;
  if ((rc = make_maps(&tq, q, ldq, B, Nq, width)) || (rc = make_maps(&tk, k, ldk, B, Nk, width)) ||
      (rc = make_maps(&tv, v, ldv, B, Nk, width)) || (rc = make_maps(&tdo, d_o, lddo, B, Nq, width)))
    return rc;
  AttnTcBwdGenParams p;
  p.Nq = Nq; p.Nk = Nk; p.H = H; p.scale = scale;
  p.lse = lse; p.delta = delta;
  p.dQ = reinterpret_cast<bf16*>(dq); p.dK = reinterpret_cast<bf16*>(dk); p.dV = reinterpret_cast<bf16*>(dv);
  p.lddq = lddq; p.lddk = lddk; p.lddv = lddv;
  return head_dim == 64 ? launch_bwd_gen<64>(tq, tk, tv, tdo, p, B, st) : launch_bwd_gen<32>(tq, tk, tv, tdo, p, B, st);
}

int attn_tc_forward_gen(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o,
                        int64_t ldo, float* lse, int B, int H, int Nq, int Nk, int head_dim, float scale, cudaStream_t st) {
  CUtensorMap tq, tk, tv;
  int rc;
  const int width = H * head_dim;
  if ((rc = make_maps(&tq, q, ldq, B, Nq, width)) || (rc = make_maps(&tk, k, ldk, B, Nk, width)) ||
      (rc = make_maps(&tv, v, ldv, B, Nk, width)))
    return rc;
  AttnTcParams p;
  p.Nq = Nq; p.Nk = Nk; p.H = H; p.scale = scale;
  p.O = reinterpret_cast<bf16*>(o);
  p.ldo = ldo;
  p.lse = lse;
  const bool two = Nk > 128;
  if (head_dim == 64) return two ? launch_fwd_gen<64, 2>(tq, tk, tv, p, B, st) : launch_fwd_gen<64, 1>(tq, tk, tv, p, B, st);
  return two ? launch_fwd_gen<32, 2>(tq, tk, tv, p, B, st) : launch_fwd_gen<32, 1>(tq, tk, tv, p, B, st);
}

int attn_tc_forward(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o, int64_t ldo,
                    float* lse, int B, int H, int Nq, int Nk, float scale, cudaStream_t st) {
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_maps(&tq, q, ldq, B, Nq, H * DH)) || (rc = make_maps(&tk, k, ldk, B, Nk, H * DH)) ||
      (rc = make_maps(&tv, v, ldv, B, Nk, H * DH)))
    return rc;
  AttnTcParams p;
  p.Nq = Nq; p.Nk = Nk; p.H = H; p.scale = scale;
  p.O = reinterpret_cast<bf16*>(o);
  p.ldo = ldo;
  p.lse = lse;
  constexpr int SMEM = 3 * TILE_BYTES + 64 + 1024;
  static bool configured = false;
  if (!configured) {
    MMAE_CUDA_OK(cudaFuncSetAttribute(attn_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  launch_k(attn_tc_fwd_kernel, dim3(H, B), 128, SMEM, st, tq, tk, tv, p);
  count_launch();
  MMAE_LAUNCH_OK();
  return MMAE_OK;
}

// experimental persistent forward (bit 4 of mmae_attention_set_tc), same contract as attn_tc_forward
int attn_tc_forward_persistent(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o,
                               int64_t ldo, float* lse, int B, int H, int Nq, int Nk, float scale, cudaStream_t st) {
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_maps(&tq, q, ldq, B, Nq, H * DH)) || (rc = make_maps(&tk, k, ldk, B, Nk, H * DH)) ||
      (rc = make_maps(&tv, v, ldv, B, Nk, H * DH)))
    return rc;
  AttnTcParams p;
  p.Nq = Nq; p.Nk = Nk; p.H = H; p.scale = scale;
  p.O = reinterpret_cast<bf16*>(o);
  p.ldo = ldo;
  p.lse = lse;
  static bool configured = false;
  if (!configured) {
    MMAE_CUDA_OK(cudaFuncSetAttribute(attn_tc_fwd_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD2_SMEM));
    configured = true;
  }
  const int items = B * H;
  const int grid = std::min(items, 2 * sm_count());
  launch_k(attn_tc_fwd_persistent_kernel, grid, 128, FWD2_SMEM, st, tq, tk, tv, p, items);
  count_launch();
  MMAE_LAUNCH_OK();
  return MMAE_OK;
}

int attn_tc_backward(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* d_o,
                     int64_t lddo, const float* lse, const float* delta, void* dq, int64_t lddq, void* dk, int64_t lddk,
                     void* dv, int64_t lddv, int B, int H, int Nq, int Nk, float scale, cudaStream_t st) {
  CUtensorMap tq, tk, tv, tdo;
  int rc;
  if ((rc = make_maps(&tq, q, ldq, B, Nq, H * DH)) || (rc = make_maps(&tk, k, ldk, B, Nk, H * DH)) ||
      (rc = make_maps(&tv, v, ldv, B, Nk, H * DH)) || (rc = make_maps(&tdo, d_o, lddo, B, Nq, H * DH)))
    return rc;
  AttnTcBwdParams p;
  p.Nq = Nq; p.Nk = Nk; p.H = H; p.scale = scale;
  p.lse = lse; p.delta = delta;
  p.dQ = reinterpret_cast<bf16*>(dq); p.dK = reinterpret_cast<bf16*>(dk); p.dV = reinterpret_cast<bf16*>(dv);
  p.lddq = lddq; p.lddk = lddk; p.lddv = lddv;
  constexpr int SMEM = ATTN_BWD_SMEM;
  static bool configured = false;
  if (!configured) {
    MMAE_CUDA_OK(cudaFuncSetAttribute(attn_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    MMAE_CUDA_OK(cudaFuncSetAttribute(attn_tc_bwd_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    configured = true;
  }
  launch_k(attn_tc_bwd_kernel, dim3(H, B), 128, SMEM, st, tq, tk, tv, tdo, p);
  count_launch();
  MMAE_LAUNCH_OK();
  return MMAE_OK;
}

}  // namespace mmae

// Warp-specialised, persistent tcgen05 attention for the MultiMAE-B/L shapes (multimae/multimae_utils.py:170-182 encoder
// self-attention, :199-214 decoder cross-attention and the decoder blocks' self-attention), forward and backward.
//
// One CTA per SM, 10 warps:
//   warp 0      TMA producer   : Q / K / V (/ dO) tiles of the next work items into a 2..3-stage shared-memory ring
//   warp 1      MMA issuer     : one elected thread issues every tcgen05.mma and the commits that drive the barriers
//   warps 2..9  compute        : 256 threads = TWO threads per query row (the TMEM lane), each owning half of the score
//                                columns: softmax / P / dS arithmetic, TMEM <-> registers, bf16 results to global memory
// The three roles only meet at mbarriers, so the loads of item i+2, the S = Q K^T (and dP = dO V^T) MMAs of item i+1, and
// the exp / pack / store work of item i overlap.  The one-CTA-per-(b, h) kernels in attention_tc.cu run the same steps
// back to back in four warps with one warp per scheduler: every tcgen05.ld / mbarrier / MUFU latency is exposed (ncu:
// 60-80 % "no eligible warp", 18-40 % issue-slot utilisation at 25 us forward / 75 us backward per encoder layer).
//
// forward  work item = (b, h, 128-query tile):  S[128 x nk16 <= 256] in TMEM (two buffers: S of item i+1 is computed while
//          item i is in its softmax) -> P bf16 in 128B-swizzled smem panels -> O = P V accumulates over the dead S columns.
// backward work item = (b, h); block = (128-key tile, 128-query tile), key tile outer.  Per block S and dP = dO V^T are
//          recomputed, P / dS are written ONCE to smem and consumed by three UMMAs in two operand roles (MN-major A for
//          dV += P^T dO and dK += dS^T Q, K-major A for dQ += dS K).  dV / dK / dQ accumulate in TMEM across the blocks of
//          an item (UMMA accumulate flag) - the general two-kernel tcgen05 backward of attention_tc.cu recomputed S / dP in
//          both kernels and allocated all 512 TMEM columns per 128-thread CTA (one CTA per SM, 264 us at 196 x 196 x 32).
// head_dim 32: shared-memory tiles stay 64 columns (one 128-byte swizzle row) and hold a PAIR of heads; S / dP use the head's
//          32-column half as K extent (descriptor start + 64 B inside the swizzled row); the N = 64 accumulators carry the
//          sibling head's columns as don't-care values that the epilogue skips.
#include "common.cuh"
#include "../../include/multimae_b200.h"

namespace mmae {
void count_launch();
namespace {

constexpr int TQ = 128;                    // rows per tile (one TMEM lane each)
constexpr int TILE_BYTES = TQ * 64 * 2;    // 16 KB: [128 x 64] bf16, 128-byte rows
constexpr float LOG2E_F = 1.4426950408889634f;
constexpr int WS_THREADS = 320;            // warp 0 TMA, warp 1 MMA, warps 2..9 compute
constexpr int NCOMPUTE = 256;

__device__ __forceinline__ uint32_t swz128(int r, int j) { return uint32_t(r) * 128u + uint32_t((j ^ (r & 7)) << 4); }

// 32 consecutive bf16 (packed 2 per u32) of row r, columns [c0, c0 + 32) of a K-major buffer made of 64-column panels
__device__ __forceinline__ void store_row32(uint8_t* buf, int r, int c0, const uint32_t (&pk)[16]) {
  uint8_t* panel = buf + (c0 >> 6) * TILE_BYTES;
  const int j0 = (c0 & 63) >> 3;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    *reinterpret_cast<uint4*>(panel + swz128(r, j0 + j)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
}

// phase time stamps of CTA 0 (first 64 items / blocks): slot = 16 * it + k
__device__ __forceinline__ void trace_stamp(long long* trace, int it, int k) {
  if (trace != nullptr && blockIdx.x == 0 && it < 64) trace[16 * it + k] = clock64();
}

__device__ __forceinline__ void compute_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ uint4 pack8(const uint32_t* r, float mul) {
  uint4 v;
  v.x = pack_bf16x2(__uint_as_float(r[0]) * mul, __uint_as_float(r[1]) * mul);
  v.y = pack_bf16x2(__uint_as_float(r[2]) * mul, __uint_as_float(r[3]) * mul);
  v.z = pack_bf16x2(__uint_as_float(r[4]) * mul, __uint_as_float(r[5]) * mul);
  v.w = pack_bf16x2(__uint_as_float(r[6]) * mul, __uint_as_float(r[7]) * mul);
  return v;
}

// This thread's share of a [128 x 64] accumulator row -> bf16 -> shared-memory staging tile in the layout of the TMA store
// map.  Per-lane global stores of 16-byte pieces touch 32 different 128-byte lines per warp instruction (one L1 tag cycle
// each): they cost 1000 (forward) to 6000 (backward: dV, dK, dQ) clocks per work item, more than everything else together
// (measured with the clock64 trace); the TMA unit writes whole rows and clips the rows past the sequence end itself.
//   HD 64: 128-byte rows, SWIZZLE_128B; thread (row, hf) owns the 16-byte chunks 4 hf .. 4 hf + 3
//   HD 32: 64-byte rows (the head's 32 columns), SWIZZLE_64B; thread (row, hf) owns chunks 2 hf, 2 hf + 1
template <int HD>
__device__ __forceinline__ void acc_to_stage(uint32_t taddr_row, int row, int hf, int hsel, uint8_t* stage, float mul) {
  if constexpr (HD == 64) {
    uint32_t r[32];
    tmem_ld_32x32(taddr_row + uint32_t(hf * 32), r);
    tc_wait_ld();
#pragma unroll
    for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(stage + swz128(row, hf * 4 + j)) = pack8(r + 8 * j, mul);
  } else {
    uint32_t r[16];
    tmem_ld_32x16(taddr_row + uint32_t(hsel * 32 + hf * 16), r);
    tc_wait_ld();
    const int sw = (row >> 1) & 3;
#pragma unroll
    for (int j = 0; j < 2; ++j)
      *reinterpret_cast<uint4*>(stage + row * 64 + (((hf * 2 + j) ^ sw) << 4)) = pack8(r + 8 * j, mul);
  }
}

// =====================================================================================================================
// forward
// =====================================================================================================================
struct WsFwdParams {
  int B, H, Nq, Nk, q_tiles, items;
  float scale;
  bf16* O;
  int64_t ldo;
  float* lse;
  long long* trace;   // diagnostics (mmae_attention_ws_set_trace): clock64 stamps of CTA 0's phases, 16 slots per item
};

template <int KBOX>
struct WsFwdCfg {
  static constexpr int STAGES = KBOX == 1 ? 3 : 2;
  static constexpr int STAGE_BYTES = (1 + 2 * KBOX) * TILE_BYTES;       // Q | K boxes | V boxes
  static constexpr int P_BYTES = 2 * KBOX * TILE_BYTES;                 // 64-key panels
  static constexpr int P_OFF = STAGES * STAGE_BYTES;
  static constexpr int X_OFF = P_OFF + P_BYTES;                         // row max / row sum exchange: 2 x 2 x 128 floats
  static constexpr int BAR_OFF = X_OFF + 2048;
  static constexpr int SLACK = KBOX == 1 ? 1024 : 768;                  // alignment slack (the kernel traps if it needs more)
  static constexpr int SMEM = BAR_OFF + 256 + SLACK;                    // + barriers + alignment slack
  static constexpr int SCOLS = 128 * KBOX;                              // TMEM columns of one S buffer
  static_assert(SMEM <= 232448, "forward exceeds 227 KB of shared memory");
};

template <int HD, int KBOX>
__global__ void __launch_bounds__(WS_THREADS, 1) attn_ws_fwd_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                    const __grid_constant__ CUtensorMap tmK,
                                                                    const __grid_constant__ CUtensorMap tmV,
                                                                    const __grid_constant__ CUtensorMap tmO,
                                                                    const WsFwdParams p) {
  pdl_launch_dependents();
  using C = WsFwdCfg<KBOX>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  if (pad > uint32_t(C::SLACK)) __trap();
  uint8_t* smem = smem_raw + pad;
  uint8_t* sP = smem + C::P_OFF;
  float* xmax = reinterpret_cast<float*>(smem + C::X_OFF);        // [2][128]
  float* xsum = xmax + 256;                                       // [2][128]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* s_full = empty + STAGES;      // [2]
  uint64_t* s_free = s_full + 2;          // [2]
  uint64_t* o_full = s_free + 2;          // [2]
  uint64_t* p_full = o_full + 2;          // [1]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(p_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk16 = (p.Nk + 15) & ~15;

  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      tma_prefetch_desc(&tmO);
    }
  } else if (warp == 1) {
    if (elect_one()) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&empty[s], 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(&s_full[b], 1);
        mbar_init(&s_free[b], NCOMPUTE / 32);
        mbar_init(&o_full[b], 1);
      }
      mbar_init(p_full, NCOMPUTE / 32);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr_smem, 2 * C::SCOLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_smem;
  pdl_wait();

  auto decode = [&](int item, int& b, int& h, int& qt) {
    qt = item % p.q_tiles;
    const int bh = item / p.q_tiles;
    h = bh % p.H;
    b = bh / p.H;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int it = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
        int b, h, qt;
        decode(item, b, h, qt);
        const int s = it % STAGES;
        const uint32_t k = uint32_t(it / STAGES);
        mbar_wait(&empty[s], (k & 1u) ^ 1u);
        trace_stamp(p.trace, it, 0);                        // loads issued
        const int tile_col = HD == 64 ? h * 64 : (h >> 1) * 64;
        uint8_t* st = smem + s * C::STAGE_BYTES;
        mbar_expect_tx(&full[s], C::STAGE_BYTES);
        tma_load_3d(st, &tmQ, &full[s], tile_col, qt * TQ, b);
#pragma unroll
        for (int kb = 0; kb < KBOX; ++kb) {
          tma_load_3d(st + (1 + kb) * TILE_BYTES, &tmK, &full[s], tile_col, kb * TQ, b);
          tma_load_3d(st + (1 + KBOX + kb) * TILE_BYTES, &tmV, &full[s], tile_col, kb * TQ, b);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint32_t idesc_s = umma_idesc_bf16(128, nk16, 0, 0);
      const uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);
      auto issue_s = [&](int it, int item) {
        int b, h, qt;
        decode(item, b, h, qt);
        const int s = it % STAGES, buf = it & 1;
        const uint32_t k = uint32_t(it / STAGES), u = uint32_t(it >> 1);
        const int sub_off = HD == 64 ? 0 : (h & 1) * 64;
        mbar_wait(&full[s], k & 1u);
        trace_stamp(p.trace, it, 1);                        // loads landed
        mbar_wait(&s_free[buf], (u & 1u) ^ 1u);
        tc_fence_after();
        trace_stamp(p.trace, it, 2);                        // S issued
        const uint32_t q_addr = smem_u32(smem + s * C::STAGE_BYTES), k_addr = q_addr + TILE_BYTES;
#pragma unroll
        for (int j = 0; j < HD / 16; ++j)
          tc_mma_f16_ss(tmem + uint32_t(buf * C::SCOLS), umma_smem_desc_sw128(q_addr + sub_off + j * 32, 16, 1024),
                        umma_smem_desc_sw128(k_addr + sub_off + j * 32, 16, 1024), idesc_s, j != 0 ? 1u : 0u);
        tc_commit(&s_full[buf]);
      };
      int it = 0;
      if (int(blockIdx.x) < p.items) issue_s(0, blockIdx.x);
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
        const int next = item + int(gridDim.x);
        if (next < p.items) issue_s(it + 1, next);          // S of the next item runs under this item's softmax
        const int s = it % STAGES, buf = it & 1;
        mbar_wait(p_full, uint32_t(it) & 1u);
        tc_fence_after();
        trace_stamp(p.trace, it, 3);                        // P V issued
        const uint32_t v_addr = smem_u32(smem + s * C::STAGE_BYTES) + (1 + KBOX) * TILE_BYTES;
        const int ksteps = nk16 / 16;
        for (int j = 0; j < ksteps; ++j) {
          // A = P (K-major, 64-key panels), B = V (MN-major: the row-major [key][dh] tile as it is; 128-key boxes)
          const uint32_t a_addr = smem_u32(sP) + (j >> 2) * TILE_BYTES + (j & 3) * 32;
          const uint32_t b_addr = v_addr + (j >> 3) * TILE_BYTES + (j & 7) * 2048;
          tc_mma_f16_ss(tmem + uint32_t(buf * C::SCOLS), umma_smem_desc_sw128(a_addr, 16, 1024),
                        umma_smem_desc_sw128(b_addr, TILE_BYTES, 1024), idesc_o, j != 0 ? 1u : 0u);
        }
        tc_commit(&o_full[buf]);
        tc_commit(&empty[s]);
      }
    }
  } else {
    // ------------------------------------------------------------------ compute: two threads per query row
    const int cw = warp - 2;                       // 0..7
    const int hf = cw >> 2;                        // which half of the score columns
    const int row = (warp & 3) * 32 + lane;        // TMEM lane quarter of a warp = warp id % 4
    const uint32_t lane_addr = uint32_t((warp & 3) * 32) << 16;
    constexpr int HALF = 64 * KBOX;                // score columns per thread
    const int cbeg = hf * HALF;
    const bool is_issuer = cw == 0 && lane == 0;   // the compute thread that issues the TMA stores
    const float sl2 = p.scale * LOG2E_F;
    int it = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
      int b, h, qt;
      decode(item, b, h, qt);
      const int buf = it & 1;
      const uint32_t u = uint32_t(it >> 1);
      const uint32_t ts = tmem + lane_addr + uint32_t(buf * C::SCOLS);
      if (cw == 0 && lane == 0) trace_stamp(p.trace, it, 4);   // compute: waiting for S
      mbar_wait(&s_full[buf], u & 1u);
      tc_fence_after();
      if (cw == 0 && lane == 0) trace_stamp(p.trace, it, 5);   // S ready
      // ---- pass 1: row maximum over this thread's columns (<= 128 keys: the scores stay in registers for pass 2)
      constexpr bool KEEP = KBOX == 1;
      uint32_t sreg[KEEP ? HALF / 32 : 1][32];
      float mx = -INFINITY;
#pragma unroll
      for (int c0 = 0; c0 < HALF; c0 += 32) {
        if (cbeg + c0 < nk16) {
          uint32_t rl[32];
          uint32_t(&r)[32] = KEEP ? sreg[KEEP ? c0 / 32 : 0] : rl;
          tmem_ld_32x32(ts + uint32_t(cbeg + c0), r);
          tc_wait_ld();
          if (cbeg + c0 + 32 <= p.Nk) {
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (cbeg + c0 + i < p.Nk) mx = fmaxf(mx, __uint_as_float(r[i]));
          }
        }
      }
      // the previous item's TMA store has read its staging tile (which aliases P panel 0) before any thread writes P
      if (is_issuer) bulk_wait_read_all();
      xmax[hf * 128 + row] = mx;
      compute_bar();
      mx = fmaxf(xmax[row], xmax[128 + row]);
      if (cw == 0 && lane == 0) trace_stamp(p.trace, it, 6);   // max known
      const float moff = mx * sl2;
      // ---- pass 2: P = exp2(S * scale * log2e - max), row sum, bf16 panels (the previous item's P V has completed:
      //      every compute thread waited for its o_full before leaving the previous iteration)
      float sum = 0.f;
#pragma unroll
      for (int c0 = 0; c0 < HALF; c0 += 32) {
        uint32_t pk[16];
        if (cbeg + c0 < nk16) {
          uint32_t rl[32];
          uint32_t(&r)[32] = KEEP ? sreg[KEEP ? c0 / 32 : 0] : rl;
          if constexpr (!KEEP) {
            tmem_ld_32x32(ts + uint32_t(cbeg + c0), rl);
            tc_wait_ld();
          }
          if (cbeg + c0 + 32 <= p.Nk) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float a = fast_exp2(fmaf(__uint_as_float(r[2 * i]), sl2, -moff));
              const float c = fast_exp2(fmaf(__uint_as_float(r[2 * i + 1]), sl2, -moff));
              sum += a + c;
              pk[i] = pack_bf16x2(a, c);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int k0 = cbeg + c0 + 2 * i;
              const float a = k0 < p.Nk ? fast_exp2(fmaf(__uint_as_float(r[2 * i]), sl2, -moff)) : 0.f;
              const float c = k0 + 1 < p.Nk ? fast_exp2(fmaf(__uint_as_float(r[2 * i + 1]), sl2, -moff)) : 0.f;
              sum += a + c;
              pk[i] = pack_bf16x2(a, c);
            }
          }
          store_row32(sP, row, cbeg + c0, pk);
        }
      }
      xsum[hf * 128 + row] = sum;
      fence_proxy_async_smem();          // generic-proxy smem writes -> visible to the tensor-core (async) proxy
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (cw == 0 && lane == 0) trace_stamp(p.trace, it, 7);   // P written
      // ---- epilogue: O / l -> bf16 -> staging tile (P panel 0: P V has finished reading P) -> one TMA store
      mbar_wait(&o_full[buf], u & 1u);
      tc_fence_after();
      if (cw == 0 && lane == 0) trace_stamp(p.trace, it, 8);   // O ready
      const float l = xsum[row] + xsum[128 + row];
      const float inv = 1.0f / l;
      const int qrow = qt * TQ + row;
      acc_to_stage<HD>(ts, row, hf, h & 1, sP, inv);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[buf]);           // this S / O buffer may be overwritten
      if (hf == 0 && p.lse != nullptr && qrow < p.Nq) p.lse[(int64_t(b) * p.H + h) * p.Nq + qrow] = mx * p.scale + logf(l);
      fence_proxy_async_smem();
      compute_bar();
      if (is_issuer) {
        tma_store_3d(&tmO, sP, h * HD, qt * TQ, b);       // rows past Nq are clipped by the TMA unit
        bulk_commit_group();
      }
      if (cw == 0 && lane == 0) trace_stamp(p.trace, it, 9);   // epilogue done
    }
    if (is_issuer) bulk_wait_all();     // the last store has left shared memory before the CTA does
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 2 * C::SCOLS);
  }
}

// =====================================================================================================================
// backward
// =====================================================================================================================
struct WsBwdParams {
  int B, H, Nq, Nk, q_tiles, k_tiles, items;
  float scale;
  const float* lse;
  const float* delta;
  bf16 *dQ, *dK, *dV;
  int64_t lddq, lddk, lddv;
  long long* trace;
};

struct WsBwdCfg {
  static constexpr int STAGES = 2;
  static constexpr int STAGE_BYTES = 5 * TILE_BYTES;                   // Q | dO | K | V | O
  static constexpr int P_OFF = STAGES * STAGE_BYTES;                    // P : two 64-key panels
  static constexpr int DS_OFF = P_OFF + 2 * TILE_BYTES;                 // dS: two 64-key panels
  static constexpr int X_OFF = DS_OFF + 2 * TILE_BYTES;                 // delta exchange: 2 x 128 floats
  static constexpr int BAR_OFF = X_OFF + 1024;
  static constexpr int SMEM = BAR_OFF + 256 + 1024;
  static_assert(SMEM <= 232448, "backward exceeds 227 KB of shared memory");
  // TMEM columns: S [0,128)  dP [128,256)  dV [256,320)  dK [320,384)  dQ of query tile t [384 + 64 t, +64), t < 2
  static constexpr uint32_t COL_DP = 128, COL_DV = 256, COL_DK = 320, COL_DQ = 384;
};

template <int HD>
__global__ void __launch_bounds__(WS_THREADS, 1) attn_ws_bwd_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                    const __grid_constant__ CUtensorMap tmK,
                                                                    const __grid_constant__ CUtensorMap tmV,
                                                                    const __grid_constant__ CUtensorMap tmdO,
                                                                    const __grid_constant__ CUtensorMap tmO,
                                                                    const __grid_constant__ CUtensorMap tmdQ,
                                                                    const __grid_constant__ CUtensorMap tmdK,
                                                                    const __grid_constant__ CUtensorMap tmdV,
                                                                    const WsBwdParams p) {
  pdl_launch_dependents();
  using C = WsBwdCfg;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sP = smem + C::P_OFF;
  uint8_t* sdS = smem + C::DS_OFF;
  float* xdel = reinterpret_cast<float*>(smem + C::X_OFF);       // [2][128]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* sdp_full = empty + STAGES;
  uint64_t* pds_full = sdp_full + 1;
  uint64_t* grads_done = pds_full + 1;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(grads_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int blocks_per_item = p.q_tiles * p.k_tiles;

  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      tma_prefetch_desc(&tmdO);
      tma_prefetch_desc(&tmO);
      tma_prefetch_desc(&tmdQ);
      tma_prefetch_desc(&tmdK);
      tma_prefetch_desc(&tmdV);
    }
  } else if (warp == 1) {
    if (elect_one()) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&empty[s], 1);
      }
      mbar_init(sdp_full, 1);
      mbar_init(pds_full, NCOMPUTE / 32);
      mbar_init(grads_done, 1);
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
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer: one stage per block
    if (elect_one()) {
      int it = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const int h = item % p.H, b = item / p.H;
        const int tile_col = HD == 64 ? h * 64 : (h >> 1) * 64;
        for (int kt = 0; kt < p.k_tiles; ++kt)
          for (int qt = 0; qt < p.q_tiles; ++qt, ++it) {
            const int s = it % STAGES;
            const uint32_t k = uint32_t(it / STAGES);
            mbar_wait(&empty[s], (k & 1u) ^ 1u);
            trace_stamp(p.trace, it, 0);
            uint8_t* st = smem + s * C::STAGE_BYTES;
            mbar_expect_tx(&full[s], C::STAGE_BYTES);
            tma_load_3d(st, &tmQ, &full[s], tile_col, qt * TQ, b);
            tma_load_3d(st + TILE_BYTES, &tmdO, &full[s], tile_col, qt * TQ, b);
            tma_load_3d(st + 2 * TILE_BYTES, &tmK, &full[s], tile_col, kt * TQ, b);
            tma_load_3d(st + 3 * TILE_BYTES, &tmV, &full[s], tile_col, kt * TQ, b);
            tma_load_3d(st + 4 * TILE_BYTES, &tmO, &full[s], tile_col, qt * TQ, b);
          }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint32_t idesc_t = umma_idesc_bf16(128, 64, 1, 1);    // dV / dK: A (P / dS) MN-major, B (dO / Q) MN-major
      const uint32_t idesc_q = umma_idesc_bf16(128, 64, 0, 1);    // dQ     : A (dS) K-major,     B (K) MN-major
      int it = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const int h = item % p.H;
        const int sub_off = HD == 64 ? 0 : (h & 1) * 64;
        for (int kt = 0; kt < p.k_tiles; ++kt) {
          const int nk_valid = min(p.Nk - kt * TQ, TQ);
          const int nk16 = (nk_valid + 15) & ~15;
          for (int qt = 0; qt < p.q_tiles; ++qt, ++it) {
            const int nq_valid = min(p.Nq - qt * TQ, TQ);
            const int nq16 = (nq_valid + 15) & ~15;
            const int s = it % STAGES;
            const uint32_t k = uint32_t(it / STAGES);
            const uint32_t q_addr = smem_u32(smem + s * C::STAGE_BYTES), do_addr = q_addr + TILE_BYTES,
                           k_addr = q_addr + 2 * TILE_BYTES, v_addr = q_addr + 3 * TILE_BYTES;
            mbar_wait(&full[s], k & 1u);
            // S / dP of the previous block have been read (its pds_full was awaited below before its gradient MMAs)
            tc_fence_after();
            trace_stamp(p.trace, it, 1);
            const uint32_t idesc_s = umma_idesc_bf16(128, nk16, 0, 0);
#pragma unroll
            for (int j = 0; j < HD / 16; ++j)      // S = Q K^T
              tc_mma_f16_ss(tmem, umma_smem_desc_sw128(q_addr + sub_off + j * 32, 16, 1024),
                            umma_smem_desc_sw128(k_addr + sub_off + j * 32, 16, 1024), idesc_s, j != 0 ? 1u : 0u);
#pragma unroll
            for (int j = 0; j < HD / 16; ++j)      // dP = dO V^T
              tc_mma_f16_ss(tmem + C::COL_DP, umma_smem_desc_sw128(do_addr + sub_off + j * 32, 16, 1024),
                            umma_smem_desc_sw128(v_addr + sub_off + j * 32, 16, 1024), idesc_s, j != 0 ? 1u : 0u);
            tc_commit(sdp_full);
            mbar_wait(pds_full, uint32_t(it) & 1u);            // P / dS are in shared memory, S / dP have been read
            tc_fence_after();
            trace_stamp(p.trace, it, 3);
            for (int j = 0; j < nq16 / 16; ++j) {
              const uint32_t acc = (qt | j) != 0 ? 1u : 0u;     // first query tile of this key tile starts dV / dK
              tc_mma_f16_ss(tmem + C::COL_DV, umma_smem_desc_sw128(smem_u32(sP) + j * 2048, TILE_BYTES, 1024),
                            umma_smem_desc_sw128(do_addr + j * 2048, TILE_BYTES, 1024), idesc_t, acc);
              tc_mma_f16_ss(tmem + C::COL_DK, umma_smem_desc_sw128(smem_u32(sdS) + j * 2048, TILE_BYTES, 1024),
                            umma_smem_desc_sw128(q_addr + j * 2048, TILE_BYTES, 1024), idesc_t, acc);
            }
            for (int j = 0; j < nk16 / 16; ++j) {
              const uint32_t a_addr = smem_u32(sdS) + (j >> 2) * TILE_BYTES + (j & 3) * 32;
              tc_mma_f16_ss(tmem + C::COL_DQ + uint32_t(qt * 64), umma_smem_desc_sw128(a_addr, 16, 1024),
                            umma_smem_desc_sw128(k_addr + j * 2048, TILE_BYTES, 1024), idesc_q, (kt | j) != 0 ? 1u : 0u);
            }
            tc_commit(grads_done);
            tc_commit(&empty[s]);
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ compute: two threads per row
    const int cw = warp - 2;
    const int hf = cw >> 2;                         // key-column half of the block = P / dS panel
    const int row = (warp & 3) * 32 + lane;
    const uint32_t trow = tmem + (uint32_t((warp & 3) * 32) << 16);
    const float sl2 = p.scale * LOG2E_F;
    const bool is_issuer = cw == 0 && lane == 0;    // the compute thread that issues the TMA stores
    bool staged = false;                            // the previous block left gradient tiles in the P / dS buffers
    float lse_next = int(blockIdx.x) < p.items && row < p.Nq ? __ldg(p.lse + int64_t(blockIdx.x) * p.Nq + row) : 0.f;
    int it = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const int h = item % p.H, b = item / p.H;
      for (int kt = 0; kt < p.k_tiles; ++kt) {
        const int nk_valid = min(p.Nk - kt * TQ, TQ);
        const int nk16 = (nk_valid + 15) & ~15;
        for (int qt = 0; qt < p.q_tiles; ++qt, ++it) {
          const int qrow = qt * TQ + row;
          const bool row_ok = qrow < p.Nq;
          // log-sum-exp of this row: loaded one block ahead (below), so its DRAM latency hides behind the previous block
          const float lse2 = row_ok ? lse_next * LOG2E_F : INFINITY;   // +inf: P = 0
          {
            int qt_n = qt + 1, kt_n = kt, item_n = item;
            if (qt_n == p.q_tiles) {
              qt_n = 0;
              if (++kt_n == p.k_tiles) {
                kt_n = 0;
                item_n += int(gridDim.x);
              }
            }
            const int qrow_n = qt_n * TQ + row;
            lse_next = item_n < p.items && qrow_n < p.Nq ? __ldg(p.lse + int64_t(item_n) * p.Nq + qrow_n) : 0.f;
          }
          if (cw == 0 && lane == 0) trace_stamp(p.trace, it, 4);
          // ---- delta = rowsum(dO o O) from the staged dO / O tiles (each thread: its half of the head's columns), exchanged
          //      between the two threads of a row.  Replaces the separate delta kernel and its re-read of O and dO.
          const int s = it % STAGES;
          mbar_wait(&full[s], uint32_t(it / STAGES) & 1u);      // the TMA writes of this stage are visible to this thread
          {
            const uint8_t* sdO_t = smem + s * C::STAGE_BYTES + TILE_BYTES;
            const uint8_t* sO_t = smem + s * C::STAGE_BYTES + 4 * TILE_BYTES;
            constexpr int NCH = HD == 64 ? 4 : 2;                // 16-byte chunks per thread
            const int ch0 = HD == 64 ? hf * 4 : (h & 1) * 4 + hf * 2;
            float part = 0.f;
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
              const uint4 a = *reinterpret_cast<const uint4*>(sdO_t + swz128(row, ch0 + j));
              const uint4 o4 = *reinterpret_cast<const uint4*>(sO_t + swz128(row, ch0 + j));
              const uint32_t* au = reinterpret_cast<const uint32_t*>(&a);
              const uint32_t* ou = reinterpret_cast<const uint32_t*>(&o4);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 x = unpack_bf16x2(au[e]), y = unpack_bf16x2(ou[e]);
                part = fmaf(x.x, y.x, part);
                part = fmaf(x.y, y.y, part);
              }
            }
            xdel[hf * 128 + row] = part;
          }
          // P / dS buffers: the previous block's gradient MMAs have completed (grads_done awaited at the end of the
          // previous iteration by every compute thread); gradient tiles staged there have been read by their TMA stores
          if (staged && is_issuer) bulk_wait_read_all();
          compute_bar();
          const float del = xdel[row] + xdel[128 + row];
          mbar_wait(sdp_full, uint32_t(it) & 1u);
          tc_fence_after();
          if (cw == 0 && lane == 0) trace_stamp(p.trace, it, 5);
#pragma unroll
          for (int c0 = 0; c0 < 64; c0 += 32) {
            const int col = hf * 64 + c0;
            uint32_t pp[16], ds[16];
            if (col < nk16) {
              uint32_t sv[32], dv[32];
              tmem_ld_32x32(trow + uint32_t(col), sv);
              tmem_ld_32x32(trow + C::COL_DP + uint32_t(col), dv);
              tc_wait_ld();
              if (col + 32 <= nk_valid) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const float p0 = fast_exp2(fmaf(__uint_as_float(sv[2 * i]), sl2, -lse2));
                  const float p1 = fast_exp2(fmaf(__uint_as_float(sv[2 * i + 1]), sl2, -lse2));
                  pp[i] = pack_bf16x2(p0, p1);
                  ds[i] = pack_bf16x2(p0 * (__uint_as_float(dv[2 * i]) - del), p1 * (__uint_as_float(dv[2 * i + 1]) - del));
                }
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const float p0 = col + 2 * i < nk_valid ? fast_exp2(fmaf(__uint_as_float(sv[2 * i]), sl2, -lse2)) : 0.f;
                  const float p1 = col + 2 * i + 1 < nk_valid ? fast_exp2(fmaf(__uint_as_float(sv[2 * i + 1]), sl2, -lse2)) : 0.f;
                  pp[i] = pack_bf16x2(p0, p1);
                  ds[i] = pack_bf16x2(p0 * (__uint_as_float(dv[2 * i]) - del), p1 * (__uint_as_float(dv[2 * i + 1]) - del));
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) pp[i] = ds[i] = 0u;
            }
            store_row32(sP, row, col, pp);
            store_row32(sdS, row, col, ds);
          }
          fence_proxy_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(pds_full);
          if (cw == 0 && lane == 0) trace_stamp(p.trace, it, 7);
          mbar_wait(grads_done, uint32_t(it) & 1u);
          tc_fence_after();
          if (cw == 0 && lane == 0) trace_stamp(p.trace, it, 8);
          // ---- epilogues: rows = keys for dV / dK (after the last query tile), queries for dQ (after the last key tile).
          //      Staging tiles alias the P / dS panels (free until the next block's P / dS are written); one TMA store
          //      per tile, rows past the sequence end clipped by the TMA unit.
          const bool last_q = qt == p.q_tiles - 1, last_k = kt == p.k_tiles - 1;
          if (last_q) {
            acc_to_stage<HD>(trow + C::COL_DV, row, hf, h & 1, sP, 1.0f);
            acc_to_stage<HD>(trow + C::COL_DK, row, hf, h & 1, sP + TILE_BYTES, p.scale);
          }
          if (last_k) acc_to_stage<HD>(trow + C::COL_DQ + uint32_t(qt * 64), row, hf, h & 1, sdS, p.scale);
          tc_fence_before();     // accumulator reads are ordered before the arrive on the next block's pds_full
          staged = last_q || last_k;
          if (staged) {
            fence_proxy_async_smem();
            compute_bar();
            if (is_issuer) {
              if (last_q) {
                tma_store_3d(&tmdV, sP, h * HD, kt * TQ, b);
                tma_store_3d(&tmdK, sP + TILE_BYTES, h * HD, kt * TQ, b);
              }
              if (last_k) tma_store_3d(&tmdQ, sdS, h * HD, qt * TQ, b);
              bulk_commit_group();
            }
          }
          if (cw == 0 && lane == 0) trace_stamp(p.trace, it, 9);
        }
      }
    }
    if (is_issuer) bulk_wait_all();     // the last stores have left shared memory before the CTA does
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

int make_maps(CUtensorMap* out, const void* ptr, int64_t ld, int B, int N, int width_cols) {
  // [B][N][width] view with row pitch ld; box = 64 columns x 128 rows x 1 sample; rows past N are zero-filled
  return make_tmap_3d_bf16(out, ptr, (uint64_t)width_cols, (uint64_t)N, (uint64_t)B, (uint64_t)ld, (uint64_t)N * ld, 64, TQ, 1);
}

// store map over the same [B][N][width] view: box = head_dim columns x 128 rows x 1 sample
int make_store_map(CUtensorMap* out, const void* ptr, int64_t ld, int B, int N, int width_cols, int head_dim) {
  return make_tmap_3d_bf16_store(out, ptr, (uint64_t)width_cols, (uint64_t)N, (uint64_t)B, (uint64_t)ld, (uint64_t)N * ld,
                                 head_dim, TQ, 1);
}

template <int HD, int KBOX>
int launch_ws_fwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& to,
                  const WsFwdParams& p, cudaStream_t st) {
  using C = WsFwdCfg<KBOX>;
  auto kern = attn_ws_fwd_kernel<HD, KBOX>;
  static bool configured = false;
  if (!configured) {
    MMAE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    configured = true;
  }
  launch_k(kern, dim3(std::min(p.items, sm_count())), WS_THREADS, C::SMEM, st, tq, tk, tv, to, p);
  count_launch();
  MMAE_LAUNCH_OK();
  return MMAE_OK;
}

template <int HD>
int launch_ws_bwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& tdo, const CUtensorMap& to,
                  const CUtensorMap& tdq, const CUtensorMap& tdk, const CUtensorMap& tdv, const WsBwdParams& p, cudaStream_t st) {
  auto kern = attn_ws_bwd_kernel<HD>;
  static bool configured = false;
  if (!configured) {
    MMAE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, WsBwdCfg::SMEM));
    configured = true;
  }
  launch_k(kern, dim3(std::min(p.items, sm_count())), WS_THREADS, WsBwdCfg::SMEM, st, tq, tk, tv, tdo, to, tdq, tdk, tdv, p);
  count_launch();
  MMAE_LAUNCH_OK();
  return MMAE_OK;
}

long long* g_ws_trace = nullptr;

}  // namespace

// forward: any Nq, Nk <= 256, head_dim 64 or 32 (even head count)
bool attn_ws_fwd_supported(int H, int Nq, int Nk, int head_dim) {
  return Nq >= 1 && Nk >= 1 && Nk <= 256 && (head_dim == 64 || (head_dim == 32 && H % 2 == 0));
}
// backward: Nq <= 256 (one dQ accumulator per query tile stays in TMEM), any Nk, head_dim 64 or 32 (even head count)
bool attn_ws_bwd_supported(int H, int Nq, int Nk, int head_dim) {
  return Nq >= 1 && Nk >= 1 && Nq <= 256 && (head_dim == 64 || (head_dim == 32 && H % 2 == 0));
}

int attn_ws_forward(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o, int64_t ldo,
                    float* lse, int B, int H, int Nq, int Nk, int head_dim, float scale, cudaStream_t st) {
  CUtensorMap tq, tk, tv, to;
  int rc;
  const int width = H * head_dim;
  if ((rc = make_maps(&tq, q, ldq, B, Nq, width)) || (rc = make_maps(&tk, k, ldk, B, Nk, width)) ||
      (rc = make_maps(&tv, v, ldv, B, Nk, width)) || (rc = make_store_map(&to, o, ldo, B, Nq, width, head_dim)))
    return rc;
  WsFwdParams p;
  p.B = B; p.H = H; p.Nq = Nq; p.Nk = Nk;
  p.q_tiles = ceil_div(Nq, TQ);
  p.items = B * H * p.q_tiles;
  p.scale = scale;
  p.O = reinterpret_cast<bf16*>(o);
  p.ldo = ldo;
  p.lse = lse;
  p.trace = g_ws_trace;
  const bool two = Nk > 128;
  if (head_dim == 64) return two ? launch_ws_fwd<64, 2>(tq, tk, tv, to, p, st) : launch_ws_fwd<64, 1>(tq, tk, tv, to, p, st);
  return two ? launch_ws_fwd<32, 2>(tq, tk, tv, to, p, st) : launch_ws_fwd<32, 1>(tq, tk, tv, to, p, st);
}

int attn_ws_backward(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* o,
                     int64_t ldo, const void* d_o, int64_t lddo, const float* lse, void* dq, int64_t lddq, void* dk, int64_t lddk,
                     void* dv, int64_t lddv, int B, int H, int Nq, int Nk, int head_dim, float scale, cudaStream_t st) {
  CUtensorMap tq, tk, tv, tdo, to, tdq, tdk, tdv;
  int rc;
  const int width = H * head_dim;
  if ((rc = make_maps(&tq, q, ldq, B, Nq, width)) || (rc = make_maps(&tk, k, ldk, B, Nk, width)) ||
      (rc = make_maps(&tv, v, ldv, B, Nk, width)) || (rc = make_maps(&tdo, d_o, lddo, B, Nq, width)) ||
      (rc = make_maps(&to, o, ldo, B, Nq, width)) ||
      (rc = make_store_map(&tdq, dq, lddq, B, Nq, width, head_dim)) ||
      (rc = make_store_map(&tdk, dk, lddk, B, Nk, width, head_dim)) ||
      (rc = make_store_map(&tdv, dv, lddv, B, Nk, width, head_dim)))
    return rc;
  WsBwdParams p;
  p.B = B; p.H = H; p.Nq = Nq; p.Nk = Nk;
  p.q_tiles = ceil_div(Nq, TQ);
  p.k_tiles = ceil_div(Nk, TQ);
  p.items = B * H;
  p.scale = scale;
  p.lse = lse; p.delta = nullptr;
  p.dQ = reinterpret_cast<bf16*>(dq); p.dK = reinterpret_cast<bf16*>(dk); p.dV = reinterpret_cast<bf16*>(dv);
  p.lddq = lddq; p.lddk = lddk; p.lddv = lddv;
  p.trace = g_ws_trace;
  return head_dim == 64 ? launch_ws_bwd<64>(tq, tk, tv, tdo, to, tdq, tdk, tdv, p, st)
                        : launch_ws_bwd<32>(tq, tk, tv, tdo, to, tdq, tdk, tdv, p, st);
}

}  // namespace mmae

// diagnostics: device buffer of 64 x 16 clock64 stamps written by CTA 0 of the next attention_ws launches (NULL: off)
extern "C" int mmae_attention_ws_set_trace(long long* device_buffer) {
  mmae::g_ws_trace = device_buffer;
  return MMAE_OK;
}

// Fused multi-head attention for the ViT encoder (self, head_dim 64) and the MAE decoder (cross,
// head_dim 32) on tcgen05 tensor cores:  O = softmax(Q K^T * scale) V,  plus its backward.
// Replaces F.scaled_dot_product_attention at cinema/vit.py:505-511 and the permute / unbind
// copies around it (:498-500, :519): Q/K/V/O are addressed in place inside the projection
// outputs through 4-D TMA tensor maps (dim, head, token, batch).
//
// Forward, one CTA per (batch, head, 256 queries):
//   warp 8  : TMA producer  (Q once, K_j / V_j double buffered)
//   warp 9  : MMA issuer    S_t = Q_t K_j^T  (128x128xd, SS)   and   O_t += P_t V_j  (128xdx128, SS)
//   warps 0-3 / 4-7 : softmax warpgroup for query tile t = 0 / 1, one query row per thread
//             (thread == TMEM lane, so row max / row sum need no shuffles): S row TMEM -> registers,
//             online softmax in the log2 domain with lazy rescaling of O (only when the running max
//             grows by more than 2^8), P -> bf16 -> TMEM (tcgen05.st) as the A operand of P.V.
// The two query tiles ping-pong on the tensor pipe: S_t(j+1) is issued as soon as warpgroup t has
// pulled S_t(j) into registers, so the tensor core runs under the exp / max / sum work.
// TMEM: S0 | S1 (128 cols each) | P0 | P1 (64 cols each, packed bf16) | O0 | O1 (head_dim cols each).
#include "../../include/cinema_b200.h"
#include "common.cuh"

#include <cstdlib>
#include <type_traits>

// In-kernel clock trace (diagnostics, cb_attention_trace): when a device buffer is registered, ONE CTA in the middle of
// the grid stamps clock64() at the phase boundaries of one softmax / compute warp per role and of its MMA warp
// (slot layout: tools/attn_trace.py).  A null pointer (the default) costs one predicated-off store per stamp.
__device__ long long* g_attn_trace = nullptr;

extern "C" int cb_attention_trace(long long* device_buf) {
  CB_CUDA(cudaMemcpyToSymbol(g_attn_trace, &device_buf, sizeof(device_buf)));
  return 0;
}

namespace {

#define CB_TR(slot)                                  \
  do {                                               \
    if (tron) trace[(slot)] = clock64();             \
  } while (0)

constexpr int TQ = 128;       // query rows per tile (UMMA M)
constexpr int TK = 128;       // keys per tile (UMMA N of S, K extent of P.V)
constexpr int FWD_THREADS = 384;  // 2 softmax warpgroups + 1 light warpgroup (TMA warp, MMA warp, 2 idle warps)
constexpr float LOG2E = 1.4426950408889634f;
constexpr float RESCALE_THRESHOLD = 8.0f;  // log2 units

struct AttnFwdArgs {
  bf16* o;
  long long o_sb, o_sn, o_sh;
  float* lse;  // [B, H, Nq]
  int B, H, Nq, Nk;
  float scale_log2;  // scale * log2(e)
};

template <int D>
struct FwdCfg {
  static constexpr int ROW_BYTES = D * 2;                    // 128 (SW128) or 64 (SW64)
  static constexpr uint64_t SWZ = D == 64 ? UMMA_SW128 : UMMA_SW64;
  static constexpr int GROUP_BYTES = 8 * ROW_BYTES;           // 8-row swizzle group (SBO)
  static constexpr int TILE_BYTES = TQ * ROW_BYTES;           // Q / K / V tile
  static constexpr int OFF_Q = 0;                             // 2 tiles
  static constexpr int OFF_K = OFF_Q + 2 * TILE_BYTES;        // 2 stages
  static constexpr int OFF_V = OFF_K + 2 * TILE_BYTES;        // 2 stages
  static constexpr int OFF_BAR = OFF_V + 2 * TILE_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
  static constexpr int TMEM_P = 256;                          // P0 as packed bf16 (64 columns); P1 at +64
  static constexpr int TMEM_O = 384;                          // column of O0; O1 at +D
};

// element (row, col) of a K-major 128B-swizzled [128 x 64] bf16 chunk -> byte offset of its 16-byte vector
__device__ __forceinline__ uint32_t sw128_vec_offset(int row, int vec /*0..7*/) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((vec ^ (row & 7)) << 4));
}

template <int D>
__global__ void __launch_bounds__(FWD_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const AttnFwdArgs p) {
  using C = FwdCfg<D>;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;   // [2]
  uint64_t* k_empty = bars + 3;  // [2]
  uint64_t* v_full = bars + 5;   // [2]
  uint64_t* v_empty = bars + 7;  // [2]
  uint64_t* s_full = bars + 9;   // [2] per query tile
  uint64_t* s_free = bars + 11;  // [2]
  uint64_t* p_full = bars + 13;  // [2]
  uint64_t* pv_done = bars + 15; // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_blk = blockIdx.x;  // 256-query block
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int q0 = q_blk * 2 * TQ;
  const int n_tiles_q = (p.Nq - q0 > TQ) ? 2 : 1;  // is the second 128-row tile populated?
  const int n_kv = (p.Nk + TK - 1) / TK;

  if (warp == 9) {
    if (lane == 0) {
      mbar_init(q_full, 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&k_full[i], 1), mbar_init(&k_empty[i], 1);
        mbar_init(&v_full[i], 1), mbar_init(&v_empty[i], 1);
        mbar_init(&s_full[i], 1), mbar_init(&s_free[i], 4);
        mbar_init(&p_full[i], 4), mbar_init(&pv_done[i], 1);
      }
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
  }
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // PDL: setup done, wait for the kernels that produce Q / K / V (/ dO) before the first load

  if (warp >= 8) {
   // the light warpgroup hands registers to the softmax warpgroups (8 x 232 + 4 x 40 == 12 x 168)
   asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
   if (warp == 8) {
    // ------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_expect_tx(q_full, n_tiles_q * C::TILE_BYTES);
      for (int t = 0; t < n_tiles_q; ++t)
        tma_load_4d(smem + C::OFF_Q + t * C::TILE_BYTES, &tm_q, q_full, 0, h, q0 + t * TQ, b);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_expect_tx(&k_full[s], C::TILE_BYTES);
        tma_load_4d(smem + C::OFF_K + s * C::TILE_BYTES, &tm_k, &k_full[s], 0, h, j * TK, b);
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_expect_tx(&v_full[s], C::TILE_BYTES);
        tma_load_4d(smem + C::OFF_V + s * C::TILE_BYTES, &tm_v, &v_full[s], 0, h, j * TK, b);
      }
    }
    __syncwarp();
   } else if (warp == 9) {
    // ------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(TQ, TK, false, false);
      constexpr uint32_t idesc_o = umma_idesc_bf16(TQ, D, false, true);
      const uint32_t q_base = smem_u32(smem + C::OFF_Q);
      const uint32_t k_base = smem_u32(smem + C::OFF_K);
      const uint32_t v_base = smem_u32(smem + C::OFF_V);
      auto issue_s = [&](int t, int j) {
        const uint32_t ks = k_base + (j & 1) * C::TILE_BYTES;
        const uint32_t qs = q_base + t * C::TILE_BYTES;
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint64_t da = umma_smem_desc(qs + kk * 32, 0, C::GROUP_BYTES, C::SWZ);
          const uint64_t db = umma_smem_desc(ks + kk * 32, 0, C::GROUP_BYTES, C::SWZ);
          umma_bf16_ss(tmem_base + t * TK, da, db, idesc_s, kk > 0 ? 1u : 0u);
        }
        umma_commit(&s_full[t]);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tcgen05_fence_after();
      for (int t = 0; t < n_tiles_q; ++t) issue_s(t, 0);
      if (n_kv == 1) umma_commit(&k_empty[0]);
      for (int j = 0; j < n_kv; ++j) {
        if (j + 1 < n_kv) {
          mbar_wait(&k_full[(j + 1) & 1], ((j + 1) >> 1) & 1);
          for (int t = 0; t < n_tiles_q; ++t) {
            mbar_wait(&s_free[t], j & 1);  // warpgroup t holds S_t(j) in registers
            tcgen05_fence_after();
            issue_s(t, j + 1);
          }
          // K stage of tile j is free once S(j) retired, which happened before s_free; the commit below
          // covers S(j+1) as well, which is harmless: it only delays the refill of the *other* stage.
          umma_commit(&k_empty[j & 1]);
        }
        mbar_wait(&v_full[j & 1], (j >> 1) & 1);
        const uint32_t vs = v_base + (j & 1) * C::TILE_BYTES;
        for (int t = 0; t < n_tiles_q; ++t) {
          mbar_wait(&p_full[t], j & 1);
          tcgen05_fence_after();
          // A = P_t straight from TMEM (16 keys = 8 packed columns per step): an smem A operand would cost 4 KB of
          // shared-memory reads per instruction and pace these narrow (N = head_dim) MMAs at ~64 clk each
#pragma unroll
          for (int kk = 0; kk < TK / 16; ++kk) {
            const uint64_t db = umma_smem_desc(vs + kk * 2 * C::GROUP_BYTES, 0, C::GROUP_BYTES, C::SWZ);
            umma_bf16_ts(tmem_base + C::TMEM_O + t * D, tmem_base + C::TMEM_P + t * 64 + kk * 8, db, idesc_o,
                         (j > 0 || kk > 0) ? 1u : 0u);
          }
          umma_commit(&pv_done[t]);
        }
        umma_commit(&v_empty[j & 1]);
      }
    }
    __syncwarp();
   }
  } else {
    // ------------------------------------------------ softmax warpgroups
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int t = warp >> 2;             // query tile of this warpgroup
    const int r = (warp & 3) * 32 + lane;  // row inside the tile == TMEM lane
    if (t < n_tiles_q) {
      const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
      float m_ref = -INFINITY;  // reference max (log2 domain) the accumulators are expressed against
      float l = 0.f;
      const float sc = p.scale_log2;
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(&s_full[t], j & 1);
        tcgen05_fence_after();
        float s[TK];  // raw scores; scale and running max are folded into the exp2 argument below
        {
          uint32_t v[4][32];
#pragma unroll
          for (int c = 0; c < TK / 32; ++c) tmem_ld_32x32b_x32(lane_addr + t * TK + c * 32, v[c]);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < TK / 32; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i) s[c * 32 + i] = __uint_as_float(v[c][i]);
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[t]);

        const int valid = p.Nk - j * TK;  // keys of this tile that exist
        if (valid < TK) {                 // ragged last tile only (CTA-uniform); kept out of line on purpose
          asm volatile("" ::: "memory");
#pragma unroll
          for (int i = 0; i < TK; ++i)
            if (i >= valid) s[i] = -INFINITY;
          asm volatile("" ::: "memory");
        }
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < TK; i += 8) {
#pragma unroll
          for (int u = 0; u < 4; ++u) mx4[u] = fmaxf(mx4[u], fmaxf(s[i + 2 * u], s[i + 2 * u + 1]));
        }
        const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * sc;  // scale > 0
        // lazy rescale: keep m_ref unless the row max ran away by more than 2^8
        const bool grow = mx > m_ref + RESCALE_THRESHOLD;
        const float m_new = grow ? mx : m_ref;
        const float alpha = grow ? ex2_approx(m_ref - m_new) : 1.0f;  // 2^(-inf) = 0 on the first tile
        if (j > 0) {
          mbar_wait(&pv_done[t], (j - 1) & 1);  // P smem and O_t are quiescent
          tcgen05_fence_after();
          if (__any_sync(0xffffffffu, grow)) {
#pragma unroll
            for (int c = 0; c < D / 16; ++c) {
              uint32_t o[16];
              tmem_ld_32x32b_x16(lane_addr + C::TMEM_O + t * D + c * 16, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st_32x32b_x16(lane_addr + C::TMEM_O + t * D + c * 16, o);
            }
            tmem_st_wait();
          }
        }
        l *= alpha;
        m_ref = m_new;
        const float neg_m = -m_ref;
        float sum4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int blk = 0; blk < TK / 32; ++blk) {  // 32 keys -> 16 packed columns of this row's P in TMEM
          uint32_t pw[16];
#pragma unroll
          for (int w = 0; w < 16; w += 2) {
            const int i0 = blk * 32 + w * 2;
            const float e0 = ex2_approx(fmaf(s[i0], sc, neg_m)), e1 = ex2_approx(fmaf(s[i0 + 1], sc, neg_m));
            const float e2 = ex2_approx(fmaf(s[i0 + 2], sc, neg_m)), e3 = ex2_approx(fmaf(s[i0 + 3], sc, neg_m));
            sum4[0] += e0, sum4[1] += e1, sum4[2] += e2, sum4[3] += e3;
            pw[w] = pack_bf16(e0, e1), pw[w + 1] = pack_bf16(e2, e3);
          }
          tmem_st_32x32b_x16(lane_addr + C::TMEM_P + t * 64 + blk * 16, pw);
        }
        tmem_st_wait();
        l += (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
      }
      // epilogue: O / l -> bf16, LSE
      mbar_wait(&pv_done[t], (n_kv - 1) & 1);
      tcgen05_fence_after();
      const int q_row = q0 + t * TQ + r;
      const float inv_l = 1.0f / l;
      bf16* o_ptr = p.o + (long long)b * p.o_sb + (long long)q_row * p.o_sn + (long long)h * p.o_sh;
#pragma unroll
      for (int c = 0; c < D / 16; ++c) {
        uint32_t o[16];
        tmem_ld_32x32b_x16(lane_addr + C::TMEM_O + t * D + c * 16, o);
        tmem_ld_wait();
        if (q_row < p.Nq) {
          uint4 w0 = make_uint4(pack_bf16(__uint_as_float(o[0]) * inv_l, __uint_as_float(o[1]) * inv_l),
                                pack_bf16(__uint_as_float(o[2]) * inv_l, __uint_as_float(o[3]) * inv_l),
                                pack_bf16(__uint_as_float(o[4]) * inv_l, __uint_as_float(o[5]) * inv_l),
                                pack_bf16(__uint_as_float(o[6]) * inv_l, __uint_as_float(o[7]) * inv_l));
          uint4 w1 = make_uint4(pack_bf16(__uint_as_float(o[8]) * inv_l, __uint_as_float(o[9]) * inv_l),
                                pack_bf16(__uint_as_float(o[10]) * inv_l, __uint_as_float(o[11]) * inv_l),
                                pack_bf16(__uint_as_float(o[12]) * inv_l, __uint_as_float(o[13]) * inv_l),
                                pack_bf16(__uint_as_float(o[14]) * inv_l, __uint_as_float(o[15]) * inv_l));
          reinterpret_cast<uint4*>(o_ptr + c * 16)[0] = w0;
          reinterpret_cast<uint4*>(o_ptr + c * 16)[1] = w1;
        }
      }
      if (q_row < p.Nq)
        p.lse[((long long)b * p.H + h) * p.Nq + q_row] = (m_ref + log2f(l)) * (1.0f / LOG2E);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 9) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// Forward, second generation (default): ONE 128-query tile per CTA, 256 threads, TWO CTAs per SM.
//   warps 0-3 : softmax warpgroup (thread == query row == TMEM lane)          216 registers (setmaxnreg)
//   warp 4    : TMA producer (Q once, K_j / V_j double buffered)              40 registers
//   warp 5    : MMA issuer (S = Q K_j^T, O += P V_j with P as the TMEM A operand)
//   warps 6-7 : idle (they only make the light warpgroup complete for setmaxnreg)
// TMEM per CTA: S (128 columns) | P (64, packed bf16) | O (head_dim)  = 256 columns, so two CTAs share the SM's 512.
// Why: the first-generation kernel (two query tiles ping-ponging inside one 384-thread CTA) spends ~7 k clocks per CTA in
// launch / TMEM allocation / first loads / epilogue and runs its two softmax warpgroups in lock step, so the MUFU pipe
// idles during every row-max / TMEM-load phase (profiles/r01_attention_clock_trace.md: 3.5 k clocks per key tile against
// a 2.0 k MUFU bound).  Two independent CTAs per SM desynchronise by themselves: one CTA's exp2 phase runs under the
// other's prologue, row max, rescale or epilogue, and the hardware block scheduler balances ragged tiles.
// Ragged shapes cost what they contain: the last key tile is evaluated over ceil(valid / 32) * 32 columns only (S MMA
// with a narrower N, fewer exponentials, fewer P.V k-steps), and warps whose 32 query rows are all past Nq skip the
// softmax arithmetic (they still take part in the barrier protocol).
// ---------------------------------------------------------------------------------------------
constexpr int FWD2_THREADS = 256;

template <int D>
struct Fwd2Cfg {
  static constexpr int ROW_BYTES = D * 2;
  static constexpr uint64_t SWZ = D == 64 ? UMMA_SW128 : UMMA_SW64;
  static constexpr int GROUP_BYTES = 8 * ROW_BYTES;
  static constexpr int TILE_BYTES = TQ * ROW_BYTES;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + TILE_BYTES;      // 2 stages
  static constexpr int OFF_V = OFF_K + 2 * TILE_BYTES;  // 2 stages
  static constexpr int OFF_BAR = OFF_V + 2 * TILE_BYTES;
  static constexpr int SMEM_USED = OFF_BAR + 256 + 1024;
  // at least 80 KB are requested so that never more than two CTAs are resident per SM: a third one would hold
  // registers while it waits for TMEM columns, and the setmaxnreg.inc of the first two could starve
  static constexpr int SMEM_BYTES = SMEM_USED > 80 * 1024 ? SMEM_USED : 80 * 1024;
  static constexpr int TMEM_COLS = 256;
  static constexpr int TMEM_S = 0, TMEM_P = 128, TMEM_O = 192;
  static_assert(TMEM_O + D <= TMEM_COLS, "TMEM budget");
};

template <int D>
__global__ void __launch_bounds__(FWD2_THREADS, 2)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                 const __grid_constant__ CUtensorMap tm_v, const AttnFwdArgs p) {
  using C = Fwd2Cfg<D>;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;   // [2]
  uint64_t* k_empty = bars + 3;  // [2]
  uint64_t* v_full = bars + 5;   // [2]
  uint64_t* v_empty = bars + 7;  // [2]
  uint64_t* s_full = bars + 9;
  uint64_t* s_free = bars + 10;  // 4 warp arrivals
  uint64_t* p_full = bars + 11;  // 4 warp arrivals
  uint64_t* pv_done = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int q0 = blockIdx.x * TQ;
  const int n_kv = (p.Nk + TK - 1) / TK;
  const int last_nc = (p.Nk - (n_kv - 1) * TK + 31) >> 5;  // 32-column chunks of the last key tile that hold keys (1..4)
  long long* const trace = g_attn_trace;
  const bool tron = trace != nullptr && lane == 0 && blockIdx.x == 5 && blockIdx.y == 3 && blockIdx.z == gridDim.z / 2;
  if (warp == 0) CB_TR(240);

  if (warp == 5) {
    if (lane == 0) {
      mbar_init(q_full, 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&k_full[i], 1), mbar_init(&k_empty[i], 1);
        mbar_init(&v_full[i], 1), mbar_init(&v_empty[i], 1);
      }
      mbar_init(s_full, 1), mbar_init(s_free, 4), mbar_init(p_full, 4), mbar_init(pv_done, 1);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, C::TMEM_COLS);
  }
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // PDL: setup done, wait for the kernels that produce Q / K / V before the first load
  if (warp == 0) CB_TR(241);

  if (warp >= 4) {
    // 2 CTAs x (4 x 32 x 216 + 4 x 32 x 40) registers == 64 K: the light warpgroup hands its share to the softmax warps
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 4) {
      // ------------------------------------------------ TMA producer
      if (lane == 0) {
        mbar_expect_tx(q_full, C::TILE_BYTES);
        tma_load_4d(smem + C::OFF_Q, &tm_q, q_full, 0, h, q0, b);
        for (int j = 0; j < n_kv; ++j) {
          const int s = j & 1;
          const uint32_t ph = (j >> 1) & 1;
          mbar_wait(&k_empty[s], ph ^ 1);
          mbar_expect_tx(&k_full[s], C::TILE_BYTES);
          tma_load_4d(smem + C::OFF_K + s * C::TILE_BYTES, &tm_k, &k_full[s], 0, h, j * TK, b);
          mbar_wait(&v_empty[s], ph ^ 1);
          mbar_expect_tx(&v_full[s], C::TILE_BYTES);
          tma_load_4d(smem + C::OFF_V + s * C::TILE_BYTES, &tm_v, &v_full[s], 0, h, j * TK, b);
        }
      }
      __syncwarp();
    } else if (warp == 5) {
      // ------------------------------------------------ MMA issuer
      // The WHOLE warp runs this loop converged (every lane waits on the barriers) and one elected lane issues: under
      // `if (lane == 0)` nvcc wraps every tcgen05 instruction in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop and
      // rebuilds its descriptors in front of it (~67 clocks per MMA, ~100 per commit measured with the clock trace);
      // with the elect.sync predicate the UTCHMMA / UTCBAR instructions are issued back to back from hoisted operands.
      const bool leader = elect_one_sync() != 0;
      constexpr uint32_t idesc_o = umma_idesc_bf16(TQ, D, false, true);
      const uint32_t q_base = smem_u32(smem + C::OFF_Q);
      const uint32_t k_base = smem_u32(smem + C::OFF_K);
      const uint32_t v_base = smem_u32(smem + C::OFF_V);
      auto issue_s = [&](int j) {
        const uint32_t idesc_s = umma_idesc_bf16(TQ, j == n_kv - 1 ? last_nc * 32 : TK, false, false);
        const uint32_t ks = k_base + (j & 1) * C::TILE_BYTES;
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint64_t da = umma_smem_desc(q_base + kk * 32, 0, C::GROUP_BYTES, C::SWZ);
          const uint64_t db = umma_smem_desc(ks + kk * 32, 0, C::GROUP_BYTES, C::SWZ);
          umma_bf16_ss(tmem_base + C::TMEM_S, da, db, idesc_s, kk > 0 ? 1u : 0u);
        }
        umma_commit(s_full);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tcgen05_fence_after();
      if (leader) issue_s(0);
      __syncwarp();
      for (int j = 0; j < n_kv; ++j) {
        if (j + 1 < n_kv) {
          mbar_wait(&k_full[(j + 1) & 1], ((j + 1) >> 1) & 1);
          mbar_wait(s_free, j & 1);  // the softmax warps hold S(j) in registers
          if (j < 8) CB_TR(16 * j + 8);
          tcgen05_fence_after();
          if (leader) {
            issue_s(j + 1);
            umma_commit(&k_empty[j & 1]);  // (covers S(j+1) too: that only delays the refill of the OTHER stage's successor)
          }
          __syncwarp();
          if (j < 8) CB_TR(16 * j + 9);
        }
        mbar_wait(&v_full[j & 1], (j >> 1) & 1);
        mbar_wait(p_full, j & 1);
        if (j < 8) CB_TR(16 * j + 10);
        tcgen05_fence_after();
        if (leader) {
          const uint32_t vs = v_base + (j & 1) * C::TILE_BYTES;
          const int n_kk = j == n_kv - 1 ? last_nc * 2 : TK / 16;  // 16 keys per step; the ragged tile stops early
#pragma unroll
          for (int kk = 0; kk < TK / 16; ++kk) {
            if (kk < n_kk) {
              const uint64_t db = umma_smem_desc(vs + kk * 2 * C::GROUP_BYTES, 0, C::GROUP_BYTES, C::SWZ);
              umma_bf16_ts(tmem_base + C::TMEM_O, tmem_base + C::TMEM_P + kk * 8, db, idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
            }
          }
          umma_commit(pv_done);
          umma_commit(&v_empty[j & 1]);
        }
        __syncwarp();
        if (j < 8) CB_TR(16 * j + 11);
      }
    }
  } else {
    // ------------------------------------------------ softmax warpgroup
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    const int r = warp * 32 + lane;  // row inside the tile == TMEM lane
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const bool warp_live = q0 + warp * 32 < p.Nq;  // warp-uniform: do any of this warp's 32 rows exist?
    float m_ref = -INFINITY;  // reference max (log2 domain) the accumulators are expressed against
    float l = 0.f;
    const float sc = p.scale_log2;

    // one key tile of NC * 32 columns
    auto tile = [&](auto nc_c, const int j) {
      constexpr int NC = decltype(nc_c)::value;
      constexpr int W = NC * 32;
      const bool trj = warp == 0 && j < 8;
      if (trj) CB_TR(16 * j + 0);
      mbar_wait(s_full, j & 1);
      if (trj) CB_TR(16 * j + 1);
      tcgen05_fence_after();
      float s[W];  // raw scores; scale and running max are folded into the exp2 argument below
      if (warp_live) {
        uint32_t v[NC][32];
#pragma unroll
        for (int c = 0; c < NC; ++c) tmem_ld_32x32b_x32(lane_addr + C::TMEM_S + c * 32, v[c]);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i) s[c * 32 + i] = __uint_as_float(v[c][i]);
      }
      if (trj) CB_TR(16 * j + 2);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);

      float alpha = 1.0f;
      bool grow = false;
      if (warp_live) {
        const int valid = p.Nk - j * TK;  // keys of this tile that exist
        if (valid < W) {                  // ragged last tile only (CTA-uniform); kept out of line on purpose
          asm volatile("" ::: "memory");
#pragma unroll
          for (int i = 0; i < W; ++i)
            if (i >= valid) s[i] = -INFINITY;
          asm volatile("" ::: "memory");
        }
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < W; i += 8) {
#pragma unroll
          for (int u = 0; u < 4; ++u) mx4[u] = fmaxf(mx4[u], fmaxf(s[i + 2 * u], s[i + 2 * u + 1]));
        }
        const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * sc;  // scale > 0
        // lazy rescale: keep m_ref unless the row max ran away by more than 2^8
        grow = mx > m_ref + RESCALE_THRESHOLD;
        const float m_new = grow ? mx : m_ref;
        alpha = grow ? ex2_approx(m_ref - m_new) : 1.0f;  // 2^(-inf) = 0 on the first tile
        m_ref = m_new;
      }
      if (trj) CB_TR(16 * j + 3);
      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1);  // P and O are quiescent
        if (trj) CB_TR(16 * j + 4);
        tcgen05_fence_after();
        if (warp_live && __any_sync(0xffffffffu, grow)) {
#pragma unroll
          for (int c = 0; c < D / 16; ++c) {
            uint32_t o[16];
            tmem_ld_32x32b_x16(lane_addr + C::TMEM_O + c * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32b_x16(lane_addr + C::TMEM_O + c * 16, o);
          }
          tmem_st_wait();
        }
      }
      if (warp_live) {
        l *= alpha;
        const float neg_m = -m_ref;
        float sum4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int blk = 0; blk < NC; ++blk) {  // 32 keys -> 16 packed columns of this row's P in TMEM
          uint32_t pw[16];
#pragma unroll
          for (int w = 0; w < 16; w += 2) {
            const int i0 = blk * 32 + w * 2;
            const float e0 = ex2_approx(fmaf(s[i0], sc, neg_m)), e1 = ex2_approx(fmaf(s[i0 + 1], sc, neg_m));
            const float e2 = ex2_approx(fmaf(s[i0 + 2], sc, neg_m)), e3 = ex2_approx(fmaf(s[i0 + 3], sc, neg_m));
            sum4[0] += e0, sum4[1] += e1, sum4[2] += e2, sum4[3] += e3;
            pw[w] = pack_bf16(e0, e1), pw[w + 1] = pack_bf16(e2, e3);
          }
          tmem_st_32x32b_x16(lane_addr + C::TMEM_P + blk * 16, pw);
        }
        tmem_st_wait();
        l += (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
      }
      if (trj) CB_TR(16 * j + 5);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    };

    for (int j = 0; j + 1 < n_kv; ++j) tile(std::integral_constant<int, 4>{}, j);
    switch (last_nc) {
      case 1: tile(std::integral_constant<int, 1>{}, n_kv - 1); break;
      case 2: tile(std::integral_constant<int, 2>{}, n_kv - 1); break;
      case 3: tile(std::integral_constant<int, 3>{}, n_kv - 1); break;
      default: tile(std::integral_constant<int, 4>{}, n_kv - 1); break;
    }

    // epilogue: O / l -> bf16, LSE
    if (warp == 0) CB_TR(242);
    mbar_wait(pv_done, (n_kv - 1) & 1);
    if (warp == 0) CB_TR(243);
    tcgen05_fence_after();
    const int q_row = q0 + r;
    if (warp_live) {
      const float inv_l = 1.0f / l;
      bf16* o_ptr = p.o + (long long)b * p.o_sb + (long long)q_row * p.o_sn + (long long)h * p.o_sh;
#pragma unroll
      for (int c = 0; c < D / 16; ++c) {
        uint32_t o[16];
        tmem_ld_32x32b_x16(lane_addr + C::TMEM_O + c * 16, o);
        tmem_ld_wait();
        if (q_row < p.Nq) {
          uint4 w0 = make_uint4(pack_bf16(__uint_as_float(o[0]) * inv_l, __uint_as_float(o[1]) * inv_l),
                                pack_bf16(__uint_as_float(o[2]) * inv_l, __uint_as_float(o[3]) * inv_l),
                                pack_bf16(__uint_as_float(o[4]) * inv_l, __uint_as_float(o[5]) * inv_l),
                                pack_bf16(__uint_as_float(o[6]) * inv_l, __uint_as_float(o[7]) * inv_l));
          uint4 w1 = make_uint4(pack_bf16(__uint_as_float(o[8]) * inv_l, __uint_as_float(o[9]) * inv_l),
                                pack_bf16(__uint_as_float(o[10]) * inv_l, __uint_as_float(o[11]) * inv_l),
                                pack_bf16(__uint_as_float(o[12]) * inv_l, __uint_as_float(o[13]) * inv_l),
                                pack_bf16(__uint_as_float(o[14]) * inv_l, __uint_as_float(o[15]) * inv_l));
          reinterpret_cast<uint4*>(o_ptr + c * 16)[0] = w0;
          reinterpret_cast<uint4*>(o_ptr + c * 16)[1] = w1;
        }
      }
      if (q_row < p.Nq)
        p.lse[((long long)b * p.H + h) * p.Nq + q_row] = (m_ref + log2f(l)) * (1.0f / LOG2E);
    }
    if (warp == 0) CB_TR(244);
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 5) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
  if (warp == 0) CB_TR(245);
}

template <int D>
int launch_fwd2(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnFwdArgs& a,
                cudaStream_t stream) {
  using C = Fwd2Cfg<D>;
  auto kern = attn_fwd2_kernel<D>;
  static bool attr_set = false;
  if (!attr_set) {
    CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  dim3 grid((a.Nq + TQ - 1) / TQ, a.H, a.B);
  cb_launch(kern, grid, FWD2_THREADS, C::SMEM_BYTES, stream, tq, tk, tv, a);
  CB_LAUNCH_CHECK();
  return 0;
}

// 4-D tensor map (dim, head, token, batch) over a strided bf16 view; box = {D, 1, rows, 1}
int make_qkv_tmap(CUtensorMap* m, const void* ptr, long long sb, long long sn, long long sh, int B, int H, int N, int D,
                  int rows) {
  uint64_t dims[4] = {(uint64_t)D, (uint64_t)H, (uint64_t)N, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)sh * 2, (uint64_t)sn * 2, (uint64_t)sb * 2};
  uint32_t box[4] = {(uint32_t)D, 1, (uint32_t)rows, 1};
  return cb_make_tmap_nd(m, ptr, 4, dims, strides, box, D * 2);
}

template <int D>
int launch_fwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnFwdArgs& a,
               cudaStream_t stream) {
  using C = FwdCfg<D>;
  auto kern = attn_fwd_kernel<D>;
  static bool attr_set = false;
  if (!attr_set) {
    CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  dim3 grid((a.Nq + 2 * TQ - 1) / (2 * TQ), a.H, a.B);
  cb_launch(kern, grid, FWD_THREADS, C::SMEM_BYTES, stream, tq, tk, tv, a);
  CB_LAUNCH_CHECK();
  return 0;
}

}  // namespace

extern "C" int cb_attention_fwd(const void* q, long long q_sb, long long q_sn, long long q_sh, const void* k,
                                long long k_sb, long long k_sn, long long k_sh, const void* v, long long v_sb,
                                long long v_sn, long long v_sh, void* o, long long o_sb, long long o_sn, long long o_sh,
                                float* lse, int B, int H, int Nq, int Nk, int head_dim, float scale, void* stream) {
  CB_CHECK_ARG(head_dim == 32 || head_dim == 64, "attention: head_dim %d not supported (32 or 64)", head_dim);
  CB_CHECK_ARG(B > 0 && H > 0 && Nq > 0 && Nk > 0, "attention: empty problem");
  CB_CHECK_ARG(o_sn % 8 == 0 && o_sh % 8 == 0 && o_sb % 8 == 0 && ((uintptr_t)o & 15) == 0,
               "attention: output view must keep 16-byte alignment per (token, head)");
  CUtensorMap tq, tk, tv;
  if (int rc = make_qkv_tmap(&tq, q, q_sb, q_sn, q_sh, B, H, Nq, head_dim, TQ)) return rc;
  if (int rc = make_qkv_tmap(&tk, k, k_sb, k_sn, k_sh, B, H, Nk, head_dim, TK)) return rc;
  if (int rc = make_qkv_tmap(&tv, v, v_sb, v_sn, v_sh, B, H, Nk, head_dim, TK)) return rc;
  AttnFwdArgs a;
  a.o = (bf16*)o, a.o_sb = o_sb, a.o_sn = o_sn, a.o_sh = o_sh, a.lse = lse;
  a.B = B, a.H = H, a.Nq = Nq, a.Nk = Nk, a.scale_log2 = scale * LOG2E;
  cudaStream_t s = (cudaStream_t)stream;
  static const int gen = [] {  // CB_ATTN_FWD=1 selects the first-generation kernel (A/B measurements)
    const char* e = getenv("CB_ATTN_FWD");
    return e != nullptr ? atoi(e) : 2;
  }();
  if (gen == 1) return head_dim == 64 ? launch_fwd<64>(tq, tk, tv, a, s) : launch_fwd<32>(tq, tk, tv, a, s);
  return head_dim == 64 ? launch_fwd2<64>(tq, tk, tv, a, s) : launch_fwd2<32>(tq, tk, tv, a, s);
}


// ============================================================================================
// Backward.  One CTA per (batch, head, 128-key tile); loop over 128-query tiles i:
//   S^T  = K Q_i^T            (M = keys, N = queries; lanes = keys)       SS, both K-major
//   dP^T = V dO_i^T                                                        SS, both K-major
//   P^T  = exp2(S^T * scale*log2e - lse_q * log2e)            \  two warpgroups, one key row per
//   dS^T = P^T o (dP^T - delta_q) * scale                     /  thread, 64 query columns each
//   dV  += P^T  dO_i          A = P^T in TMEM (packed bf16, TS-mode MMA), B = dO_i tile read MN-major
//   dK  += dS^T Q_i           A = dS^T in TMEM (head_dim 32) or smem (K-major), B = Q_i tile read MN-major
//   dQ_i = dS   K             A = dS^T smem read MN-major, B = K tile read MN-major
// dQ_i is drained from TMEM with fp32 red.global.add into dq_acc (summed over key tiles by the
// atomics) and converted to bf16 by a small follow-up kernel; dK / dV are written once at the end.
// The same Q_i / dO_i / K smem tiles serve as K-major and as MN-major operands: only the UMMA
// descriptors differ, nothing is transposed in memory.
// TMEM: S^T (128) | dP^T (128) | P^T (64, packed) | [dS^T (64, packed)] | dV (D) | dK (D) | dQ (D).
// ============================================================================================
namespace {

constexpr int BWD_THREADS = 384;  // 2 compute warpgroups + 1 light warpgroup (TMA warp, MMA warp, 2 idle warps)

struct AttnBwdArgs {
  const float* nl2;  // [B*H, NqP]  -lse * log2(e)      (attn_delta_kernel; NqP = Nq rounded up to 128, pads = -inf)
  const float* dsc;  // [B*H, NqP]  delta * scale       (pads = 0)
  int NqP;
  float* dq_acc;  // [B, H, Nq, D] fp32, pre-zeroed
  bf16* dk;
  long long dk_sb, dk_sn, dk_sh;
  bf16* dv;
  long long dv_sb, dv_sn, dv_sh;
  int B, H, Nq, Nk;
  float scale, scale_log2;
  int token;  // second-generation kernel: alternate the exp2 section between the two compute warpgroups
  float* dk_cs;  // optional [H * D] fp32: += column sums of the stored bf16 dK / dV (k / v projection bias gradients)
  float* dv_cs;
};

template <int D>
struct BwdCfg {
  static constexpr int ROW_BYTES = D * 2;
  static constexpr uint64_t SWZ = D == 64 ? UMMA_SW128 : UMMA_SW64;
  static constexpr int GROUP_BYTES = 8 * ROW_BYTES;
  static constexpr int TILE_BYTES = 128 * ROW_BYTES;
  static constexpr int PS_BYTES = 128 * 128 * 2;  // dS^T, two 64-column chunks of 16 KB
  static constexpr int OFF_K = 0;
  static constexpr int OFF_V = OFF_K + TILE_BYTES;
  static constexpr int OFF_Q = OFF_V + TILE_BYTES;        // 2 stages
  static constexpr int OFF_DO = OFF_Q + 2 * TILE_BYTES;   // 2 stages
  static constexpr int OFF_DS = OFF_DO + 2 * TILE_BYTES;
  static constexpr int OFF_VEC = OFF_DS + PS_BYTES;       // lse / delta: [2 stages][2][128] floats
  static constexpr int OFF_STG = OFF_VEC + 2 * 2 * 128 * 4;  // dQ drain transposition: 8 warps x (32 rows x 64 B)
  static constexpr int OFF_BAR = OFF_STG + 8 * 2048;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
  // A operands in TMEM (TS-mode MMA): P^T always; dS^T too when the 512 columns allow it (head_dim 32).  An smem A
  // operand costs a 4 KB shared-memory read per K=16 step, which paces these narrow (N = head_dim) MMAs at ~64 clk.
  static constexpr bool DS_TMEM = D <= 32;
  static constexpr int TM_S = 0, TM_DP = 128, TM_PT = 256, TM_DST = 320;
  static constexpr int TM_DV = DS_TMEM ? 384 : 320, TM_DK = TM_DV + D, TM_DQ = TM_DV + 2 * D;
  static_assert(TM_DQ + D <= 512, "TMEM budget");
};

__device__ __forceinline__ void red_add_v4f(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int D>
__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                const AttnBwdArgs p) {
  using C = BwdCfg<D>;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* kv_full = bars + 0;
  uint64_t* qdo_full = bars + 1;   // [2]
  uint64_t* qdo_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* dp_full = bars + 6;
  uint64_t* sdp_free = bars + 7;   // 8 warp arrivals
  uint64_t* ps_ready = bars + 8;   // 8 warp arrivals
  uint64_t* mma2_done = bars + 9;
  uint64_t* dq_free = bars + 10;   // 8 warp arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);
  float* vec = reinterpret_cast<float*>(smem + C::OFF_VEC);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kv0 = blockIdx.x * 128;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int n_q = (p.Nq + 127) / 128;

  if (warp == 9) {
    if (lane == 0) {
      mbar_init(kv_full, 1);
      for (int i = 0; i < 2; ++i) mbar_init(&qdo_full[i], 1), mbar_init(&qdo_empty[i], 1);
      mbar_init(s_full, 1), mbar_init(dp_full, 1), mbar_init(mma2_done, 1);
      mbar_init(sdp_free, 8), mbar_init(ps_ready, 8), mbar_init(dq_free, 8);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
  }
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tm_q), tma_prefetch_desc(&tm_k), tma_prefetch_desc(&tm_v), tma_prefetch_desc(&tm_do);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // PDL: setup done, wait for the kernels that produce Q / K / V (/ dO) before the first load

  if (warp >= 8) {
   // the light warpgroup hands registers to the compute warpgroups (8 x 232 + 4 x 40 == 12 x 168)
   asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
   if (warp == 8) {
    // ------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_expect_tx(kv_full, 2 * C::TILE_BYTES);
      tma_load_4d(smem + C::OFF_K, &tm_k, kv_full, 0, h, kv0, b);
      tma_load_4d(smem + C::OFF_V, &tm_v, kv_full, 0, h, kv0, b);
      for (int i = 0; i < n_q; ++i) {
        const int s = i & 1;
        mbar_wait(&qdo_empty[s], ((i >> 1) & 1) ^ 1);
        mbar_expect_tx(&qdo_full[s], 2 * C::TILE_BYTES + 1024);
        tma_load_4d(smem + C::OFF_Q + s * C::TILE_BYTES, &tm_q, &qdo_full[s], 0, h, i * 128, b);
        tma_load_4d(smem + C::OFF_DO + s * C::TILE_BYTES, &tm_do, &qdo_full[s], 0, h, i * 128, b);
        const long long voff = ((long long)b * p.H + h) * p.NqP + i * 128;  // per-query vectors of this tile
        bulk_load_1d(vec + s * 256, p.nl2 + voff, 512, &qdo_full[s]);
        bulk_load_1d(vec + s * 256 + 128, p.dsc + voff, 512, &qdo_full[s]);
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, false, false);
      constexpr uint32_t idesc_kv = umma_idesc_bf16(128, D, false, true);  // dV, dK
      constexpr uint32_t idesc_dq = umma_idesc_bf16(128, D, true, true);   // dQ
      const uint32_t k_base = smem_u32(smem + C::OFF_K);
      const uint32_t v_base = smem_u32(smem + C::OFF_V);
      const uint32_t q_base = smem_u32(smem + C::OFF_Q);
      const uint32_t do_base = smem_u32(smem + C::OFF_DO);
      const uint32_t ds_base = smem_u32(smem + C::OFF_DS);
      auto issue_s_dp = [&](int i) {
        const uint32_t qs = q_base + (i & 1) * C::TILE_BYTES;
        const uint32_t dos = do_base + (i & 1) * C::TILE_BYTES;
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk)
          umma_bf16_ss(tmem_base + C::TM_S, umma_smem_desc(k_base + kk * 32, 0, C::GROUP_BYTES, C::SWZ),
                       umma_smem_desc(qs + kk * 32, 0, C::GROUP_BYTES, C::SWZ), idesc_s, kk > 0 ? 1u : 0u);
        umma_commit(s_full);
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk)
          umma_bf16_ss(tmem_base + C::TM_DP, umma_smem_desc(v_base + kk * 32, 0, C::GROUP_BYTES, C::SWZ),
                       umma_smem_desc(dos + kk * 32, 0, C::GROUP_BYTES, C::SWZ), idesc_s, kk > 0 ? 1u : 0u);
        umma_commit(dp_full);
      };
      mbar_wait(kv_full, 0);
      mbar_wait(&qdo_full[0], 0);
      tcgen05_fence_after();
      issue_s_dp(0);
      for (int i = 0; i < n_q; ++i) {
        if (i + 1 < n_q) {
          mbar_wait(&qdo_full[(i + 1) & 1], ((i + 1) >> 1) & 1);
          mbar_wait(sdp_free, i & 1);  // both warpgroups hold S^T_i / dP^T_i in registers
          tcgen05_fence_after();
          issue_s_dp(i + 1);
        }
        mbar_wait(ps_ready, i & 1);
        if (i > 0) mbar_wait(dq_free, (i - 1) & 1);
        tcgen05_fence_after();
        const uint32_t qs = q_base + (i & 1) * C::TILE_BYTES;
        const uint32_t dos = do_base + (i & 1) * C::TILE_BYTES;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {  // dV += P^T dO_i        (A = P^T from TMEM, 16 queries = 8 packed columns)
          const uint64_t db = umma_smem_desc(dos + kk * 2 * C::GROUP_BYTES, 0, C::GROUP_BYTES, C::SWZ);
          umma_bf16_ts(tmem_base + C::TM_DV, tmem_base + C::TM_PT + kk * 8, db, idesc_kv, (i > 0 || kk > 0) ? 1u : 0u);
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {  // dK += dS^T Q_i
          const uint64_t db = umma_smem_desc(qs + kk * 2 * C::GROUP_BYTES, 0, C::GROUP_BYTES, C::SWZ);
          if constexpr (C::DS_TMEM) {
            umma_bf16_ts(tmem_base + C::TM_DK, tmem_base + C::TM_DST + kk * 8, db, idesc_kv, (i > 0 || kk > 0) ? 1u : 0u);
          } else {
            const uint64_t da = umma_smem_desc(ds_base + (kk >> 2) * 16384 + (kk & 3) * 32, 0, 1024, UMMA_SW128);
            umma_bf16_ss(tmem_base + C::TM_DK, da, db, idesc_kv, (i > 0 || kk > 0) ? 1u : 0u);
          }
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {  // dQ_i = dS K   (A = dS^T read MN-major: M = queries contiguous)
          const uint64_t da = umma_smem_desc(ds_base + kk * 2048, 16384, 1024, UMMA_SW128);
          const uint64_t db = umma_smem_desc(k_base + kk * 2 * C::GROUP_BYTES, 0, C::GROUP_BYTES, C::SWZ);
          umma_bf16_ss(tmem_base + C::TM_DQ, da, db, idesc_dq, kk > 0 ? 1u : 0u);
        }
        umma_commit(mma2_done);
        umma_commit(&qdo_empty[i & 1]);
      }
    }
    __syncwarp();
   }
  } else {
    // ------------------------------------------------ two compute warpgroups
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int wg = warp >> 2;                 // query-column half handled by this warpgroup
    const int r = (warp & 3) * 32 + lane;     // key row inside the tile == TMEM lane
    const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    float* dq_g = p.dq_acc + ((long long)b * p.H + h) * p.Nq * D;

    // dQ tile i: TMEM lanes = query rows, this warpgroup takes D/2 columns.  The accumulator chunk (thread == row, 16
    // fp32) goes through a per-warp swizzled smem tile so that each fp32 reduction instruction covers 8 rows x 64
    // contiguous bytes (whole sectors) instead of 32 rows x 16 bytes: the L2 reduction rate, not the tensor pipe,
    // bounds this kernel, and it is counted in sector requests.
    const uint32_t stg = smem_u32(smem + C::OFF_STG + warp * 2048);
    const int sub = lane >> 2, c16 = lane & 3;
    auto drain_dq = [&](int i) {
      const int row0 = i * 128 + (warp & 3) * 32;  // first query row of this warp
#pragma unroll
      for (int c = 0; c < D / 32; ++c) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(lane_addr + C::TM_DQ + wg * (D / 2) + c * 16, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)),
                       "r"(v[4 * j]), "r"(v[4 * j + 1]), "r"(v[4 * j + 2]), "r"(v[4 * j + 3])
                       : "memory");
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int rr = k * 8 + sub;
          float x0, x1, x2, x3;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(x0), "=f"(x1), "=f"(x2), "=f"(x3)
                       : "r"(stg + rr * 64 + ((c16 ^ ((rr >> 1) & 3)) << 4)));
          const int q_row = row0 + rr;
          if (q_row < p.Nq) red_add_v4f(dq_g + (long long)q_row * D + wg * (D / 2) + c * 16 + c16 * 4, x0, x1, x2, x3);
        }
        __syncwarp();
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_free);
    };

    const uint32_t ds_s32 = smem_u32(smem + C::OFF_DS + wg * 16384);
    for (int i = 0; i < n_q; ++i) {
      // -lse*log2e and delta*scale of this tile's 128 queries arrive with the Q / dO stage (bulk copies)
      const uint32_t nl_s32 = smem_u32(vec + (i & 1) * 256) + wg * 256;  // this warpgroup's 64 queries (16 float4)
      const uint32_t ds_v32 = nl_s32 + 512;
      float s[64], dp[64];
      mbar_wait(s_full, i & 1);
      tcgen05_fence_after();
      {
        uint32_t v[2][32];
        tmem_ld_32x32b_x32(lane_addr + C::TM_S + wg * 64, v[0]);
        tmem_ld_32x32b_x32(lane_addr + C::TM_S + wg * 64 + 32, v[1]);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 64; ++k) s[k] = __uint_as_float(v[k >> 5][k & 31]);
      }
      mbar_wait(dp_full, i & 1);
      tcgen05_fence_after();
      {
        uint32_t v[2][32];
        tmem_ld_32x32b_x32(lane_addr + C::TM_DP + wg * 64, v[0]);
        tmem_ld_32x32b_x32(lane_addr + C::TM_DP + wg * 64 + 32, v[1]);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 64; ++k) dp[k] = __uint_as_float(v[k >> 5][k & 31]);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sdp_free);
      mbar_wait(&qdo_full[i & 1], (i >> 1) & 1);  // the vectors travelled with this stage

      uint32_t pk[32], dsk[32];  // packed bf16 P^T and dS^T of this row half
#pragma unroll
      for (int g = 0; g < 16; ++g) {
        float4 nl, dsv;  // explicit shared-memory loads (broadcast): a generic LD here serialises on its latency
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(nl.x), "=f"(nl.y), "=f"(nl.z), "=f"(nl.w) : "r"(nl_s32 + g * 16));
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(dsv.x), "=f"(dsv.y), "=f"(dsv.z), "=f"(dsv.w) : "r"(ds_v32 + g * 16));
        const float p0 = ex2_approx(fmaf(s[4 * g + 0], p.scale_log2, nl.x));
        const float p1 = ex2_approx(fmaf(s[4 * g + 1], p.scale_log2, nl.y));
        const float p2 = ex2_approx(fmaf(s[4 * g + 2], p.scale_log2, nl.z));
        const float p3 = ex2_approx(fmaf(s[4 * g + 3], p.scale_log2, nl.w));
        pk[2 * g] = pack_bf16(p0, p1);
        pk[2 * g + 1] = pack_bf16(p2, p3);
        dsk[2 * g] = pack_bf16(p0 * fmaf(dp[4 * g + 0], p.scale, -dsv.x), p1 * fmaf(dp[4 * g + 1], p.scale, -dsv.y));
        dsk[2 * g + 1] = pack_bf16(p2 * fmaf(dp[4 * g + 2], p.scale, -dsv.z), p3 * fmaf(dp[4 * g + 3], p.scale, -dsv.w));
      }
      if (i > 0) {
        mbar_wait(mma2_done, (i - 1) & 1);  // P / dS smem reusable, dQ_{i-1} complete
        tcgen05_fence_after();
        drain_dq(i - 1);
      }
      // P^T (and dS^T for head_dim 32) -> TMEM as packed bf16: this warpgroup's 64 queries = 32 columns
      tmem_st_32x32b_x32(lane_addr + C::TM_PT + wg * 32, pk);
      if constexpr (C::DS_TMEM) tmem_st_32x32b_x32(lane_addr + C::TM_DST + wg * 32, dsk);
#pragma unroll
      for (int vcol = 0; vcol < 8; ++vcol) {  // dS^T -> smem: dQ = dS K reads it transposed (MN-major A)
        const uint32_t off = sw128_vec_offset(r, vcol);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ds_s32 + off), "r"(dsk[4 * vcol]),
                     "r"(dsk[4 * vcol + 1]), "r"(dsk[4 * vcol + 2]), "r"(dsk[4 * vcol + 3])
                     : "memory");
      }
      tmem_st_wait();
      fence_proxy_async_smem();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ps_ready);
    }
    mbar_wait(mma2_done, (n_q - 1) & 1);
    tcgen05_fence_after();
    drain_dq(n_q - 1);

    // dK / dV epilogue: lanes = key rows, this warpgroup takes D/2 columns of each
    const int kv_row = kv0 + r;
    bf16* dk_ptr = p.dk + (long long)b * p.dk_sb + (long long)kv_row * p.dk_sn + (long long)h * p.dk_sh + wg * (D / 2);
    bf16* dv_ptr = p.dv + (long long)b * p.dv_sb + (long long)kv_row * p.dv_sn + (long long)h * p.dv_sh + wg * (D / 2);
#pragma unroll
    for (int c = 0; c < D / 32; ++c) {
      uint32_t a[16], v[16];
      tmem_ld_32x32b_x16(lane_addr + C::TM_DK + wg * (D / 2) + c * 16, a);
      tmem_ld_32x32b_x16(lane_addr + C::TM_DV + wg * (D / 2) + c * 16, v);
      tmem_ld_wait();
      if (kv_row < p.Nk) {
        uint4 w0 = make_uint4(pack_bf16(__uint_as_float(a[0]), __uint_as_float(a[1])), pack_bf16(__uint_as_float(a[2]), __uint_as_float(a[3])),
                              pack_bf16(__uint_as_float(a[4]), __uint_as_float(a[5])), pack_bf16(__uint_as_float(a[6]), __uint_as_float(a[7])));
        uint4 w1 = make_uint4(pack_bf16(__uint_as_float(a[8]), __uint_as_float(a[9])), pack_bf16(__uint_as_float(a[10]), __uint_as_float(a[11])),
                              pack_bf16(__uint_as_float(a[12]), __uint_as_float(a[13])), pack_bf16(__uint_as_float(a[14]), __uint_as_float(a[15])));
        reinterpret_cast<uint4*>(dk_ptr + c * 16)[0] = w0;
        reinterpret_cast<uint4*>(dk_ptr + c * 16)[1] = w1;
        uint4 x0 = make_uint4(pack_bf16(__uint_as_float(v[0]), __uint_as_float(v[1])), pack_bf16(__uint_as_float(v[2]), __uint_as_float(v[3])),
                              pack_bf16(__uint_as_float(v[4]), __uint_as_float(v[5])), pack_bf16(__uint_as_float(v[6]), __uint_as_float(v[7])));
        uint4 x1 = make_uint4(pack_bf16(__uint_as_float(v[8]), __uint_as_float(v[9])), pack_bf16(__uint_as_float(v[10]), __uint_as_float(v[11])),
                              pack_bf16(__uint_as_float(v[12]), __uint_as_float(v[13])), pack_bf16(__uint_as_float(v[14]), __uint_as_float(v[15])));
        reinterpret_cast<uint4*>(dv_ptr + c * 16)[0] = x0;
        reinterpret_cast<uint4*>(dv_ptr + c * 16)[1] = x1;
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 9) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// Backward, second generation (default).  Same tiling and MMAs as above (one CTA per 128-key tile, loop over 128-query
// tiles, lanes = keys), re-plumbed so that nothing on the SM waits in lock step (profiles/r01_attention_clock_trace.md:
// the first generation spent 3.8 k clocks per query tile in ONE serial chain of its eight compute warps -- TMEM loads,
// 64 exponentials per thread, the dQ drain, TMEM / smem stores -- with the tensor pipe idle for most of it):
//   * 512 threads: warpgroups 0 / 1 own the two 64-query HALVES of every tile with their own barriers
//     (dp_full / sdp_free / ps_ready / pdk_done per half), so they drift apart and one half's exp2 phase runs under the
//     other's TMEM traffic and barrier waits; S^T / dP^T (N = 64) and dV / dK are issued per half.  Each half is
//     processed in two 32-query chunks (64 accumulator registers in flight, no spills) and its dS^T vectors are
//     stored to shared memory as they are produced, so the generic->async proxy fence finds nothing in flight.
//   * warpgroup 2 drains dQ (TMEM -> swizzled smem transposition -> coalesced fp32 red.global.add) off the critical
//     path; at head_dim 32 the dQ accumulator is double buffered in TMEM, so dQ(i+1) is issued while dQ(i) drains.
//   * dS^T in shared memory (the MN-major A operand of dQ = dS K) is double buffered by tile parity.
//   * warpgroup 3: TMA producer warp (three Q / dO stages) and TWO MMA issuer warps -- one issues S^T / dP^T of the
//     next tile the moment a half's registers are loaded, the other dV / dK per half and dQ per tile -- each running
//     converged with one elected lane, so that the tcgen05 instructions go out back to back (the single `lane == 0`
//     issuer of the first cut needed ~3 k clocks per query tile for its 32 MMAs + 9 commits and sat on the critical
//     path together with the exposed TMA latency of a two-stage ring).  Registers: 2 x 192 + 64 + 64 (setmaxnreg).
//   * the second half of a last query tile that holds no queries (Nq % 128 in 1..64) skips its arithmetic.
// TMEM: S^T (2 x 64) | dP^T (2 x 64) | P^T (2 x 32 packed) | [dS^T (2 x 32 packed)] | dV (D) | dK (D) | dQ (D) [| dQ' (D)].
// ---------------------------------------------------------------------------------------------
constexpr int BWD2_THREADS = 512;

template <int D>
struct Bwd2Cfg {
  static constexpr int ROW_BYTES = D * 2;
  static constexpr uint64_t SWZ = D == 64 ? UMMA_SW128 : UMMA_SW64;
  static constexpr int GROUP_BYTES = 8 * ROW_BYTES;
  static constexpr int TILE_BYTES = 128 * ROW_BYTES;
  static constexpr int PS_BYTES = 128 * 128 * 2;  // one dS^T buffer: two 64-query chunks of 16 KB
  static constexpr int OFF_K = 0;
  static constexpr int OFF_V = OFF_K + TILE_BYTES;
  static constexpr int STAGES = 3;                           // Q / dO / per-query vector stages
  static constexpr int OFF_Q = OFF_V + TILE_BYTES;
  static constexpr int OFF_DO = OFF_Q + STAGES * TILE_BYTES;
  static constexpr int OFF_DS = OFF_DO + STAGES * TILE_BYTES;  // 2 buffers (tile parity)
  static constexpr int OFF_VEC = OFF_DS + 2 * PS_BYTES;        // lse / delta: [STAGES][2][128] floats
  static constexpr int OFF_STG = (OFF_VEC + STAGES * 2 * 128 * 4 + 1023) / 1024 * 1024;  // dQ drain: 4 warps x (32 rows x 128 B)
  static constexpr int OFF_BAR = OFF_STG + 4 * 4096;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
  static constexpr bool DS_TMEM = D <= 32;  // dS^T as a TMEM A operand for dK (an smem A operand paces N = 32 MMAs at ~64 clk)
  static constexpr bool DQ2 = D <= 32;      // two dQ accumulators
  static constexpr int TM_S = 0, TM_DP = 128, TM_PT = 256, TM_DST = 320;
  static constexpr int TM_DV = DS_TMEM ? 384 : 320, TM_DK = TM_DV + D, TM_DQ = TM_DV + 2 * D;
  static_assert(TM_DQ + (DQ2 ? 2 : 1) * D <= 512, "TMEM budget");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

template <int D>
__global__ void __launch_bounds__(BWD2_THREADS, 1)
attn_bwd2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                 const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                 const __grid_constant__ CUtensorMap tm_dq, const AttnBwdArgs p) {
  using C = Bwd2Cfg<D>;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* kv_full = bars + 0;
  uint64_t* qdo_full = bars + 1;    // [3] stages
  uint64_t* qdo_empty = bars + 4;   // [3]
  uint64_t* s_full = bars + 7;      // [2] halves
  uint64_t* dp_full = bars + 9;     // [2]
  uint64_t* sdp_free = bars + 11;   // [2], 4 warp arrivals each
  uint64_t* ps_ready = bars + 13;   // [2], 4 warp arrivals each
  uint64_t* pdk_done = bars + 15;   // [2]: dV / dK MMAs of this half retired -> its P^T / dS^T TMEM columns are free
  uint64_t* ds_free = bars + 17;    // [2] smem dS^T buffers (tile parity): dQ MMAs of that tile retired
  uint64_t* dq_full = bars + 19;    // [2] dQ accumulators
  uint64_t* dq_free = bars + 21;    // [2], 4 warp arrivals each
  uint64_t* all_done = bars + 23;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
  float* vec = reinterpret_cast<float*>(smem + C::OFF_VEC);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kv0 = blockIdx.x * 128;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int n_q = (p.Nq + 127) / 128;
  const int last_halves = (p.Nq - (n_q - 1) * 128 + 63) >> 6;  // halves of the last query tile that hold queries (1 or 2)
  long long* const trace = g_attn_trace;
  const bool tron = trace != nullptr && lane == 0 && blockIdx.x == 2 && blockIdx.y == 3 && blockIdx.z == gridDim.z / 2;
  if (warp == 0) CB_TR(1000);

  if (warp == 13) {
    if (lane == 0) {
      mbar_init(kv_full, 1), mbar_init(all_done, 1);
      for (int i = 0; i < C::STAGES; ++i) mbar_init(&qdo_full[i], 1), mbar_init(&qdo_empty[i], 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&s_full[i], 1), mbar_init(&dp_full[i], 1), mbar_init(&sdp_free[i], 4), mbar_init(&ps_ready[i], 4);
        mbar_init(&pdk_done[i], 1), mbar_init(&ds_free[i], 1), mbar_init(&dq_full[i], 1), mbar_init(&dq_free[i], 4);
      }
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
  }
  if (warp == 12 && lane == 0) {
    tma_prefetch_desc(&tm_q), tma_prefetch_desc(&tm_k), tma_prefetch_desc(&tm_v), tma_prefetch_desc(&tm_do);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // PDL: setup done, wait for the kernels that produce Q / K / V / dO before the first load
  if (warp == 0) CB_TR(1001);

  if (warp >= 12) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");  // 128 x (2 x 192 + 64 + 64) == 64 K registers
    if (warp == 12) {
      // ------------------------------------------------ TMA producer
      if (lane == 0) {
        mbar_expect_tx(kv_full, 2 * C::TILE_BYTES);
        tma_load_4d(smem + C::OFF_K, &tm_k, kv_full, 0, h, kv0, b);
        tma_load_4d(smem + C::OFF_V, &tm_v, kv_full, 0, h, kv0, b);
        for (int i = 0; i < n_q; ++i) {
          const int s = i % C::STAGES;
          if (i >= C::STAGES) mbar_wait(&qdo_empty[s], ((i / C::STAGES) & 1) ^ 1);  // (the first fills find the ring empty)
          mbar_expect_tx(&qdo_full[s], 2 * C::TILE_BYTES + 1024);
          tma_load_4d(smem + C::OFF_Q + s * C::TILE_BYTES, &tm_q, &qdo_full[s], 0, h, i * 128, b);
          tma_load_4d(smem + C::OFF_DO + s * C::TILE_BYTES, &tm_do, &qdo_full[s], 0, h, i * 128, b);
          const long long voff = ((long long)b * p.H + h) * p.NqP + i * 128;  // per-query vectors of this tile
          bulk_load_1d(vec + s * 256, p.nl2 + voff, 512, &qdo_full[s]);
          bulk_load_1d(vec + s * 256 + 128, p.dsc + voff, 512, &qdo_full[s]);
        }
      }
      __syncwarp();
    } else if (warp == 13) {
      // ------------------------------------------------ MMA issuer A: S^T / dP^T of every half, as early as possible
      // (whole warp converged, one elected lane issues: see attn_fwd2_kernel)
      const bool leader = elect_one_sync() != 0;
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, false, false);  // S^T / dP^T of one half
      const uint32_t k_base = smem_u32(smem + C::OFF_K);
      const uint32_t v_base = smem_u32(smem + C::OFF_V);
      const uint32_t q_base = smem_u32(smem + C::OFF_Q);
      const uint32_t do_base = smem_u32(smem + C::OFF_DO);
      mbar_wait(kv_full, 0);
      for (int i = 0; i < n_q; ++i) {
        const int st = i % C::STAGES;
        mbar_wait(&qdo_full[st], (i / C::STAGES) & 1);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          if (i > 0) mbar_wait(&sdp_free[hf], (i - 1) & 1);  // warpgroup hf holds S^T / dP^T of tile i - 1 in registers
          tcgen05_fence_after();
          if (leader) {  // queries [64 hf, 64 hf + 64) of tile i: rows of the Q / dO tiles
            const uint32_t qs = q_base + st * C::TILE_BYTES + hf * 8 * C::GROUP_BYTES;
            const uint32_t dos = do_base + st * C::TILE_BYTES + hf * 8 * C::GROUP_BYTES;
#pragma unroll
            for (int kk = 0; kk < D / 16; ++kk)
              umma_bf16_ss(tmem_base + C::TM_S + hf * 64, umma_smem_desc(k_base + kk * 32, 0, C::GROUP_BYTES, C::SWZ),
                           umma_smem_desc(qs + kk * 32, 0, C::GROUP_BYTES, C::SWZ), idesc_s, kk > 0 ? 1u : 0u);
#pragma unroll
            for (int kk = 0; kk < D / 16; ++kk)
              umma_bf16_ss(tmem_base + C::TM_DP + hf * 64, umma_smem_desc(v_base + kk * 32, 0, C::GROUP_BYTES, C::SWZ),
                           umma_smem_desc(dos + kk * 32, 0, C::GROUP_BYTES, C::SWZ), idesc_s, kk > 0 ? 1u : 0u);
            umma_commit(&dp_full[hf]);  // one arrival for S^T and dP^T of this half
          }
          __syncwarp();
        }
        if (i < 24) CB_TR(32 * i + 16);
      }
    } else if (warp == 14) {
      // ------------------------------------------------ MMA issuer B: dV / dK per half, then dQ of the tile
      const bool leader = elect_one_sync() != 0;
      constexpr uint32_t idesc_kv = umma_idesc_bf16(128, D, false, true);   // dV, dK
      constexpr uint32_t idesc_dq = umma_idesc_bf16(128, D, true, true);    // dQ
      const uint32_t k_base = smem_u32(smem + C::OFF_K);
      const uint32_t q_base = smem_u32(smem + C::OFF_Q);
      const uint32_t do_base = smem_u32(smem + C::OFF_DO);
      const uint32_t ds_base0 = smem_u32(smem + C::OFF_DS);
      mbar_wait(kv_full, 0);
      for (int i = 0; i < n_q; ++i) {
        const int st = i % C::STAGES;
        mbar_wait(&qdo_full[st], (i / C::STAGES) & 1);  // (long complete: the compute warps needed it; orders the TMA writes)
        const uint32_t qs = q_base + st * C::TILE_BYTES;
        const uint32_t dos = do_base + st * C::TILE_BYTES;
        const uint32_t ds_base = ds_base0 + (i & 1) * C::PS_BYTES;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          mbar_wait(&ps_ready[hf], i & 1);
          if (i < 24) CB_TR(32 * i + 17 + 2 * hf);
          tcgen05_fence_after();
          if (leader) {
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {  // dV += P^T dO_i   (A = P^T from TMEM: 16 queries = 8 packed columns per step)
              const int kk = hf * 4 + k4;
              const uint64_t db = umma_smem_desc(dos + kk * 2 * C::GROUP_BYTES, 0, C::GROUP_BYTES, C::SWZ);
              umma_bf16_ts(tmem_base + C::TM_DV, tmem_base + C::TM_PT + kk * 8, db, idesc_kv, (i > 0 || kk > 0) ? 1u : 0u);
            }
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {  // dK += dS^T Q_i
              const int kk = hf * 4 + k4;
              const uint64_t db = umma_smem_desc(qs + kk * 2 * C::GROUP_BYTES, 0, C::GROUP_BYTES, C::SWZ);
              if constexpr (C::DS_TMEM) {
                umma_bf16_ts(tmem_base + C::TM_DK, tmem_base + C::TM_DST + kk * 8, db, idesc_kv, (i > 0 || kk > 0) ? 1u : 0u);
              } else {
                const uint64_t da = umma_smem_desc(ds_base + (kk >> 2) * 16384 + (kk & 3) * 32, 0, 1024, UMMA_SW128);
                umma_bf16_ss(tmem_base + C::TM_DK, da, db, idesc_kv, (i > 0 || kk > 0) ? 1u : 0u);
              }
            }
            umma_commit(&pdk_done[hf]);
            if (hf == 1) umma_commit(&qdo_empty[st]);  // Q_i / dO_i (and the vectors) have no reader left
          }
          __syncwarp();
          if (i < 24) CB_TR(32 * i + 18 + 2 * hf);
        }
        // dQ_i = dS K   (A = dS^T of both halves read MN-major from smem: M = queries contiguous)
        int buf = 0;
        if constexpr (C::DQ2) {
          buf = i & 1;
          if (i >= 2) mbar_wait(&dq_free[buf], ((i >> 1) & 1) ^ 1);  // tile i - 2 drained from this accumulator
        } else {
          if (i >= 1) mbar_wait(&dq_free[0], (i - 1) & 1);
        }
        if (i < 24) CB_TR(32 * i + 21);
        tcgen05_fence_after();
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint64_t da = umma_smem_desc(ds_base + kk * 2048, 16384, 1024, UMMA_SW128);
            const uint64_t db = umma_smem_desc(k_base + kk * 2 * C::GROUP_BYTES, 0, C::GROUP_BYTES, C::SWZ);
            umma_bf16_ss(tmem_base + C::TM_DQ + buf * D, da, db, idesc_dq, kk > 0 ? 1u : 0u);
          }
          umma_commit(&dq_full[buf]);
          umma_commit(&ds_free[i & 1]);
          if (i == n_q - 1) umma_commit(all_done);
        }
        __syncwarp();
        if (i < 24) CB_TR(32 * i + 22);
      }
    }
  } else if (warp >= 8) {
    // ------------------------------------------------ dQ drain warpgroup: lanes = query rows, D columns per warp
    // TMEM -> registers -> a 128B-swizzled [32 rows x 32 fp32] staging tile per warp -> ONE bulk-tensor reduce-add per
    // 32 columns (cp.reduce.async.bulk.tensor, fp32 add in L2).  The first cut drained with red.global.add.v4 from the
    // LSU, which retires ~12 B per clock and SM (2.6 k clocks per 128 x 64 tile: it set the pace of the head_dim 64
    // kernel, whose dQ accumulator cannot be double buffered); the TMA unit takes the tile off the SM in one
    // instruction, and the 3-D tensor map (dim, query, batch * head) clips the ragged last query tile by itself.
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    const int w4 = warp & 3;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(w4 * 32) << 16);
    uint8_t* stg = smem + C::OFF_STG + w4 * 4096;
    const uint32_t stg32 = smem_u32(stg);
    const int bh = b * p.H + h;
    if (warp == 8 && lane == 0) tma_prefetch_desc(&tm_dq);
    for (int i = 0; i < n_q; ++i) {
      int buf = 0;
      uint32_t ph = i & 1;
      if constexpr (C::DQ2) buf = i & 1, ph = (i >> 1) & 1;
      mbar_wait(&dq_full[buf], ph);
      if (warp == 8 && i < 24) CB_TR(32 * i + 24);
      tcgen05_fence_after();
      const int row0 = i * 128 + w4 * 32;  // first query row of this warp
      if (row0 < p.Nq) {
#pragma unroll
        for (int half = 0; half < D / 32; ++half) {
          if (lane == 0) bulk_wait_group_read0();  // the previous reduce has read the staging tile
          __syncwarp();
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t v[16];
            tmem_ld_32x32b_x16(lane_addr + C::TM_DQ + buf * D + half * 32 + c * 16, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 4; ++j)  // row = lane, 16-byte chunk (c * 4 + j) of the 128-byte row, 128B swizzle
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg32 + lane * 128 + (((c * 4 + j) ^ (lane & 7)) << 4)),
                           "r"(v[4 * j]), "r"(v[4 * j + 1]), "r"(v[4 * j + 2]), "r"(v[4 * j + 3])
                           : "memory");
          }
          if (half == D / 32 - 1) {  // the accumulator is in registers / shared memory: hand it back to the MMA warp
            tcgen05_fence_before();
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (half == D / 32 - 1) mbar_arrive(&dq_free[buf]);
            tma_reduce_add_3d(&tm_dq, stg, half * 32, row0, bh);
            bulk_commit_group();
          }
        }
      } else {
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&dq_free[buf]);
      }
      if (warp == 8 && i < 24) CB_TR(32 * i + 25);
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // reductions complete before the CTA retires
  } else {
    // ------------------------------------------------ two compute warpgroups: half hf = 64 query columns of every tile
    asm volatile("setmaxnreg.inc.sync.aligned.u32 192;");
    const int hf = warp >> 2;
    const int r = (warp & 3) * 32 + lane;  // key row inside the tile == TMEM lane
    const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    // The exp2-heavy section is handed back and forth between the two warpgroups with named barriers 1 / 2 (256 threads:
    // one warpgroup syncs, the other arrives), half 0 first.  Left alone the halves run in lock step -- both in the
    // MUFU-bound section (1.6 k clocks for the pair), then both in barrier waits / TMEM loads / stores with the MUFU pipe
    // idle (0.85 k); with the token they settle half an iteration apart and one half's arithmetic runs under the
    // other's memory phases.  CB_ATTN_BWD_TOKEN=0 (AttnBwdArgs::token) switches it off for A/B measurements.
    const bool token = p.token != 0;
    if (token && hf == 1) asm volatile("bar.arrive 1, 256;" ::: "memory");
    for (int i = 0; i < n_q; ++i) {
      // the second half of a ragged last tile may hold no query at all: P^T = dS^T = 0 there by construction (padded
      // vectors), so the TMEM loads and the arithmetic are skipped and zeros are stored
      const bool live = !(i == n_q - 1 && hf >= last_halves);
      const uint32_t nl_s32 = smem_u32(vec + (i % C::STAGES) * 256) + hf * 256;  // this half's 64 queries (16 float4)
      const uint32_t ds_v32 = nl_s32 + 512;
      const uint32_t ds_s32 = smem_u32(smem + C::OFF_DS + (i & 1) * C::PS_BYTES + hf * 16384);
      const bool tri = (warp & 3) == 0 && i < 24;
      const int tb = 32 * i + 8 * hf;
      if (tri) CB_TR(tb + 0);
      mbar_wait(&dp_full[hf], i & 1);                                // S^T and dP^T of this half (one commit covers both)
      mbar_wait(&qdo_full[i % C::STAGES], (i / C::STAGES) & 1);      // the per-query vectors travelled with this stage
      if (i >= 2) mbar_wait(&ds_free[i & 1], ((i >> 1) & 1) ^ 1);    // dQ (dK) of tile i - 2 have read this smem buffer
      if (tri) CB_TR(tb + 1);
      tcgen05_fence_after();
      uint32_t pk[32], dsk[32];  // packed bf16 P^T and dS^T of this half
      if (live) {
        // two chunks of 32 queries: 64 accumulator registers in flight instead of 128 (no spills), and the dS^T
        // vectors go to shared memory as they are produced, far ahead of the proxy fence below
        uint32_t vs_all[2][32], vd_all[2][32];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          tmem_ld_32x32b_x32(lane_addr + C::TM_S + hf * 64 + c * 32, vs_all[c]);
          tmem_ld_32x32b_x32(lane_addr + C::TM_DP + hf * 64 + c * 32, vd_all[c]);
        }
        tmem_ld_wait();
        // S^T / dP^T of this half are in registers: issuer A may overwrite them with tile i + 1 right away
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sdp_free[hf]);
        if (tri) CB_TR(tb + 3);
        if (token) {
          if (hf == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
          else asm volatile("bar.sync 2, 256;" ::: "memory");
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const uint32_t* vs = vs_all[c];
          const uint32_t* vd = vd_all[c];
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            float4 nl, dsv;  // explicit shared-memory loads (broadcast): a generic LD here serialises on its latency
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(nl.x), "=f"(nl.y), "=f"(nl.z), "=f"(nl.w) : "r"(nl_s32 + (c * 8 + g) * 16));
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(dsv.x), "=f"(dsv.y), "=f"(dsv.z), "=f"(dsv.w) : "r"(ds_v32 + (c * 8 + g) * 16));
            const float p0 = ex2_approx(fmaf(__uint_as_float(vs[4 * g + 0]), p.scale_log2, nl.x));
            const float p1 = ex2_approx(fmaf(__uint_as_float(vs[4 * g + 1]), p.scale_log2, nl.y));
            const float p2 = ex2_approx(fmaf(__uint_as_float(vs[4 * g + 2]), p.scale_log2, nl.z));
            const float p3 = ex2_approx(fmaf(__uint_as_float(vs[4 * g + 3]), p.scale_log2, nl.w));
            const int o = c * 16 + 2 * g;
            pk[o] = pack_bf16(p0, p1);
            pk[o + 1] = pack_bf16(p2, p3);
            dsk[o] = pack_bf16(p0 * fmaf(__uint_as_float(vd[4 * g + 0]), p.scale, -dsv.x),
                               p1 * fmaf(__uint_as_float(vd[4 * g + 1]), p.scale, -dsv.y));
            dsk[o + 1] = pack_bf16(p2 * fmaf(__uint_as_float(vd[4 * g + 2]), p.scale, -dsv.z),
                                   p3 * fmaf(__uint_as_float(vd[4 * g + 3]), p.scale, -dsv.w));
          }
#pragma unroll
          for (int v4 = 0; v4 < 4; ++v4) {  // dS^T -> smem: dQ = dS K reads it transposed (MN-major A)
            const int vcol = c * 4 + v4;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ds_s32 + sw128_vec_offset(r, vcol)), "r"(dsk[4 * vcol]),
                         "r"(dsk[4 * vcol + 1]), "r"(dsk[4 * vcol + 2]), "r"(dsk[4 * vcol + 3])
                         : "memory");
          }
        }
        if (token) {  // math done: pass the token (the very last pass of half 1 has no taker)
          if (hf == 0) asm volatile("bar.arrive 2, 256;" ::: "memory");
          else if (i + 1 < n_q) asm volatile("bar.arrive 1, 256;" ::: "memory");
        }
      } else {
#pragma unroll
        for (int g = 0; g < 32; ++g) pk[g] = 0u, dsk[g] = 0u;
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sdp_free[hf]);
        if (token) {  // keep the token protocol balanced on the dead half: take it and pass it on
          if (hf == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
          else asm volatile("bar.sync 2, 256;" ::: "memory");
          if (hf == 0) asm volatile("bar.arrive 2, 256;" ::: "memory");
          else if (i + 1 < n_q) asm volatile("bar.arrive 1, 256;" ::: "memory");
        }
#pragma unroll
        for (int vcol = 0; vcol < 8; ++vcol)
          asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(ds_s32 + sw128_vec_offset(r, vcol)), "r"(0u) : "memory");
      }
      if (tri) CB_TR(tb + 4);
      if (i > 0) mbar_wait(&pdk_done[hf], (i - 1) & 1);  // this half's P^T / dS^T TMEM columns are free
      if (tri) CB_TR(tb + 5);
      tcgen05_fence_after();
      // P^T (and dS^T for head_dim 32) -> TMEM as packed bf16: this half's 64 queries = 32 columns
      tmem_st_32x32b_x32(lane_addr + C::TM_PT + hf * 32, pk);
      if constexpr (C::DS_TMEM) tmem_st_32x32b_x32(lane_addr + C::TM_DST + hf * 32, dsk);
      tmem_st_wait();
      fence_proxy_async_smem();
      if (tri) CB_TR(tb + 6);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ps_ready[hf]);
    }
    if (warp == 0) CB_TR(1002);
    mbar_wait(all_done, 0);
    tcgen05_fence_after();

    // dK / dV epilogue: lanes = key rows, this warpgroup takes D/2 columns of each
    const int kv_row = kv0 + r;
    bf16* dk_ptr = p.dk + (long long)b * p.dk_sb + (long long)kv_row * p.dk_sn + (long long)h * p.dk_sh + hf * (D / 2);
    bf16* dv_ptr = p.dv + (long long)b * p.dv_sb + (long long)kv_row * p.dv_sn + (long long)h * p.dv_sh + hf * (D / 2);
#pragma unroll
    for (int c = 0; c < D / 32; ++c) {
      uint32_t a[16], v[16];
      tmem_ld_32x32b_x16(lane_addr + C::TM_DK + hf * (D / 2) + c * 16, a);
      tmem_ld_32x32b_x16(lane_addr + C::TM_DV + hf * (D / 2) + c * 16, v);
      tmem_ld_wait();
      if (kv_row < p.Nk) {
        uint4 w0 = make_uint4(pack_bf16(__uint_as_float(a[0]), __uint_as_float(a[1])), pack_bf16(__uint_as_float(a[2]), __uint_as_float(a[3])),
                              pack_bf16(__uint_as_float(a[4]), __uint_as_float(a[5])), pack_bf16(__uint_as_float(a[6]), __uint_as_float(a[7])));
        uint4 w1 = make_uint4(pack_bf16(__uint_as_float(a[8]), __uint_as_float(a[9])), pack_bf16(__uint_as_float(a[10]), __uint_as_float(a[11])),
                              pack_bf16(__uint_as_float(a[12]), __uint_as_float(a[13])), pack_bf16(__uint_as_float(a[14]), __uint_as_float(a[15])));
        reinterpret_cast<uint4*>(dk_ptr + c * 16)[0] = w0;
        reinterpret_cast<uint4*>(dk_ptr + c * 16)[1] = w1;
        uint4 x0 = make_uint4(pack_bf16(__uint_as_float(v[0]), __uint_as_float(v[1])), pack_bf16(__uint_as_float(v[2]), __uint_as_float(v[3])),
                              pack_bf16(__uint_as_float(v[4]), __uint_as_float(v[5])), pack_bf16(__uint_as_float(v[6]), __uint_as_float(v[7])));
        uint4 x1 = make_uint4(pack_bf16(__uint_as_float(v[8]), __uint_as_float(v[9])), pack_bf16(__uint_as_float(v[10]), __uint_as_float(v[11])),
                              pack_bf16(__uint_as_float(v[12]), __uint_as_float(v[13])), pack_bf16(__uint_as_float(v[14]), __uint_as_float(v[15])));
        reinterpret_cast<uint4*>(dv_ptr + c * 16)[0] = x0;
        reinterpret_cast<uint4*>(dv_ptr + c * 16)[1] = x1;
      }
      if (p.dk_cs != nullptr || p.dv_cs != nullptr) {
        // bias gradients of the k / v projections: column sums of the ROUNDED values just stored, over this warp's 32
        // key rows (shuffle tree), one fp32 atomic per column and warp
        const bool ok = kv_row < p.Nk;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float sk = ok ? bf16_round(__uint_as_float(a[j])) : 0.f;
          float sv = ok ? bf16_round(__uint_as_float(v[j])) : 0.f;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            sk += __shfl_xor_sync(0xffffffffu, sk, o);
            sv += __shfl_xor_sync(0xffffffffu, sv, o);
          }
          if (lane == j) {  // spread the 16 atomics of a chunk over 16 lanes
            const int col = h * D + hf * (D / 2) + c * 16 + j;
            if (p.dk_cs != nullptr) atomicAdd(p.dk_cs + col, sk);
            if (p.dv_cs != nullptr) atomicAdd(p.dv_cs + col, sv);
          }
        }
      }
    }
    if (warp == 0) CB_TR(1003);
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 13) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// Per-query vectors of the backward, D / 8 lanes per (b, q, h) row, 16-byte loads:
//   dsc[bh, q] = scale * sum_d O[b,q,h,d] * dO[b,q,h,d]      nl2[bh, q] = -lse[b,h,q] * log2(e)
// rows are padded to NqP = 128 * ceil(Nq / 128) entries (nl2 = -inf -> P = 0, dsc = 0) so that the main kernel can
// fetch whole 128-query vectors with one bulk copy each.  Threads walk the rows in MEMORY order (batch, query, head):
// the heads of one query are adjacent in the projection output, so a warp reads 512 contiguous bytes of O and of dO
// (the first cut walked (batch, head, query) and touched 64 useful bytes per kilobyte: 2.0 TB/s); the two small outputs
// are written transposed.
template <int D>
__global__ void attn_delta_kernel(const bf16* __restrict__ o, long long o_sb, long long o_sn, long long o_sh,
                                  const bf16* __restrict__ d_o, long long do_sb, long long do_sn, long long do_sh,
                                  const float* __restrict__ lse, float* __restrict__ nl2, float* __restrict__ dsc, int B,
                                  int H, int Nq, int NqP, float scale) {
  pdl_prologue();  // PDL: release the next launch, then wait for the previous kernel's results
  constexpr int LPR = D / 8;  // lanes per row
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long row = gid / LPR;  // over B * NqP * H, head fastest
  const int sub = (int)(gid % LPR);
  const long long total = (long long)B * H * NqP;
  float acc = 0.f;
  const int h = (int)(row % H);
  const int q = (int)((row / H) % NqP);
  const int b = (int)(row / ((long long)H * NqP));
  const bool live = row < total && q < Nq;
  if (live) {
    const bf16* op = o + b * o_sb + (long long)q * o_sn + (long long)h * o_sh + sub * 8;
    const bf16* dp = d_o + b * do_sb + (long long)q * do_sn + (long long)h * do_sh + sub * 8;
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(op));
    const uint4 g = __ldg(reinterpret_cast<const uint4*>(dp));
    const uint32_t au[4] = {a.x, a.y, a.z, a.w}, gu[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 x = unpack_bf16(au[k]), y = unpack_bf16(gu[k]);
      acc += x.x * y.x + x.y * y.y;
    }
  }
#pragma unroll
  for (int m = 1; m < LPR; m <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
  if (row < total && sub == 0) {
    const long long out = ((long long)b * H + h) * NqP + q;
    dsc[out] = live ? acc * scale : 0.f;
    nl2[out] = live ? -lse[((long long)b * H + h) * Nq + q] * LOG2E : -INFINITY;
  }
}

// dq[b,q,h,:] = bf16(dq_acc[b,h,q,:]); optionally colsum[h * D + c] += sum over (b, q) of the rounded values (the q
// projection's bias gradient).  Block = one (b, h) and a strided set of its query rows; D / 8 threads per row.
template <int D>
__global__ void __launch_bounds__(256)
attn_dq_convert_kernel(const float* __restrict__ acc, bf16* __restrict__ dq, long long sb, long long sn, long long sh, int B,
                       int H, int Nq, float* __restrict__ colsum) {
  pdl_prologue();  // PDL: release the next launch, then wait for the previous kernel's results
  constexpr int VPR = D / 8;       // 16-byte output vectors per row
  constexpr int RPB = 256 / VPR;   // rows per block pass
  const int vcol = threadIdx.x % VPR, rsub = threadIdx.x / VPR;
  const int h = blockIdx.y, b = blockIdx.z;
  const float* a = acc + ((long long)b * H + h) * Nq * D;
  bf16* out = dq + b * sb + (long long)h * sh + vcol * 8;
  float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int q = blockIdx.x * RPB + rsub; q < Nq; q += gridDim.x * RPB) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(a + (long long)q * D + vcol * 8));
    const float4 y = __ldg(reinterpret_cast<const float4*>(a + (long long)q * D + vcol * 8 + 4));
    const uint4 pk = make_uint4(pack_bf16(x.x, x.y), pack_bf16(x.z, x.w), pack_bf16(y.x, y.y), pack_bf16(y.z, y.w));
    *reinterpret_cast<uint4*>(out + (long long)q * sn) = pk;
    if (colsum != nullptr) {
      const float2 p0 = unpack_bf16(pk.x), p1 = unpack_bf16(pk.y), p2 = unpack_bf16(pk.z), p3 = unpack_bf16(pk.w);
      cs[0] += p0.x, cs[1] += p0.y, cs[2] += p1.x, cs[3] += p1.y, cs[4] += p2.x, cs[5] += p2.y, cs[6] += p3.x, cs[7] += p3.y;
    }
  }
  if (colsum == nullptr) return;
  __shared__ float red[256][9];  // (+1: conflict-free column reads)
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.x][j] = cs[j];
  __syncthreads();
  if (threadIdx.x < D) {
    const int vc = threadIdx.x >> 3, j = threadIdx.x & 7;
    float t = 0.f;
    for (int r = 0; r < RPB; ++r) t += red[r * VPR + vc][j];
    atomicAdd(colsum + h * D + threadIdx.x, t);
  }
}

template <int D>
int launch_bwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& tdo,
               const AttnBwdArgs& a, const bf16* o, long long o_sb, long long o_sn, long long o_sh, const bf16* d_o,
               long long do_sb, long long do_sn, long long do_sh, const float* lse, bf16* dq, long long dq_sb,
               long long dq_sn, long long dq_sh, float* dq_cs, cudaStream_t stream) {
  using C = BwdCfg<D>;
  const long long rows = (long long)a.B * a.H * a.Nq;
  const long long rows_p = (long long)a.B * a.H * a.NqP;
  cb_launch(attn_delta_kernel<D>, (unsigned)((rows_p * (D / 8) + 255) / 256), 256, 0, stream, 
      o, o_sb, o_sn, o_sh, d_o, do_sb, do_sn, do_sh, lse, const_cast<float*>(a.nl2), const_cast<float*>(a.dsc), a.B, a.H, a.Nq,
      a.NqP, a.scale);
  CB_LAUNCH_CHECK();
  CB_CUDA(cudaMemsetAsync(a.dq_acc, 0, (size_t)rows * D * sizeof(float), stream));
  static const int gen = [] {  // CB_ATTN_BWD=1 selects the first-generation kernel (A/B measurements)
    const char* e = getenv("CB_ATTN_BWD");
    return e != nullptr ? atoi(e) : 2;
  }();
  dim3 grid((a.Nk + 127) / 128, a.H, a.B);
  if (gen == 1) {
    auto kern = attn_bwd_kernel<D>;
    static bool attr_set = false;
    if (!attr_set) {
      CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
      attr_set = true;
    }
    cb_launch(kern, grid, BWD_THREADS, C::SMEM_BYTES, stream, tq, tk, tv, tdo, a);
  } else {
    using C2 = Bwd2Cfg<D>;
    auto kern = attn_bwd2_kernel<D>;
    static bool attr_set = false;
    if (!attr_set) {
      CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C2::SMEM_BYTES));
      attr_set = true;
    }
    // dq_acc [B * H, Nq, D] fp32 as a 3-D tensor: box = 32 columns x 32 rows, 128B swizzle (the drain's staging tile)
    CUtensorMap tdq;
    {
      uint64_t dims[3] = {(uint64_t)D, (uint64_t)a.Nq, (uint64_t)a.B * a.H};
      uint64_t strides[2] = {(uint64_t)D * 4, (uint64_t)a.Nq * D * 4};
      uint32_t box[3] = {32, 32, 1};
      if (int rc = cb_make_tmap_nd_f32(&tdq, a.dq_acc, 3, dims, strides, box, 128)) return rc;
    }
    cb_launch(kern, grid, BWD2_THREADS, C2::SMEM_BYTES, stream, tq, tk, tv, tdo, tdq, a);
  }
  CB_LAUNCH_CHECK();
  {
    constexpr int RPB = 256 / (D / 8);
    const int xb = (a.Nq + 4 * RPB - 1) / (4 * RPB);  // ~4 row passes per block
    cb_launch(attn_dq_convert_kernel<D>, dim3(xb < 1 ? 1 : xb, a.H, a.B), 256, 0, stream, a.dq_acc, dq, dq_sb, dq_sn, dq_sh,
              a.B, a.H, a.Nq, dq_cs);
  }
  CB_LAUNCH_CHECK();
  return 0;
}

}  // namespace

extern "C" int cb_attention_bwd(const void* q, long long q_sb, long long q_sn, long long q_sh, const void* k,
                                long long k_sb, long long k_sn, long long k_sh, const void* v, long long v_sb,
                                long long v_sn, long long v_sh, const void* o, long long o_sb, long long o_sn,
                                long long o_sh, const void* d_o, long long do_sb, long long do_sn, long long do_sh,
                                const float* lse, void* dq, long long dq_sb, long long dq_sn, long long dq_sh, void* dk,
                                long long dk_sb, long long dk_sn, long long dk_sh, void* dv, long long dv_sb,
                                long long dv_sn, long long dv_sh, float* delta, float* dq_acc, int B, int H, int Nq,
                                int Nk, int head_dim, float scale, float* dq_colsum, float* dk_colsum, float* dv_colsum,
                                void* stream) {
  CB_CHECK_ARG(head_dim == 32 || head_dim == 64, "attention_bwd: head_dim %d not supported (32 or 64)", head_dim);
  CB_CHECK_ARG(B > 0 && H > 0 && Nq > 0 && Nk > 0, "attention_bwd: empty problem");
  CB_CHECK_ARG(delta != nullptr && dq_acc != nullptr, "attention_bwd: workspaces missing");
  CUtensorMap tq, tk, tv, tdo;
  if (int rc = make_qkv_tmap(&tq, q, q_sb, q_sn, q_sh, B, H, Nq, head_dim, 128)) return rc;
  if (int rc = make_qkv_tmap(&tk, k, k_sb, k_sn, k_sh, B, H, Nk, head_dim, 128)) return rc;
  if (int rc = make_qkv_tmap(&tv, v, v_sb, v_sn, v_sh, B, H, Nk, head_dim, 128)) return rc;
  if (int rc = make_qkv_tmap(&tdo, d_o, do_sb, do_sn, do_sh, B, H, Nq, head_dim, 128)) return rc;
  AttnBwdArgs a;
  a.NqP = (Nq + 127) / 128 * 128;
  a.nl2 = delta, a.dsc = delta + (long long)B * H * a.NqP, a.dq_acc = dq_acc;
  a.dk = (bf16*)dk, a.dk_sb = dk_sb, a.dk_sn = dk_sn, a.dk_sh = dk_sh;
  a.dv = (bf16*)dv, a.dv_sb = dv_sb, a.dv_sn = dv_sn, a.dv_sh = dv_sh;
  a.B = B, a.H = H, a.Nq = Nq, a.Nk = Nk, a.scale = scale, a.scale_log2 = scale * LOG2E;
  static const int token = [] {
    const char* e = getenv("CB_ATTN_BWD_TOKEN");
    return e != nullptr ? atoi(e) : 1;
  }();
  a.token = token;
  a.dk_cs = dk_colsum, a.dv_cs = dv_colsum;
  {
    const char* e = getenv("CB_ATTN_BWD");
    CB_CHECK_ARG((dk_colsum == nullptr && dv_colsum == nullptr) || e == nullptr || atoi(e) != 1,
                 "attention_bwd: the first-generation kernel (CB_ATTN_BWD=1) has no fused dK / dV column sums");
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (head_dim == 64)
    return launch_bwd<64>(tq, tk, tv, tdo, a, (const bf16*)o, o_sb, o_sn, o_sh, (const bf16*)d_o, do_sb, do_sn, do_sh,
                          lse, (bf16*)dq, dq_sb, dq_sn, dq_sh, dq_colsum, s);
  return launch_bwd<32>(tq, tk, tv, tdo, a, (const bf16*)o, o_sb, o_sn, o_sh, (const bf16*)d_o, do_sb, do_sn, do_sh, lse,
                        (bf16*)dq, dq_sb, dq_sn, dq_sh, dq_colsum, s);
}

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
//             grows by more than 2^8), P -> bf16 -> 128B-swizzled smem as the A operand of P.V.
// The two query tiles ping-pong on the tensor pipe: S_t(j+1) is issued as soon as warpgroup t has
// pulled S_t(j) into registers, so the tensor core runs under the exp / max / sum work.
// TMEM: S0 | S1 (128 cols each) | O0 | O1 (head_dim cols each).
#include "../../include/cinema_b200.h"
#include "common.cuh"

namespace {

constexpr int TQ = 128;       // query rows per tile (UMMA M)
constexpr int TK = 128;       // keys per tile (UMMA N of S, K extent of P.V)
constexpr int FWD_THREADS = 320;
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
  static constexpr int P_BYTES = TQ * TK * 2;                 // 32 KB, two 64-column chunks of 16 KB
  static constexpr int OFF_Q = 0;                             // 2 tiles
  static constexpr int OFF_K = OFF_Q + 2 * TILE_BYTES;        // 2 stages
  static constexpr int OFF_V = OFF_K + 2 * TILE_BYTES;        // 2 stages
  static constexpr int OFF_P = OFF_V + 2 * TILE_BYTES;        // 2 tiles
  static constexpr int OFF_BAR = OFF_P + 2 * P_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
  static constexpr int TMEM_O = 256;                          // column of O0; O1 at +D
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
      const uint32_t p_base = smem_u32(smem + C::OFF_P);
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
          const uint32_t ps = p_base + t * C::P_BYTES;
#pragma unroll
          for (int kk = 0; kk < TK / 16; ++kk) {
            const uint64_t da = umma_smem_desc(ps + (kk >> 2) * 16384 + (kk & 3) * 32, 0, 1024, UMMA_SW128);
            const uint64_t db = umma_smem_desc(vs + kk * 2 * C::GROUP_BYTES, 0, C::GROUP_BYTES, C::SWZ);
            umma_bf16_ss(tmem_base + C::TMEM_O + t * D, da, db, idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
          }
          umma_commit(&pv_done[t]);
        }
        umma_commit(&v_empty[j & 1]);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------ softmax warpgroups
    const int t = warp >> 2;             // query tile of this warpgroup
    const int r = (warp & 3) * 32 + lane;  // row inside the tile == TMEM lane
    if (t < n_tiles_q) {
      const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
      uint8_t* p_smem = smem + C::OFF_P + t * C::P_BYTES;
      float m_ref = -INFINITY;  // reference max (log2 domain) the accumulators are expressed against
      float l = 0.f;
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(&s_full[t], j & 1);
        tcgen05_fence_after();
        float s[TK];
#pragma unroll
        for (int c = 0; c < TK / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(lane_addr + t * TK + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) s[c * 32 + i] = __uint_as_float(v[i]) * p.scale_log2;
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[t]);

        const int valid = p.Nk - j * TK;  // keys of this tile that exist
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < TK; ++i) {
          if (i >= valid) s[i] = -INFINITY;
          mx = fmaxf(mx, s[i]);
        }
        // lazy rescale: keep m_ref unless the row max ran away by more than 2^8
        const bool grow = mx > m_ref + RESCALE_THRESHOLD;
        const float m_new = grow ? mx : m_ref;
        const float alpha = grow ? exp2f(m_ref - m_new) : 1.0f;  // exp2f(-inf) = 0 on the first tile
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
        float sum = 0.f;
#pragma unroll
        for (int vcol = 0; vcol < TK / 8; ++vcol) {
          float e[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            e[i] = exp2f(s[vcol * 8 + i] - m_ref);
            sum += e[i];
          }
          const uint4 pk = make_uint4(pack_bf16(e[0], e[1]), pack_bf16(e[2], e[3]), pack_bf16(e[4], e[5]),
                                      pack_bf16(e[6], e[7]));
          *reinterpret_cast<uint4*>(p_smem + (vcol >> 3) * 16384 + sw128_vec_offset(r, vcol & 7)) = pk;
        }
        l += sum;
        fence_proxy_async_smem();
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
  kern<<<grid, FWD_THREADS, C::SMEM_BYTES, stream>>>(tq, tk, tv, a);
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
  return head_dim == 64 ? launch_fwd<64>(tq, tk, tv, a, s) : launch_fwd<32>(tq, tk, tv, a, s);
}

extern "C" int cb_attention_bwd(const void*, long long, long long, long long, const void*, long long, long long,
                                long long, const void*, long long, long long, long long, const void*, long long,
                                long long, long long, const void*, long long, long long, long long, const float*, void*,
                                long long, long long, long long, void*, long long, long long, long long, void*,
                                long long, long long, long long, float*, float*, int, int, int, int, int, float, void*) {
  cb_set_error("attention_bwd: not built yet");
  return -1;
}

// bf16 x bf16 -> fp32 GEMM on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM),
// operands staged by TMA into 128B-swizzled shared memory, persistent warp-specialised CTAs.
//
//   C[M,N] = alpha * sum_k A(m,k) * B(n,k)      (+ bias[n]) (GELU / GELU') (+ residual[m,n])
//
// Either operand may be K-major (contraction contiguous) or MN-major (its own M/N index
// contiguous), which is all a Linear layer needs without ever materialising a transpose:
//   forward  y  = x W^T : A = x  [M,K]   K-major, B = W  [N,K]  K-major
//   dgrad    dx = dy W  : A = dy [M,N']  K-major, B = W  [N',K'] read MN-major
//   wgrad    dW = dy^T x: A = dy [T,N']  MN-major, B = x [T,K']  MN-major (split-K over the tokens T,
//                          fp32 red.global accumulation straight into the gradient buffer)
// Replaces the cuBLASLt calls behind nn.Linear at cinema/vit.py:472-477, timm Mlp (cinema/vit.py:570-575),
// cinema/mae/mae.py:395,435-440 and cinema/convvit.py:121,294-298 of the reference.
//
// CTA = 576 threads: warp 0 TMA producer, warp 1 MMA issuer (+ TMEM allocator), warps 2..17 epilogue
// (four warps per TMEM lane quarter, each owning a quarter of the tile's columns).  Two TMEM accumulator
// stages let the epilogue of tile i overlap the main loop of tile i+1.
#include "../../include/cinema_b200.h"
#include "common.cuh"

#include <cstdlib>
#include <type_traits>

// In-kernel clock trace (diagnostics, cb_gemm_trace; same idea as cb_attention_trace): when a device buffer is registered,
// the leader CTA of one pair in the middle of the grid stamps clock64() at the phase boundaries of its producer warp, its MMA
// warp and its first epilogue warp for the first eight tiles it processes (slot layout: tools/gemm_trace.py).
// The stamps cost registers in the epilogue warps (56 bytes of spills, +15 % on the short-K launches), so they are compiled
// in only with -DCB_GEMM_TRACE (CB_GEMM_TRACE=1 python -m cinema_b200.build --force); the entry point always exists.
#ifdef CB_GEMM_TRACE
__device__ long long* g_gemm_trace = nullptr;
#endif

extern "C" int cb_gemm_trace(long long* device_buf) {
#ifdef CB_GEMM_TRACE
  CB_CUDA(cudaMemcpyToSymbol(g_gemm_trace, &device_buf, sizeof(device_buf)));
  return 0;
#else
  CB_CHECK_ARG(device_buf == nullptr, "gemm_trace: the library was built without -DCB_GEMM_TRACE");
  return 0;
#endif
}

namespace {

#ifdef CB_GEMM_TRACE
#define CB_GTR(slot)                      \
  do {                                    \
    if (tron) trace[(slot)] = clock64();  \
  } while (0)
#define CB_GTRE(slot)                              \
  do {                                             \
    if (tre && ti < 8) trace[(slot)] = clock64();  \
  } while (0)
#else
#define CB_GTR(slot) do {} while (0)
#define CB_GTRE(slot) do {} while (0)
#endif

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int EPI_WARPS = 16;
constexpr int NUM_THREADS = (2 + EPI_WARPS) * 32;

struct GemmArgs {
  int M, N, K;
  int num_m_tiles, num_n_tiles, splits, kb_per_split, num_kb;
  void* out;
  long long ldo;
  int out_fp32, atomic_add;
  bf16* out2;
  long long ldo2;
  const float* bias;
  const float* residual;
  long long ldr;
  const bf16* aux;
  long long ldaux;
  int epi;
  float alpha;
  float* colsum;  // optional: column sums of the bf16 output (bias gradient of the consumer)
  const float* row_scale;  // optional: per row-group factor on (alpha * acc + bias) before the residual (stochastic depth)
  int rows_per_group;
  int wide;  // wide-store epilogue path allowed (CB_GEMM_WIDE, default 1)
  // CONV instantiations only (appended: the layout seen by the GEMM instantiations is unchanged): K = taps * C_in, the A
  // tile of k-block kb is the 2-D box of the [rows, C_in] tensor map at row m0 + conv_off[kb / conv_kb_per_tap]
  int conv_kb_per_tap;
  int conv_off[27];
};

// CTAS == 2: a CTA pair (cluster of two SMs of one TPC) computes a 256 x BN tile with tcgen05.mma.cta_group::2; each
// CTA stages its own 128 rows of A and HALF of the B tile, so every SM receives 16 + BN/16 KB per k-block instead of
// 16 + BN/8 KB -- the L2 -> SM bandwidth is what bounds the single-CTA kernel (~1000 TFLOP/s with 128 x 256 tiles).
template <int BN, int CTAS = 1>
struct Cfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN / CTAS * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // WIDE: the 256-wide pair kernel (every hot launch of the step) has a second epilogue path for plain bf16 outputs that
  // stores full 128-byte row segments (run_tile_wide below); its per-warp staging tile is 32 rows x 128 B, paid for with one
  // pipeline stage (five instead of six: the long-K launches measured the same)
  static constexpr bool WIDE = CTAS == 2 && BN == 256;
  static constexpr int STAGES = CTAS == 2 ? (BN == 256 ? 5 : 8) : (BN == 256 ? 4 : (BN == 128 ? 6 : 8));
  static constexpr int TMEM_COLS = BN == 256 ? 512 : (BN == 128 ? 256 : 128);
  static constexpr int STAGING_PER_WARP = WIDE ? 4096 : 2048;  // 32 x 16 fp32 transposition tile / 32 x 64 bf16 output tile
  static constexpr int STAGING_BYTES = EPI_WARPS * STAGING_PER_WARP;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int BN, bool A_MN, bool B_MN, int CTAS, bool CONV = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const GemmArgs p) {
  using C = Cfg<BN, CTAS>;
  constexpr bool PAIR = CTAS == 2;
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;  // 0 = leader (issues the MMAs of the pair)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES + C::STAGING_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tfull_bar = empty_bar + C::STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
#ifdef CB_GEMM_TRACE
  long long* const trace = g_gemm_trace;
  const bool tron = trace != nullptr && lane == 0 && blockIdx.x == ((gridDim.x / 2) & ~1u);
#endif
  if (warp == 0) CB_GTR(1000);

  pdl_launch_dependents();  // the next kernel may be scheduled as SMs drain; it waits for this grid before reading
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < C::STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(&tfull_bar[s], 1);
        mbar_init(&tempty_bar[s], EPI_WARPS * CTAS);  // pair: the leader collects the epilogue warps of both CTAs
      }
      mbar_fence_init();
    }
    __syncwarp();
    if constexpr (PAIR) tmem_alloc_pair(tmem_slot, C::TMEM_COLS);
    else tmem_alloc(tmem_slot, C::TMEM_COLS);
  }
  tcgen05_fence_before();
  __syncthreads();                     // this CTA: barrier words and the TMEM slot written by warp 1 are visible
  if constexpr (PAIR) cluster_sync();  // both CTAs' barriers are initialised before any remote arrive / multicast commit
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // barriers, TMEM and descriptors are set up: now wait for the producer kernels of A / B / residual
  if (warp == 0) CB_GTR(1001);

  const int tiles = p.num_m_tiles * p.num_n_tiles;  // pair: num_m_tiles counts 256-row tiles
  const int items = tiles * p.splits;
  const int worker = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;      // tile-scheduler slot of this CTA (pair)
  const int n_workers = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  constexpr int TILE_M = BM * CTAS;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    // (producer and MMA issuer run as converged warps with one elected lane: under `lane == 0` nvcc wraps every TMA /
    // tcgen05 instruction in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop and rebuilds its operands in front of it --
    // ~67 clocks per instruction measured in the attention kernels, which short-K GEMMs cannot hide)
    const bool leader = elect_one_sync() != 0;
    {
      int stage = 0;
      uint32_t phase = 0;
      [[maybe_unused]] int ti = 0;
      for (int item = worker; item < items; item += n_workers, ++ti) {
        const int split = item % p.splits;
        const int tile = item / p.splits;
        const int n_tile = tile % p.num_n_tiles;
        const int m_tile = tile / p.num_n_tiles;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.num_kb);
        const int m0 = m_tile * TILE_M + (int)cta_rank * BM;          // this CTA's rows of A
        const int n0 = n_tile * BN + (int)cta_rank * (BN / CTAS);     // this CTA's share of the B tile
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (ti < 8 && kb - kb0 < 16) CB_GTR(64 * ti + (kb - kb0));
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + C::A_BYTES;
          if (leader) {
          // pair: both CTAs' loads complete on the LEADER's barrier, which expects the bytes of both
          if (!PAIR || cta_rank == 0) mbar_expect_tx(&full_bar[stage], C::STAGE_BYTES * CTAS);
          if constexpr (CONV) {
            // implicit im2col: filter tap t reads the SAME 2-D matrix shifted by a constant number of rows; rows outside
            // the matrix (negative or past the end) are zero-filled by TMA, which is the convolution's zero padding
            const int tap = kb / p.conv_kb_per_tap;
            tma_load_2d_g<CTAS>(sa, &tma_a, &full_bar[stage], (kb - tap * p.conv_kb_per_tap) * BK, m0 + p.conv_off[tap]);
          } else if constexpr (!A_MN) {
            tma_load_2d_g<CTAS>(sa, &tma_a, &full_bar[stage], kb * BK, m0);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)
              tma_load_2d_g<CTAS>(sa + j * 8192, &tma_a, &full_bar[stage], m0 + j * 64, kb * BK);
          }
          if constexpr (!B_MN) {
            tma_load_2d_g<CTAS>(sb, &tma_b, &full_bar[stage], kb * BK, n0);
          } else {
#pragma unroll
            for (int j = 0; j < BN / CTAS / 64; ++j)
              tma_load_2d_g<CTAS>(sb + j * 8192, &tma_b, &full_bar[stage], n0 + j * 64, kb * BK);
          }
          }
          __syncwarp();
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    const bool leader = elect_one_sync() != 0;
    if (cta_rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(TILE_M, BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      [[maybe_unused]] int ti = 0;
      for (int item = worker; item < items; item += n_workers, ++ti) {
        const int split = item % p.splits;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.num_kb);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        if (ti < 8) CB_GTR(64 * ti + 32);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          if (ti < 8 && kb - kb0 < 16) CB_GTR(64 * ti + 16 + (kb - kb0));
          tcgen05_fence_after();
          if (leader) {
            const uint32_t a_base = smem_u32(smem + stage * C::STAGE_BYTES);
            const uint32_t b_base = a_base + C::A_BYTES;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t da = A_MN ? umma_smem_desc(a_base + k * 2048, 8192, 1024, UMMA_SW128)
                                       : umma_smem_desc(a_base + k * 32, 0, 1024, UMMA_SW128);
              const uint64_t db = B_MN ? umma_smem_desc(b_base + k * 2048, 8192, 1024, UMMA_SW128)
                                       : umma_smem_desc(b_base + k * 32, 0, 1024, UMMA_SW128);
              umma_bf16_ss_g<CTAS>(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            }
            umma_commit_g<CTAS>(&empty_bar[stage]);  // smem slot (of both CTAs) reusable once these MMAs retire
            if (kb + 1 == kb1) umma_commit_g<CTAS>(&tfull_bar[acc]);  // accumulator complete -> epilogue (of both CTAs)
          }
          __syncwarp();
          if (ti < 8 && kb + 1 == kb1) CB_GTR(64 * ti + 33);
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------ epilogue ------------------------------
    // 16 warps: four per TMEM lane quarter, each owning a quarter of the tile's columns in chunks of 16.
    // TMEM -> registers (thread == accumulator row) -> per-warp swizzled smem tile -> registers in a row-coalesced
    // layout (4 lanes x 16 B cover 64 B of one row, 8 rows per instruction), so that every global access of the
    // epilogue (bias, residual, GELU' input, fp32 / bf16 stores, split-K reductions) moves whole 32-byte sectors.
    // Residual / GELU' operands of chunk c+1 are fetched into registers before the math of chunk c (and those of
    // a tile's first chunk before the wait on its accumulator), which hides their latency behind the main loop.
    const int ew = warp - 2;
    const int q = warp & 3;    // TMEM lane quarter this warp may touch
    const int cq = ew >> 2;    // which quarter of the tile's columns
    constexpr int COLS_PER_WARP = BN / 4;
    constexpr int CHUNKS = COLS_PER_WARP / 16;
    const uint32_t stg = smem_u32(smem + C::STAGES * C::STAGE_BYTES + ew * C::STAGING_PER_WARP);
    const int sub = lane >> 2;  // row inside a group of eight
    const int c16 = lane & 3;   // 16-byte column (4 fp32) inside the 16-column chunk
    const bool has_res = p.residual != nullptr;
    const bool has_aux = p.epi == CB_EPI_GELU_BWD;
    const bool has_bias = p.bias != nullptr;
    const bool has_out2 = p.out2 != nullptr;
    const bool has_rs = p.row_scale != nullptr;
    // all epilogue addressing is base pointer + 32-bit element offset (host checks M * ld < 2^31): one IMAD.WIDE per
    // access instead of 64-bit row arithmetic, which used to be half of the epilogue's instructions
    const int so8 = (int)(8 * p.ldo), s28 = (int)(8 * p.ldo2), sr8 = (int)(8 * p.ldr), sa8 = (int)(8 * p.ldaux);
    float* const out32 = reinterpret_cast<float*>(p.out);
    bf16* const out16 = reinterpret_cast<bf16*>(p.out);
    int acc = 0;
    uint32_t acc_phase = 0;
#ifdef CB_GEMM_TRACE
    int ti = 0;
    const bool tre = tron && ew == 0;
#endif
    float4 pre_res[4];
    uint2 pre_aux[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) pre_res[i] = make_float4(0.f, 0.f, 0.f, 0.f), pre_aux[i] = make_uint2(0u, 0u);

    // one 128 x BN accumulator tile; MODE is a compile-time constant so that each epilogue flavour gets its own loop
    auto run_tile = [&](auto mode_c, int item) {
      constexpr int MODE = decltype(mode_c)::value;  // 0 bf16 out, 1 GELU, 2 GELU', 3 fp32 out, 4 fp32 atomic add
      const int tile = item / p.splits;
      const int n_tile = tile % p.num_n_tiles;
      const int m_tile = tile / p.num_n_tiles;
      const int r0 = m_tile * TILE_M + (int)cta_rank * BM + q * 32 + sub;  // this lane's first row (i = 0)
      const int colw = n_tile * BN + cq * COLS_PER_WARP + c16 * 4;         // this lane's first column in chunk 0
      uint32_t vmask = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) vmask |= (r0 + 8 * i < p.M ? 1u : 0u) << i;
      float rs[4] = {1.f, 1.f, 1.f, 1.f};
      if ((MODE == 0 || MODE == 3) && has_rs) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if ((vmask >> i) & 1u) rs[i] = __ldg(p.row_scale + (r0 + 8 * i) / p.rows_per_group);
      }
      const int oo = (int)((long long)r0 * p.ldo) + colw;
      const int o2 = (int)((long long)r0 * p.ldo2) + colw;
      const int ro = (int)((long long)r0 * p.ldr) + colw;
      const int ao = (int)((long long)r0 * p.ldaux) + colw;
      auto prefetch = [&](int c) {  // residual / GELU' operands of chunk c into registers
        if (colw + c * 16 >= p.N) return;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if ((vmask >> i) & 1u) {
            if ((MODE == 0 || MODE == 3) && has_res)
              pre_res[i] = __ldg(reinterpret_cast<const float4*>(p.residual + (ro + i * sr8 + c * 16)));
            if (MODE == 2) pre_aux[i] = __ldg(reinterpret_cast<const uint2*>(p.aux + (ao + i * sa8 + c * 16)));
          }
        }
      };
      if (MODE == 2 || ((MODE == 0 || MODE == 3) && has_res)) {
        prefetch(0);
        // pull the NEXT tile's residual / GELU' operand rows of this warp into L2 (one 128-byte line per lane and step)
        const int nitem = item + n_workers;
        if (nitem < items) {
          const int ntile = nitem / p.splits;
          const long long nrow = (long long)(ntile / p.num_n_tiles) * TILE_M + (long long)cta_rank * BM + q * 32 + lane;
          const int ncol = (ntile % p.num_n_tiles) * BN + cq * COLS_PER_WARP;
          if (nrow < p.M && ncol < p.N) {
            if (MODE == 2) {
              const char* a = reinterpret_cast<const char*>(p.aux + nrow * p.ldaux + ncol);
#pragma unroll
              for (int o = 0; o < COLS_PER_WARP * 2; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + o));
            } else {
              const char* a = reinterpret_cast<const char*>(p.residual + nrow * p.ldr + ncol);
#pragma unroll
              for (int o = 0; o < COLS_PER_WARP * 4; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + o));
            }
          }
        }
      }
      CB_GTRE(64 * ti + 40);
      mbar_wait(&tfull_bar[acc], acc_phase);
      CB_GTRE(64 * ti + 41);
      tcgen05_fence_after();
#pragma unroll 1
      for (int c = 0; c < CHUNKS; ++c) {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + cq * COLS_PER_WARP + c * 16;
        uint32_t r[16];
        tmem_ld_32x32b_x16(taddr, r);
        tmem_ld_wait();
        const int col = colw + c * 16;
        const bool col_ok = col < p.N;  // N is a multiple of 8
#pragma unroll
        for (int j = 0; j < 4; ++j)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)),
                       "r"(r[4 * j]), "r"(r[4 * j + 1]), "r"(r[4 * j + 2]), "r"(r[4 * j + 3])
                       : "memory");
        __syncwarp();
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (has_bias && col_ok) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
        float x[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rr = i * 8 + sub;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(x[i][0]), "=f"(x[i][1]), "=f"(x[i][2]), "=f"(x[i][3])
                       : "r"(stg + rr * 64 + ((c16 ^ ((rr >> 1) & 3)) << 4)));
        }
        __syncwarp();  // staging tile is rewritten by the next chunk
        float4 res[4];
        uint2 aux[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) res[i] = pre_res[i], aux[i] = pre_aux[i];
        if ((MODE == 2 || ((MODE == 0 || MODE == 3) && has_res)) && c + 1 < CHUNKS) prefetch(c + 1);
        const bool want_cs = (MODE == 0 || MODE == 2) && p.colsum != nullptr;
        if (!col_ok && !want_cs) continue;  // (with colsum every lane stays for the shuffles below)
        float cs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (!((vmask >> i) & 1u) || !col_ok) continue;
          float* v = x[i];
          v[0] = fmaf(v[0], p.alpha, bias4.x), v[1] = fmaf(v[1], p.alpha, bias4.y);
          v[2] = fmaf(v[2], p.alpha, bias4.z), v[3] = fmaf(v[3], p.alpha, bias4.w);
          const int eo = oo + i * so8 + c * 16;
          if constexpr (MODE == 1) {
            // pre-activation is rounded to bf16 first (what autocast feeds nn.GELU), saved for backward
            const uint2 pre = make_uint2(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]));
            if (p.out != nullptr) *reinterpret_cast<uint2*>(out16 + eo) = pre;
            const float2 p0 = unpack_bf16(pre.x), p1 = unpack_bf16(pre.y);
            *reinterpret_cast<uint2*>(p.out2 + (o2 + i * s28 + c * 16)) =
                make_uint2(pack_bf16(gelu_fast(p0.x), gelu_fast(p0.y)), pack_bf16(gelu_fast(p1.x), gelu_fast(p1.y)));
          } else if constexpr (MODE == 2) {
            const float2 a0 = unpack_bf16(aux[i].x), a1 = unpack_bf16(aux[i].y);
            v[0] *= gelu_grad_fast(a0.x), v[1] *= gelu_grad_fast(a0.y);
            v[2] *= gelu_grad_fast(a1.x), v[3] *= gelu_grad_fast(a1.y);
            const uint2 pk = make_uint2(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]));
            *reinterpret_cast<uint2*>(out16 + eo) = pk;
            if (want_cs) {
              const float2 c0 = unpack_bf16(pk.x), c1 = unpack_bf16(pk.y);
              cs[0] += c0.x, cs[1] += c0.y, cs[2] += c1.x, cs[3] += c1.y;
            }
          } else if constexpr (MODE == 4) {
            red_add_v4(out32 + eo, v[0], v[1], v[2], v[3]);
          } else {
            if (has_rs) v[0] *= rs[i], v[1] *= rs[i], v[2] *= rs[i], v[3] *= rs[i];
            if (has_res) v[0] += res[i].x, v[1] += res[i].y, v[2] += res[i].z, v[3] += res[i].w;
            const uint2 pk = make_uint2(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]));
            if constexpr (MODE == 3) {
              *reinterpret_cast<float4*>(out32 + eo) = make_float4(v[0], v[1], v[2], v[3]);
              if (has_out2) *reinterpret_cast<uint2*>(p.out2 + (o2 + i * s28 + c * 16)) = pk;  // bf16 shadow
            } else {
              *reinterpret_cast<uint2*>(out16 + eo) = pk;
              if (has_out2) *reinterpret_cast<uint2*>(p.out2 + (o2 + i * s28 + c * 16)) = pk;
              if (want_cs) {
                const float2 c0 = unpack_bf16(pk.x), c1 = unpack_bf16(pk.y);
                cs[0] += c0.x, cs[1] += c0.y, cs[2] += c1.x, cs[3] += c1.y;
              }
            }
          }
        }
        if (want_cs) {  // 32 rows x 4 columns per lane quad: reduce over the eight row groups, one red.v4 per quad
#pragma unroll
          for (int o = 4; o < 32; o <<= 1) {
#pragma unroll
            for (int k = 0; k < 4; ++k) cs[k] += __shfl_xor_sync(0xffffffffu, cs[k], o);
          }
          if (sub == 0 && col_ok) red_add_v4(p.colsum + col, cs[0], cs[1], cs[2], cs[3]);
        }
        CB_GTRE(64 * ti + 42 + c);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR && cta_rank != 0) mbar_arrive_cluster(&tempty_bar[acc], 0);  // the leader's barrier gates the MMAs
        else mbar_arrive(&tempty_bar[acc]);
      }
      CB_GTRE(64 * ti + 50);
#ifdef CB_GEMM_TRACE
      ++ti;
#endif
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    };

    // ---- wide-store path (256-wide pair kernel, plain bf16 output: no residual / second output / column sums / row scale).
    // The in-kernel clock trace (profiles/r02_gemm_clock_trace.md) showed the epilogue of the short-K launches bound by the
    // NUMBER of global store requests: every STG of the path above covers eight rows x 32 bytes = eight requests; a plain
    // bf16 tile took 1.5 k clocks per 16-column chunk, 7.5 k per tile against 4.3 k clocks of MMA at K = 512.  Here the
    // arithmetic runs in the accumulator's own layout (thread == row, 16 consecutive columns per chunk, the bias is a
    // warp-uniform load), the packed bf16 rows of all four chunks are collected in a 128B-swizzled [32 rows x 128 B] tile
    // per warp and leave as 16-byte vectors covering whole 128-byte row segments: 4 x fewer requests, 3.3 k clocks per tile
    // (K = 512, N = 2048: 71.7 -> 61.4 us).  The accumulator stage goes back to the MMA warp as soon as its last chunk is in
    // registers.  (The GELU flavours did NOT profit from the same treatment -- their epilogue is bound by instruction issue
    // and the MUFU pipe, 1.85 k clocks of arithmetic per chunk -- and stay on the path above.)
    auto run_tile_wide = [&](int item) {
      const int tile = item / p.splits;
      const int n_tile = tile % p.num_n_tiles;
      const int m_tile = tile / p.num_n_tiles;
      const int row0 = m_tile * TILE_M + (int)cta_rank * BM + q * 32;  // first row of this warp
      const int col0 = n_tile * BN + cq * COLS_PER_WARP;               // first column of this warp
      const uint32_t my_row = stg + lane * 128;                        // this thread's row of the staging tile
      const uint32_t sw = (uint32_t)(lane & 7);
      CB_GTRE(64 * ti + 40);
      mbar_wait(&tfull_bar[acc], acc_phase);
      CB_GTRE(64 * ti + 41);
      tcgen05_fence_after();
      const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + cq * COLS_PER_WARP;
#pragma unroll 1
      for (int c = 0; c < CHUNKS; ++c) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(taddr0 + c * 16, r);
        const int col = col0 + c * 16;
        float4 b4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          b4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (has_bias && col + 4 * k < p.N) b4[k] = __ldg(reinterpret_cast<const float4*>(p.bias + col + 4 * k));
        }
        tmem_ld_wait();
        if (c == CHUNKS - 1) {  // the accumulator is in registers: release the stage before the last chunk's arithmetic
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (PAIR && cta_rank != 0) mbar_arrive_cluster(&tempty_bar[acc], 0);
            else mbar_arrive(&tempty_bar[acc]);
          }
        }
        uint32_t w[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          w[2 * k] = pack_bf16(fmaf(__uint_as_float(r[4 * k + 0]), p.alpha, b4[k].x), fmaf(__uint_as_float(r[4 * k + 1]), p.alpha, b4[k].y));
          w[2 * k + 1] = pack_bf16(fmaf(__uint_as_float(r[4 * k + 2]), p.alpha, b4[k].z), fmaf(__uint_as_float(r[4 * k + 3]), p.alpha, b4[k].w));
        }
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2)  // 16 columns = 32 bytes = units 2c, 2c + 1 of this thread's row
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my_row + ((((uint32_t)(2 * c + h2)) ^ sw) << 4)),
                       "r"(w[4 * h2]), "r"(w[4 * h2 + 1]), "r"(w[4 * h2 + 2]), "r"(w[4 * h2 + 3])
                       : "memory");
        CB_GTRE(64 * ti + 42 + c);
      }
      // flush: lane = (row within a group of four, 16-byte unit), 8 lanes cover one 128-byte row of the tile
      {
        const int f_row = lane >> 3, f_u = lane & 7;
        const int rows_left = p.M - row0 - f_row;  // row f_row + 4 it exists iff 4 it < rows_left
        bf16* d = out16 + (long long)(row0 + f_row) * p.ldo + col0 + f_u * 8;
        const long long f_step = 4 * p.ldo;
        const bool ok = col0 + f_u * 8 < p.N;
        const uint32_t f_src = stg + f_row * 128;
        __syncwarp();
        uint4 val[8];
#pragma unroll
        for (int it = 0; it < 8; ++it)  // row = 4 it + f_row: (row & 7) = 4 (it & 1) + f_row
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(val[it].x), "=r"(val[it].y), "=r"(val[it].z), "=r"(val[it].w)
                       : "r"(f_src + it * 512 + (((uint32_t)f_u ^ (uint32_t)(4 * (it & 1) + f_row)) << 4)));
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          if (ok && 4 * it < rows_left) *reinterpret_cast<uint4*>(d) = val[it];
          d += f_step;
        }
        __syncwarp();  // the tile is rewritten by the next item
      }
      CB_GTRE(64 * ti + 50);
#ifdef CB_GEMM_TRACE
      ++ti;
#endif
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    };
    const int mode = p.epi == CB_EPI_GELU ? 1 : (has_aux ? 2 : (p.out_fp32 ? (p.atomic_add ? 4 : 3) : 0));
    // (`CB_GEMM_WIDE=0` -> GemmArgs::wide = 0 keeps everything on the narrow path for A/B measurements)
    const bool wide = C::WIDE && p.wide != 0 && mode == 0 && !has_res && !has_out2 && !has_rs && p.colsum == nullptr;
    for (int item = worker; item < items; item += n_workers) {
      if constexpr (C::WIDE) {
        if (wide) {
          run_tile_wide(item);
          continue;
        }
      }
      switch (mode) {
        case 0: run_tile(std::integral_constant<int, 0>{}, item); break;
        case 1: run_tile(std::integral_constant<int, 1>{}, item); break;
        case 2: run_tile(std::integral_constant<int, 2>{}, item); break;
        case 3: run_tile(std::integral_constant<int, 3>{}, item); break;
        default: run_tile(std::integral_constant<int, 4>{}, item); break;
      }
    }
  }

  if (warp == 0) CB_GTR(1002);
  tcgen05_fence_before();
  if constexpr (PAIR) cluster_sync();  // no CTA of the pair leaves (or frees TMEM) while its peer may still touch it
  else __syncthreads();
  if (warp == 0) CB_GTR(1003);
  if (warp == 1) {
    tcgen05_fence_after();
    if constexpr (PAIR) tmem_dealloc_pair(tmem_base, C::TMEM_COLS);
    else tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

template <int BN, bool A_MN, bool B_MN, int CTAS, bool CONV = false>
int launch(const void* A, long long lda, const void* B, long long ldb, GemmArgs& p, cudaStream_t stream) {
  using C = Cfg<BN, CTAS>;
  static_assert(!CONV || (!A_MN && !B_MN), "the convolution producer reads K-major operands");
  CUtensorMap ta, tb;
  int rc;
  if (CONV)  // the activation matrix itself: [rows, C_in]; K of the GEMM is taps * C_in
    rc = cb_make_tmap_2d(&ta, A, (uint64_t)p.conv_kb_per_tap * BK, (uint64_t)p.M, (uint64_t)lda * 2, BK, BM, 128);
  else if (!A_MN)
    rc = cb_make_tmap_2d(&ta, A, (uint64_t)p.K, (uint64_t)p.M, (uint64_t)lda * 2, BK, BM, 128);
  else
    rc = cb_make_tmap_2d(&ta, A, (uint64_t)p.M, (uint64_t)p.K, (uint64_t)lda * 2, 64, BK, 128);
  if (rc) return rc;
  if (!B_MN)
    rc = cb_make_tmap_2d(&tb, B, (uint64_t)p.K, (uint64_t)p.N, (uint64_t)ldb * 2, BK, BN / CTAS, 128);
  else
    rc = cb_make_tmap_2d(&tb, B, (uint64_t)p.N, (uint64_t)p.K, (uint64_t)ldb * 2, 64, BK, 128);
  if (rc) return rc;
  auto kern = gemm_bf16_kernel<BN, A_MN, B_MN, CTAS, CONV>;
  static bool attr_set = false;  // per template instantiation
  if (!attr_set) {
    CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  const int items = p.num_m_tiles * p.num_n_tiles * p.splits;
  if constexpr (CTAS == 1) {
    const int grid = items < cb_sm_count() ? items : cb_sm_count();
    cb_launch(kern, grid, NUM_THREADS, C::SMEM_BYTES, stream, ta, tb, p);
  } else {
    const int pairs = cb_sm_count() / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * (unsigned)(items < pairs ? items : pairs));
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr, cfg.numAttrs = cb_pdl_enabled() ? 2 : 1;
    CB_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, p));
  }
  CB_LAUNCH_CHECK();
  return 0;
}

template <int BN, int CTAS>
int dispatch_major(int a_mn, int b_mn, const void* A, long long lda, const void* B, long long ldb, GemmArgs& p,
                   cudaStream_t s) {
  if (!a_mn && !b_mn) return launch<BN, false, false, CTAS>(A, lda, B, ldb, p, s);
  if (!a_mn && b_mn) return launch<BN, false, true, CTAS>(A, lda, B, ldb, p, s);
  if (a_mn && b_mn) return launch<BN, true, true, CTAS>(A, lda, B, ldb, p, s);
  return launch<BN, true, false, CTAS>(A, lda, B, ldb, p, s);
}

}  // namespace

extern "C" int cb_gemm_bf16(const void* A, long long lda, int a_mn_major, const void* B, long long ldb, int b_mn_major,
                            int M, int N, int K, void* out, long long ldo, int out_dtype, int accumulate, void* out2,
                            long long ldo2, const float* bias, const float* residual, long long ldr, const void* aux,
                            long long ldaux, int epilogue, float alpha, int split_k, int block_n, float* colsum,
                            const float* row_scale, int rows_per_group, void* stream) {
  CB_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  CB_CHECK_ARG(N % 8 == 0, "gemm: N=%d must be a multiple of 8", N);
  CB_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0, "gemm: lda/ldb must be multiples of 8 elements (16 B TMA pitch)");
  CB_CHECK_ARG(out_dtype == CB_DT_BF16 || out_dtype == CB_DT_F32, "gemm: bad out dtype %d", out_dtype);
  CB_CHECK_ARG(!(accumulate && out_dtype != CB_DT_F32), "gemm: accumulate needs an fp32 output");
  CB_CHECK_ARG(epilogue >= 0 && epilogue <= CB_EPI_GELU_BWD, "gemm: bad epilogue %d", epilogue);
  CB_CHECK_ARG(epilogue != CB_EPI_GELU || (out2 != nullptr && out_dtype == CB_DT_BF16),
               "gemm: GELU epilogue writes bf16 pre-activation to out (optional) and activation to out2");
  CB_CHECK_ARG(epilogue != CB_EPI_GELU_BWD || aux != nullptr, "gemm: GELU' epilogue needs aux (pre-activation)");
  CB_CHECK_ARG(out != nullptr || epilogue == CB_EPI_GELU, "gemm: out is null");
  CB_CHECK_ARG(ldo % 8 == 0 && (out2 == nullptr || ldo2 % 8 == 0) && (residual == nullptr || ldr % 4 == 0) &&
                   (aux == nullptr || ldaux % 8 == 0),
               "gemm: epilogue leading dimensions must keep 16-byte row alignment");

  {
    const long long lim = (1ll << 31) - 1;
    const long long rows = (long long)M + 2 * BM;  // tile overhang is never dereferenced but is multiplied
    CB_CHECK_ARG(rows * ldo < lim && rows * ldo2 < lim && rows * ldr < lim && rows * ldaux < lim,
                 "gemm: epilogue tensors must stay below 2^31 elements (32-bit epilogue offsets)");
  }
  GemmArgs p;
  p.M = M, p.N = N, p.K = K;
  p.num_kb = (K + BK - 1) / BK;
  p.out = out, p.ldo = ldo, p.out_fp32 = out_dtype == CB_DT_F32, p.atomic_add = accumulate;
  p.out2 = reinterpret_cast<bf16*>(out2), p.ldo2 = ldo2;
  p.bias = bias, p.residual = residual, p.ldr = ldr;
  p.aux = reinterpret_cast<const bf16*>(aux), p.ldaux = ldaux;
  p.epi = epilogue, p.alpha = alpha, p.colsum = colsum;
  p.row_scale = row_scale, p.rows_per_group = rows_per_group > 0 ? rows_per_group : 1;
  static const int wide_ok = [] {
    const char* e = getenv("CB_GEMM_WIDE");
    return e != nullptr ? atoi(e) : 1;
  }();
  p.wide = wide_ok;
  CB_CHECK_ARG(row_scale == nullptr || (epilogue == CB_EPI_NONE && !accumulate && rows_per_group > 0),
               "gemm: row_scale needs the plain epilogue, a non-accumulating output and rows_per_group > 0");
  CB_CHECK_ARG(colsum == nullptr || (out_dtype == CB_DT_BF16 && epilogue != CB_EPI_GELU && !accumulate),
               "gemm: colsum needs a bf16 output without the GELU epilogue");
  CB_CHECK_ARG(((uintptr_t)colsum & 15) == 0, "gemm: colsum must be 16-byte aligned");
  p.num_m_tiles = (M + BM - 1) / BM;

  const int sms = cb_sm_count();
  // Tile width and split-K factor from a small cost model (clocks per CTA): a k-block costs the slower of the MMA
  // (135 clk per 128 x 256 x 16 instruction) and of its operand bytes at the per-SM share of the L2 -> SM bandwidth
  // (~42 B/clk: B300_MICROARCH.md "LTS throughput cap" / 148 SMs), plus a fixed per-item prologue / epilogue.
  auto plan = [&](int c, int ctas, int want_splits, int& out_splits, int& out_kb_per_split) {
    const int workers = sms / ctas;
    const long long tiles = (long long)((M + BM * ctas - 1) / (BM * ctas)) * ((N + c - 1) / c);
    int sp = want_splits;
    if (sp <= 0) {
      sp = 1;
      if (accumulate && tiles < workers) {  // wgrad-like: few tiles, long contraction -> split K
        sp = (int)(workers / tiles);
        const int max_splits = p.num_kb / 4 > 0 ? p.num_kb / 4 : 1;
        if (sp > max_splits) sp = max_splits;
      }
    }
    if (sp > p.num_kb) sp = p.num_kb;
    const int kbs = (p.num_kb + sp - 1) / sp;
    sp = (p.num_kb + kbs - 1) / kbs;
    out_splits = sp, out_kb_per_split = kbs;
    const long long items = tiles * sp;
    const long long waves = (items + workers - 1) / workers;
    const double t_mma = 542.0 * c / 256.0;
    const double t_load = (16384.0 + 128.0 * c / ctas) / 42.0;
    const double t_kb = t_mma > t_load ? t_mma : t_load;
    return (double)waves * (kbs * t_kb + 1500.0 + 6.0 * c);
  };
  static const int force_ctas = [] {
    const char* e = getenv("CB_GEMM_CTAS");
    return e != nullptr ? atoi(e) : 0;
  }();
  int bn = block_n, ctas = 1, splits = 1;
  {
    double best = 1e30;
    const int cands[3] = {256, 128, 64};
    for (int i = 0; i < 3; ++i) {
      const int c = cands[i];
      if (block_n == 64 || block_n == 128 || block_n == 256) {
        if (c != block_n) continue;
      } else if (c > 64 && N <= c / 2) {
        continue;
      }
      for (int g = 1; g <= 2; ++g) {
        if (g == 2 && (c == 64 || M <= BM)) continue;  // pairs: 256-row tiles, B halves of at least 64 columns
        if (force_ctas != 0 && g != force_ctas && !(force_ctas == 2 && (c == 64 || M <= BM))) continue;
        int sp, kbs;
        const double cost = plan(c, g, split_k, sp, kbs);
        if (cost < best) best = cost, bn = c, ctas = g;
      }
    }
  }
  plan(bn, ctas, split_k, splits, p.kb_per_split);
  p.num_m_tiles = (M + BM * ctas - 1) / (BM * ctas);
  p.num_n_tiles = (N + bn - 1) / bn;
  p.splits = splits;
  CB_CHECK_ARG(splits == 1 || accumulate, "gemm: split-K needs accumulate=1 (fp32 atomics into a zeroed/accumulated out)");
  CB_CHECK_ARG(p.splits == 1 || (bias == nullptr && residual == nullptr && epilogue == CB_EPI_NONE),
               "gemm: split-K supports no bias/residual/activation");

  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (ctas == 2) {
    if (bn == 256) return dispatch_major<256, 2>(a_mn_major, b_mn_major, A, lda, B, ldb, p, s);
    return dispatch_major<128, 2>(a_mn_major, b_mn_major, A, lda, B, ldb, p, s);
  }
  switch (bn) {
    case 256: return dispatch_major<256, 1>(a_mn_major, b_mn_major, A, lda, B, ldb, p, s);
    case 128: return dispatch_major<128, 1>(a_mn_major, b_mn_major, A, lda, B, ldb, p, s);
    default: return dispatch_major<64, 1>(a_mn_major, b_mn_major, A, lda, B, ldb, p, s);
  }
}

// EXPERIMENTAL (not on any product path yet, not validated on a GPU in round 1; see DESIGN.md section 8b and
// tools/conv_rowspace_prototype.py for the arithmetic): 3^n "same" convolution over a zero-haloed channel-last row space
// as one K-concatenated GEMM.  X [rows, c_in] bf16 (ldx elements per row), W [c_out, n_taps * c_in] bf16 tap-major,
// row_off[n_taps] (host array) the constant row offset of each tap; out [rows, c_out] bf16 / fp32 (+ bias, + residual).
extern "C" int cb_conv_gemm_bf16(const void* X, long long ldx, long long rows, int c_in, const void* W, long long ldw,
                                 int c_out, int n_taps, const int* row_off, void* out, long long ldo, int out_dtype,
                                 const float* bias, const float* residual, long long ldr, int block_n, void* stream) {
  CB_CHECK_ARG(rows > 0 && rows < (1ll << 31) - 2 * BM && c_in > 0 && c_out > 0, "conv_gemm: bad problem size");
  CB_CHECK_ARG(c_in % BK == 0, "conv_gemm: c_in=%d must be a multiple of %d (one k-block never straddles two taps)", c_in, BK);
  CB_CHECK_ARG(c_out % 8 == 0 && ldx % 8 == 0 && ldw % 8 == 0 && ldo % 8 == 0 && (residual == nullptr || ldr % 4 == 0),
               "conv_gemm: channel counts / leading dimensions must keep 16-byte row alignment");
  CB_CHECK_ARG(n_taps >= 1 && n_taps <= 27 && row_off != nullptr, "conv_gemm: 1..27 taps");
  CB_CHECK_ARG(out != nullptr && (out_dtype == CB_DT_BF16 || out_dtype == CB_DT_F32), "conv_gemm: bad output");
  CB_CHECK_ARG((rows + 2 * BM) * ldo < (1ll << 31) - 1 && (rows + 2 * BM) * ldr < (1ll << 31) - 1,
               "conv_gemm: epilogue tensors must stay below 2^31 elements");
  GemmArgs p = {};
  p.M = (int)rows, p.N = c_out, p.K = n_taps * c_in;
  p.num_kb = p.K / BK;
  p.out = out, p.ldo = ldo, p.out_fp32 = out_dtype == CB_DT_F32, p.atomic_add = 0;
  p.bias = bias, p.residual = residual, p.ldr = ldr;
  p.epi = CB_EPI_NONE, p.alpha = 1.0f, p.rows_per_group = 1;
  p.conv_kb_per_tap = c_in / BK;
  for (int t = 0; t < n_taps; ++t) p.conv_off[t] = row_off[t];
  int bn = block_n;
  if (bn != 64 && bn != 128 && bn != 256) bn = c_out > 128 ? 256 : (c_out > 64 ? 128 : 64);
  const long long pair_tiles = (rows + 2 * BM - 1) / (2 * BM) * ((c_out + bn - 1) / bn);
  const int ctas = (bn > 64 && pair_tiles >= cb_sm_count() / 2) ? 2 : 1;  // pairs once every pair slot has a tile
  p.num_m_tiles = (int)((rows + (long long)BM * ctas - 1) / (BM * ctas));
  p.num_n_tiles = (c_out + bn - 1) / bn;
  p.splits = 1, p.kb_per_split = p.num_kb;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (ctas == 2) {
    if (bn == 256) return launch<256, false, false, 2, true>(X, ldx, W, ldw, p, s);
    return launch<128, false, false, 2, true>(X, ldx, W, ldw, p, s);
  }
  switch (bn) {
    case 256: return launch<256, false, false, 1, true>(X, ldx, W, ldw, p, s);
    case 128: return launch<128, false, false, 1, true>(X, ldx, W, ldw, p, s);
    default: return launch<64, false, false, 1, true>(X, ldx, W, ldw, p, s);
  }
}

// ConvMAE stem on the VISIBLE patches only (reference: cinema/conv.py:349-415 MaskedConvBlock,
// cinema/convvit.py:165-207 DownsampleEncoder).  The reference evaluates the stem densely (NCDHW, cuDNN) on the
// full image and zeroes masked positions before the depth-wise conv (cinema/conv.py:410-411); outputs at masked
// positions are never consumed (skips and tokens are gathered with ~mask, cinema/mae/mae.py:550,
// cinema/convvit.py:288) and masked positions never influence visible ones, so evaluating only the visible ViT
// patches is exact.  Layout: token-major, channel-last rows  x[(b, i), p, c]  with i the rank of the visible token
// inside its sample and p the position inside the token's block at this stem level (row-major over the block).
// All 1x1 convs / LayerNorms / MLPs of the stem are then the same row-wise GEMM / LN kernels as the ViT blocks; this
// file holds what is left: the depth-wise 5^n convolution with neighbour lookup through the token map, its
// weight gradient, and the token -> level-position index expansion.
#include "../../include/cinema_b200.h"
#include "common.cuh"

namespace {

constexpr int KS = 5;  // depth-wise kernel size (cinema/conv.py:385)
constexpr int HALO = KS / 2;

struct DwGeom {
  int nd;           // 2 or 3
  int gt[3];        // token grid
  int f[3];         // level positions per token per axis
  int ext[3];       // f + 2*HALO  (1 for the unused third axis in 2-D)
  int taps;         // 5^nd
  int P;            // prod f
  int R;            // prod ext
  int n_tok;        // prod gt
  int B, nk, C;
};

// stage the (f + 4)^nd neighbourhood of token (b, i) into smem as [R][C] bf16 (zeros outside the image / at
// masked tokens), 16-byte vectors along the channels
__device__ __forceinline__ void stage_region(const bf16* __restrict__ in, bf16* __restrict__ tile, const DwGeom& g,
                                             const unsigned char* __restrict__ mask, const int* __restrict__ slot, int b,
                                             int t) {
  int tg[3];
  {
    int r = t;
    for (int a = g.nd - 1; a >= 0; --a) tg[a] = r % g.gt[a], r /= g.gt[a];
    for (int a = g.nd; a < 3; ++a) tg[a] = 0;
  }
  const int vec_per_row = g.C >> 3;
  for (int e = threadIdx.x; e < g.R * vec_per_row; e += blockDim.x) {
    const int r = e / vec_per_row;
    const int v = e - r * vec_per_row;
    int rr = r, tok = 0, pin = 0;
    bool ok = true;
    int ra[3];
    for (int a = 2; a >= 0; --a) ra[a] = rr % g.ext[a], rr /= g.ext[a];
    for (int a = 0; a < g.nd; ++a) {
      const int L = tg[a] * g.f[a] - HALO + ra[a];
      ok = ok && L >= 0 && L < g.gt[a] * g.f[a];
      const int ta = L / g.f[a];
      tok = tok * g.gt[a] + ta;
      pin = pin * g.f[a] + (L - ta * g.f[a]);
    }
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (ok) {
      const long long mt = (long long)b * g.n_tok + tok;
      if (mask[mt] == 0) {
        const long long row = ((long long)b * g.nk + slot[mt]) * g.P + pin;
        val = __ldg(reinterpret_cast<const uint4*>(in + row * g.C) + v);
      }
    }
    reinterpret_cast<uint4*>(tile + (long long)r * g.C)[v] = val;
  }
}

// out[(b,i), p, c] = bias[c] + sum_taps w[c, tap] * in[neighbour(p, tap), c]      (flip: transposed conv = d input)
__global__ void __launch_bounds__(256)
dwconv_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, const bf16* __restrict__ w,
              const float* __restrict__ bias, const unsigned char* __restrict__ mask, const int* __restrict__ slot,
              const int* __restrict__ keep, DwGeom g, int flip) {
  pdl_prologue();  // PDL: release the next launch, then wait for the previous kernel's results
  extern __shared__ __align__(16) uint8_t smem[];
  bf16* wsm = reinterpret_cast<bf16*>(smem);                 // [taps][C]
  bf16* tile = wsm + (size_t)g.taps * g.C;                   // [R][C]
  for (int e = threadIdx.x; e < g.taps * g.C; e += blockDim.x) {
    const int c = e / g.taps, tap = e - c * g.taps;          // global layout (C, 1, 5, 5[, 5])
    wsm[(flip ? g.taps - 1 - tap : tap) * g.C + c] = w[e];
  }
  const int c2n = g.C >> 1;
  const long long n_items = (long long)g.B * g.nk;
  for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int b = (int)(item / g.nk);
    const int t = keep[item];
    __syncthreads();  // previous tile fully consumed (and weights staged on the first pass)
    stage_region(in, tile, g, mask, slot, b, t);
    __syncthreads();
    for (int o = threadIdx.x; o < g.P * c2n; o += blockDim.x) {
      const int p = o / c2n;
      const int c2 = o - p * c2n;
      int pa[3] = {0, 0, 0};
      {
        int r = p;
        for (int a = g.nd - 1; a >= 0; --a) pa[a] = r % g.f[a], r /= g.f[a];
      }
      float2 acc = make_float2(0.f, 0.f);
      if (bias != nullptr) acc = make_float2(bias[2 * c2], bias[2 * c2 + 1]);
      int tap = 0;
      const int e2 = g.nd == 3 ? KS : 1;
      for (int d0 = 0; d0 < KS; ++d0)
        for (int d1 = 0; d1 < KS; ++d1)
          for (int d2 = 0; d2 < e2; ++d2, ++tap) {
            const int r = ((pa[0] + d0) * g.ext[1] + (pa[1] + d1)) * g.ext[2] + (pa[2] + d2);
            const float2 x = __bfloat1622float2(reinterpret_cast<const bf162*>(tile + (size_t)r * g.C)[c2]);
            const float2 ww = __bfloat1622float2(reinterpret_cast<const bf162*>(wsm + (size_t)tap * g.C)[c2]);
            acc.x = fmaf(x.x, ww.x, acc.x);
            acc.y = fmaf(x.y, ww.y, acc.y);
          }
      reinterpret_cast<bf162*>(out + (item * g.P + p) * g.C)[c2] = __floats2bfloat162_rn(acc.x, acc.y);
    }
  }
}

// dW[c, tap] += sum_rows dy[row, c] * in[neighbour(row, tap), c];   db[c] += sum_rows dy[row, c]
// thread = (channel pair, tap group); accumulators live in registers across the block's tokens.
constexpr int MAX_TAPS_PER_THREAD = 64;

__global__ void __launch_bounds__(256)
dwconv_wgrad_kernel(const bf16* __restrict__ in, const bf16* __restrict__ dy, float* __restrict__ dw,
                    float* __restrict__ db, const unsigned char* __restrict__ mask, const int* __restrict__ slot,
                    const int* __restrict__ keep, DwGeom g) {
  pdl_prologue();  // PDL: release the next launch, then wait for the previous kernel's results
  extern __shared__ __align__(16) uint8_t smem[];
  bf16* tile = reinterpret_cast<bf16*>(smem);                // [R][C]
  bf16* dys = tile + (size_t)g.R * g.C;                      // [P][C]
  const int c2n = g.C >> 1;
  const int groups = blockDim.x / c2n;                       // tap groups (host guarantees >= 1)
  const int c2 = threadIdx.x % c2n;
  const int tg = threadIdx.x / c2n;
  const bool active = tg < groups;
  float2 acc[MAX_TAPS_PER_THREAD];
#pragma unroll
  for (int i = 0; i < MAX_TAPS_PER_THREAD; ++i) acc[i] = make_float2(0.f, 0.f);
  float2 accb = make_float2(0.f, 0.f);
  const int e2 = g.nd == 3 ? KS : 1;
  const long long n_items = (long long)g.B * g.nk;
  for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int b = (int)(item / g.nk);
    const int t = keep[item];
    __syncthreads();
    stage_region(in, tile, g, mask, slot, b, t);
    for (int e = threadIdx.x; e < g.P * (g.C >> 3); e += blockDim.x)
      reinterpret_cast<uint4*>(dys)[e] = __ldg(reinterpret_cast<const uint4*>(dy + item * g.P * g.C) + e);
    __syncthreads();
    if (!active) continue;
#pragma unroll 1
    for (int p = 0; p < g.P; ++p) {
      int pa[3] = {0, 0, 0};
      {
        int r = p;
        for (int a = g.nd - 1; a >= 0; --a) pa[a] = r % g.f[a], r /= g.f[a];
      }
      const float2 d = __bfloat1622float2(reinterpret_cast<const bf162*>(dys + (size_t)p * g.C)[c2]);
      if (tg == 0) accb.x += d.x, accb.y += d.y;
#pragma unroll
      for (int i = 0; i < MAX_TAPS_PER_THREAD; ++i) {
        const int tap = tg + i * groups;
        if (tap < g.taps) {
          const int d2 = tap % e2;
          const int d1 = (tap / e2) % KS;
          const int d0 = tap / (e2 * KS);
          const int r = ((pa[0] + d0) * g.ext[1] + (pa[1] + d1)) * g.ext[2] + (pa[2] + d2);
          const float2 x = __bfloat1622float2(reinterpret_cast<const bf162*>(tile + (size_t)r * g.C)[c2]);
          acc[i].x = fmaf(d.x, x.x, acc[i].x);
          acc[i].y = fmaf(d.y, x.y, acc[i].y);
        }
      }
    }
  }
  if (!active) return;
#pragma unroll
  for (int i = 0; i < MAX_TAPS_PER_THREAD; ++i) {
    const int tap = tg + i * groups;
    if (tap < g.taps) {
      atomicAdd(dw + (size_t)(2 * c2) * g.taps + tap, acc[i].x);
      atomicAdd(dw + (size_t)(2 * c2 + 1) * g.taps + tap, acc[i].y);
    }
  }
  if (tg == 0 && db != nullptr) {
    atomicAdd(db + 2 * c2, accb.x);
    atomicAdd(db + 2 * c2 + 1, accb.y);
  }
}


// ------------------------------------------------------------------------------------------------------------
// Fast path: square in-plane token blocks F x F (x 1), F in {2, 4}, channels in groups of 64.
// One WARP per (visible token, 64-channel group): lane = channel pair.  Every neighbour row is one coalesced
// 128-byte read from L2 (the level tensor is <= 40 MB and L2-resident); the 3 x 3 in-plane neighbour tokens are
// resolved once per z-plane by nine lanes (mask + slot lookup) and broadcast with shuffles, so the inner loops
// are fully unrolled register FMAs: (F+4) row loads feed F * 25 FMAs each.
// ------------------------------------------------------------------------------------------------------------
template <int F>
struct Fast {
  static constexpr int EXT = F + 2 * HALO;
  __host__ __device__ static constexpr int off(int r) { return (r - HALO + F) / F - 1; }  // neighbour token offset
  __host__ __device__ static constexpr int pin(int r) { return (r - HALO + F) % F; }      // position inside it
};

// nb[j] (j = (o0+1)*3 + (o1+1)) = first row of neighbour token (t0+o0, t1+o1, z) in the level tensor, or -1
template <int ND>
__device__ __forceinline__ void neighbour_rows(const DwGeom& g, const unsigned char* __restrict__ mask,
                                               const int* __restrict__ slot, int b, int t0, int t1, int z, int lane,
                                               int (&nb)[9]) {
  int mine = -1;
  if (lane < 9) {
    const int a0 = t0 + lane / 3 - 1, a1 = t1 + lane % 3 - 1;
    if (a0 >= 0 && a0 < g.gt[0] && a1 >= 0 && a1 < g.gt[1]) {
      const int tok = ND == 3 ? (a0 * g.gt[1] + a1) * g.gt[2] + z : a0 * g.gt[1] + a1;
      const long long mt = (long long)b * g.n_tok + tok;
      if (mask[mt] == 0) mine = (b * g.nk + slot[mt]) * g.P;
    }
  }
#pragma unroll
  for (int j = 0; j < 9; ++j) nb[j] = __shfl_sync(0xffffffffu, mine, j);
}

template <int F>
__device__ __forceinline__ void load_row(const bf16* __restrict__ in, const int (&nb)[9], int r0, int C, int coff,
                                         float2 (&x)[F + 2 * HALO]) {
  using T = Fast<F>;
  // r0 is a compile-time constant at every call site (fully unrolled callers)
#pragma unroll
  for (int r1 = 0; r1 < T::EXT; ++r1) {
    const int base = nb[(T::off(r0) + 1) * 3 + T::off(r1) + 1];
    x[r1] = make_float2(0.f, 0.f);
    if (base >= 0) {
      const long long row = (long long)base + T::pin(r0) * F + T::pin(r1);
      x[r1] = unpack_bf16(__ldg(reinterpret_cast<const unsigned int*>(in + row * C + coff)));
    }
  }
}

template <int F, int ND>
__global__ void __launch_bounds__(128, 4)  // <= 128 registers: 16 warps per SM hide the L2 latency of the row loads
dwconv_fast_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, const bf16* __restrict__ w,
                   const float* __restrict__ bias, const unsigned char* __restrict__ mask, const int* __restrict__ slot,
                   const int* __restrict__ keep, DwGeom g, int flip) {
  pdl_prologue();  // PDL: release the next launch, then wait for the previous kernel's results
  using T = Fast<F>;
  constexpr int E2 = ND == 3 ? KS : 1;
  extern __shared__ __align__(16) uint8_t smem[];
  bf16* wsm = reinterpret_cast<bf16*>(smem);  // [taps][C], already flipped for the transposed conv
  for (int e = threadIdx.x; e < g.taps * g.C; e += blockDim.x) {
    const int c = e / g.taps, tap = e - c * g.taps;
    wsm[(flip ? g.taps - 1 - tap : tap) * g.C + c] = w[e];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  const int cgn = g.C >> 6;
  const long long n_tasks = (long long)g.B * g.nk * cgn;
  for (long long task = (long long)blockIdx.x * warps + (threadIdx.x >> 5); task < n_tasks;
       task += (long long)gridDim.x * warps) {
    const long long item = task / cgn;
    const int coff = (int)(task - item * cgn) * 64 + lane * 2;
    const int b = (int)(item / g.nk);
    int t = keep[item];
    int t2 = 0;
    if (ND == 3) t2 = t % g.gt[2], t /= g.gt[2];
    const int t1 = t % g.gt[1], t0 = t / g.gt[1];
    float2 acc[F][F];
    {
      float2 b0 = make_float2(0.f, 0.f);
      if (bias != nullptr) b0 = *reinterpret_cast<const float2*>(bias + coff);
#pragma unroll
      for (int i = 0; i < F; ++i)
#pragma unroll
        for (int j = 0; j < F; ++j) acc[i][j] = b0;
    }
#pragma unroll 1
    for (int d2 = 0; d2 < E2; ++d2) {
      const int z = t2 + d2 - (ND == 3 ? HALO : 0);
      if (ND == 3 && (z < 0 || z >= g.gt[2])) continue;
      int nb[9];
      neighbour_rows<ND>(g, mask, slot, b, t0, t1, z, lane, nb);
      float2 wr[KS * KS];
#pragma unroll
      for (int k = 0; k < KS * KS; ++k)
        wr[k] = unpack_bf16(*reinterpret_cast<const unsigned int*>(wsm + (size_t)(k * E2 + d2) * g.C + coff));
#pragma unroll
      for (int r0 = 0; r0 < T::EXT; ++r0) {
        float2 x[T::EXT];
        load_row<F>(in, nb, r0, g.C, coff, x);
#pragma unroll
        for (int d0 = 0; d0 < KS; ++d0) {
          const int i0 = r0 - d0;
          if (i0 < 0 || i0 >= F) continue;
#pragma unroll
          for (int d1 = 0; d1 < KS; ++d1)
#pragma unroll
            for (int i1 = 0; i1 < F; ++i1) {
              acc[i0][i1].x = fmaf(wr[d0 * KS + d1].x, x[i1 + d1].x, acc[i0][i1].x);
              acc[i0][i1].y = fmaf(wr[d0 * KS + d1].y, x[i1 + d1].y, acc[i0][i1].y);
            }
        }
      }
    }
#pragma unroll
    for (int i0 = 0; i0 < F; ++i0)
#pragma unroll
      for (int i1 = 0; i1 < F; ++i1)
        *reinterpret_cast<unsigned int*>(out + (item * g.P + i0 * F + i1) * g.C + coff) = pack_bf16(acc[i0][i1].x, acc[i0][i1].y);
  }
}

// Weight gradient, fast path.  A warp owns one (z-plane d2, 64-channel group) for its whole life and keeps the 25
// in-plane taps of its channel pair in registers while it strides over the visible tokens; the block then merges
// its warps in shared memory (layout == global dw) and issues coalesced fp32 reductions.
template <int F, int ND>
__global__ void __launch_bounds__(320)
dwconv_wgrad_fast_kernel(const bf16* __restrict__ in, const bf16* __restrict__ dy, float* __restrict__ dw,
                         float* __restrict__ db, const unsigned char* __restrict__ mask, const int* __restrict__ slot,
                         const int* __restrict__ keep, DwGeom g) {
  pdl_prologue();  // PDL: release the next launch, then wait for the previous kernel's results
  using T = Fast<F>;
  constexpr int E2 = ND == 3 ? KS : 1;
  extern __shared__ __align__(16) uint8_t smem[];
  float* dws = reinterpret_cast<float*>(smem);  // [C][taps]
  float* dbs = dws + (size_t)g.C * g.taps;      // [C]
  for (int e = threadIdx.x; e < g.C * g.taps + g.C; e += blockDim.x) dws[e] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int warps = blockDim.x >> 5;
  const int cgn = g.C >> 6;
  const int combos = E2 * cgn;            // host: warps % combos == 0
  const int combo = warp % combos;
  const int d2 = combo % E2;
  const int coff = (combo / E2) * 64 + lane * 2;
  const int streams_per_block = warps / combos;
  const long long n_items = (long long)g.B * g.nk;
  float2 acc[KS * KS];
#pragma unroll
  for (int k = 0; k < KS * KS; ++k) acc[k] = make_float2(0.f, 0.f);
  float2 accb = make_float2(0.f, 0.f);
  const bool bias_warp = db != nullptr && d2 == (ND == 3 ? HALO : 0);
  for (long long item = (long long)blockIdx.x * streams_per_block + warp / combos; item < n_items;
       item += (long long)gridDim.x * streams_per_block) {
    const int b = (int)(item / g.nk);
    int t = keep[item];
    int t2 = 0;
    if (ND == 3) t2 = t % g.gt[2], t /= g.gt[2];
    const int t1 = t % g.gt[1], t0 = t / g.gt[1];
    const int z = t2 + d2 - (ND == 3 ? HALO : 0);
    if (ND == 3 && (z < 0 || z >= g.gt[2])) continue;
    int nb[9];
    neighbour_rows<ND>(g, mask, slot, b, t0, t1, z, lane, nb);
    float2 d[F][F];
#pragma unroll
    for (int i0 = 0; i0 < F; ++i0)
#pragma unroll
      for (int i1 = 0; i1 < F; ++i1)
        d[i0][i1] = unpack_bf16(__ldg(reinterpret_cast<const unsigned int*>(dy + (item * g.P + i0 * F + i1) * g.C + coff)));
    if (bias_warp) {
#pragma unroll
      for (int i0 = 0; i0 < F; ++i0)
#pragma unroll
        for (int i1 = 0; i1 < F; ++i1) accb.x += d[i0][i1].x, accb.y += d[i0][i1].y;
    }
#pragma unroll
    for (int r0 = 0; r0 < T::EXT; ++r0) {
      float2 x[T::EXT];
      load_row<F>(in, nb, r0, g.C, coff, x);
#pragma unroll
      for (int d0 = 0; d0 < KS; ++d0) {
        const int i0 = r0 - d0;
        if (i0 < 0 || i0 >= F) continue;
#pragma unroll
        for (int d1 = 0; d1 < KS; ++d1)
#pragma unroll
          for (int i1 = 0; i1 < F; ++i1) {
            acc[d0 * KS + d1].x = fmaf(d[i0][i1].x, x[i1 + d1].x, acc[d0 * KS + d1].x);
            acc[d0 * KS + d1].y = fmaf(d[i0][i1].y, x[i1 + d1].y, acc[d0 * KS + d1].y);
          }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < KS * KS; ++k) {
    atomicAdd(dws + (size_t)coff * g.taps + k * E2 + d2, acc[k].x);
    atomicAdd(dws + (size_t)(coff + 1) * g.taps + k * E2 + d2, acc[k].y);
  }
  if (bias_warp) {
    atomicAdd(dbs + coff, accb.x);
    atomicAdd(dbs + coff + 1, accb.y);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < g.C * g.taps; e += blockDim.x) atomicAdd(dw + e, dws[e]);
  if (db != nullptr)
    for (int e = threadIdx.x; e < g.C; e += blockDim.x) atomicAdd(db + e, dbs[e]);
}

// 0 = not covered by the fast path
inline int fast_f(const DwGeom& g) {
  if (g.C % 64 != 0 || g.f[0] != g.f[1] || g.f[2] != 1) return 0;
  return (g.f[0] == 2 || g.f[0] == 4) ? g.f[0] : 0;
}

template <int F, int ND>
int launch_fast_fwd(const void* in, void* out, const void* w, const float* bias, const unsigned char* mask,
                    const int* slot, const int* keep, const DwGeom& g, int transpose, cudaStream_t stream) {
  const size_t smem = (size_t)g.taps * g.C * sizeof(bf16);
  auto kern = dwconv_fast_kernel<F, ND>;
  static size_t configured = 0;
  if (smem > configured) {
    CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const long long tasks = (long long)g.B * g.nk * (g.C / 64);
  long long blocks = (tasks + 3) / 4;
  const long long cap = (long long)cb_sm_count() * 4;
  if (blocks > cap) blocks = cap;
  cb_launch(kern, (unsigned)blocks, 128, smem, stream, (const bf16*)in, (bf16*)out, (const bf16*)w, bias, mask, slot, keep, g,
                                               transpose);
  CB_LAUNCH_CHECK();
  return 0;
}

template <int F, int ND>
int launch_fast_wgrad(const void* in, const void* dy, float* dw, float* db, const unsigned char* mask, const int* slot,
                      const int* keep, const DwGeom& g, cudaStream_t stream) {
  const int combos = (ND == 3 ? KS : 1) * (g.C / 64);
  if (combos > 10) return -2;  // caller falls back to the generic kernel
  const int warps = combos * (10 / combos);
  const size_t smem = ((size_t)g.C * g.taps + g.C) * sizeof(float);
  if (smem > 200 * 1024) return -2;
  auto kern = dwconv_wgrad_fast_kernel<F, ND>;
  static size_t configured = 0;
  if (smem > configured) {
    CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const long long items = (long long)g.B * g.nk;
  const int spb = warps / combos;
  long long blocks = (items + spb - 1) / spb;
  const long long cap = (long long)cb_sm_count() * 2;
  if (blocks > cap) blocks = cap;
  cb_launch(kern, (unsigned)blocks, warps * 32, smem, stream, (const bf16*)in, (const bf16*)dy, dw, db, mask, slot, keep, g);
  CB_LAUNCH_CHECK();
  return 0;
}

// out[b, i*P + p] = flattened position id, in the level grid (gt * f), of position p of visible token keep[b, i]
__global__ void expand_index_kernel(const int* __restrict__ keep, long long n_items, DwGeom g, int* __restrict__ out) {
  pdl_prologue();  // PDL: release the next launch, then wait for the previous kernel's results
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_items * g.P) return;
  const long long item = e / g.P;
  int p = (int)(e - item * g.P);
  int t = keep[item];
  int tg[3], pa[3];
  for (int a = g.nd - 1; a >= 0; --a) {
    tg[a] = t % g.gt[a], t /= g.gt[a];
    pa[a] = p % g.f[a], p /= g.f[a];
  }
  int id = 0;
  for (int a = 0; a < g.nd; ++a) id = id * (g.gt[a] * g.f[a]) + tg[a] * g.f[a] + pa[a];
  out[e] = id;
}

int fill(DwGeom& g, int B, int nk, int C, int ndim, const int* grid_tok, const int* f) {
  CB_CHECK_ARG(ndim == 2 || ndim == 3, "stem: ndim %d must be 2 or 3", ndim);
  g.nd = ndim, g.B = B, g.nk = nk, g.C = C, g.P = 1, g.R = 1, g.n_tok = 1, g.taps = 1;
  for (int a = 0; a < 3; ++a) g.gt[a] = g.f[a] = g.ext[a] = 1;
  for (int a = 0; a < ndim; ++a) {
    CB_CHECK_ARG(grid_tok[a] > 0 && f[a] > 0, "stem: bad geometry");
    g.gt[a] = grid_tok[a], g.f[a] = f[a], g.ext[a] = f[a] + 2 * HALO;
    g.P *= f[a], g.R *= g.ext[a], g.n_tok *= grid_tok[a], g.taps *= KS;
  }
  return 0;
}

}  // namespace

extern "C" int cb_expand_token_index(const int* keep, int B, int nk, int ndim, const int* grid_tok, const int* f,
                                     int* out, void* stream) {
  DwGeom g;
  if (int rc = fill(g, B, nk, 8, ndim, grid_tok, f)) return rc;
  const long long total = (long long)B * nk * g.P;
  if (total <= 0) return 0;
  cb_launch(expand_index_kernel, (unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream, keep, (long long)B * nk, g, out);
  CB_LAUNCH_CHECK();
  return 0;
}

extern "C" int cb_dwconv_tokens(const void* in, void* out, const void* w, const float* bias, const unsigned char* mask,
                                const int* slot, const int* keep, int B, int nk, int C, int ndim, const int* grid_tok,
                                const int* f, int transpose, void* stream) {
  DwGeom g;
  if (int rc = fill(g, B, nk, C, ndim, grid_tok, f)) return rc;
  if ((long long)B * nk <= 0) return 0;
  CB_CHECK_ARG(C % 8 == 0, "dwconv: C=%d must be a multiple of 8", C);
  if (const int ff = fast_f(g)) {
    cudaStream_t st = (cudaStream_t)stream;
    if (ff == 4) return ndim == 3 ? launch_fast_fwd<4, 3>(in, out, w, bias, mask, slot, keep, g, transpose, st)
                                  : launch_fast_fwd<4, 2>(in, out, w, bias, mask, slot, keep, g, transpose, st);
    return ndim == 3 ? launch_fast_fwd<2, 3>(in, out, w, bias, mask, slot, keep, g, transpose, st)
                     : launch_fast_fwd<2, 2>(in, out, w, bias, mask, slot, keep, g, transpose, st);
  }
  const size_t smem = ((size_t)g.taps + g.R) * C * sizeof(bf16);
  CB_CHECK_ARG(smem <= 220 * 1024, "dwconv: tile of %zu bytes does not fit in shared memory", smem);
  static size_t configured = 0;
  if (smem > configured) {
    CB_CUDA(cudaFuncSetAttribute(dwconv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const long long items = (long long)B * nk;
  const int blocks = (int)(items < (long long)cb_sm_count() * 4 ? items : (long long)cb_sm_count() * 4);
  cb_launch(dwconv_kernel, blocks, 256, smem, (cudaStream_t)stream, (const bf16*)in, (bf16*)out, (const bf16*)w, bias, mask, slot,
                                                            keep, g, transpose);
  CB_LAUNCH_CHECK();
  return 0;
}

extern "C" int cb_dwconv_tokens_wgrad(const void* in, const void* dy, float* dw, float* db, const unsigned char* mask,
                                      const int* slot, const int* keep, int B, int nk, int C, int ndim,
                                      const int* grid_tok, const int* f, void* stream) {
  DwGeom g;
  if (int rc = fill(g, B, nk, C, ndim, grid_tok, f)) return rc;
  if ((long long)B * nk <= 0) return 0;
  CB_CHECK_ARG(C % 8 == 0 && C <= 512, "dwconv_wgrad: C=%d must be a multiple of 8 and <= 512", C);
  if (const int ff = fast_f(g)) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if (ff == 4) rc = ndim == 3 ? launch_fast_wgrad<4, 3>(in, dy, dw, db, mask, slot, keep, g, st)
                                : launch_fast_wgrad<4, 2>(in, dy, dw, db, mask, slot, keep, g, st);
    else rc = ndim == 3 ? launch_fast_wgrad<2, 3>(in, dy, dw, db, mask, slot, keep, g, st)
                        : launch_fast_wgrad<2, 2>(in, dy, dw, db, mask, slot, keep, g, st);
    if (rc != -2) return rc;
  }
  const int c2n = C / 2;
  const int groups = 256 / c2n;
  CB_CHECK_ARG(groups >= 1 && (g.taps + groups - 1) / groups <= MAX_TAPS_PER_THREAD,
               "dwconv_wgrad: %d taps over %d groups exceed the per-thread accumulator budget", g.taps, groups);
  const size_t smem = ((size_t)g.R + g.P) * C * sizeof(bf16);
  CB_CHECK_ARG(smem <= 220 * 1024, "dwconv_wgrad: tile of %zu bytes does not fit in shared memory", smem);
  static size_t configured = 0;
  if (smem > configured) {
    CB_CUDA(cudaFuncSetAttribute(dwconv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const long long items = (long long)B * nk;
  const int blocks = (int)(items < (long long)cb_sm_count() * 2 ? items : (long long)cb_sm_count() * 2);
  cb_launch(dwconv_wgrad_kernel, blocks, 256, smem, (cudaStream_t)stream, (const bf16*)in, (const bf16*)dy, dw, db, mask, slot,
                                                                  keep, g);
  CB_LAUNCH_CHECK();
  return 0;
}

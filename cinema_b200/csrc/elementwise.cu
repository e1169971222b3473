// HBM-bound data-movement kernels of the MAE path: weight cast, mask -> index lists,
// row gather / scatter, decoder-embedding assembly, N-d patchify / unpatchify, the fused
// "patchify + visible-token gather", and bias-gradient column sums.
// All copies are bit-exact; all accesses are coalesced 16-byte vectors where the layout allows.
#include "../../include/cinema_b200.h"
#include "common.cuh"

namespace {

constexpr int MAX_ND = 4;

struct NdGeom {
  int ndim;
  int C;
  int spatial[MAX_ND];
  int patch[MAX_ND];
  int grid[MAX_ND];
  long long sb, sc;            // source batch / channel strides (elements)
  long long sstride[MAX_ND];   // source spatial strides (elements)
  int n_tok;                   // prod(grid)
  int E;                       // prod(patch) * C
};

// ---------------------------------------------------------------------------------------
// fp32 -> bf16 flat cast
// ---------------------------------------------------------------------------------------
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n) {
  pdl_prologue();  // PDL: release the next launch, then wait for the previous kernel's results
  const long long n8 = n >> 3;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src) + 2 * i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 2 * i + 1);
    reinterpret_cast<uint4*>(dst)[i] =
        make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {
    const long long i = (n8 << 3) + threadIdx.x;
    dst[i] = __float2bfloat16_rn(src[i]);
  }
}

// ---------------------------------------------------------------------------------------
// mask -> ascending keep / drop index lists (one warp per batch row, ballot compaction)
// ---------------------------------------------------------------------------------------
__global__ void mask_to_index_kernel(const unsigned char* __restrict__ mask, int B, int n, int n_keep,
                                     int* __restrict__ keep_idx, int* __restrict__ drop_idx, int* __restrict__ slot) {
  pdl_prologue();  // PDL: release the next launch, then wait for the previous kernel's results
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const unsigned char* m = mask + (long long)b * n;
  const int n_drop = n - n_keep;
  int nk = 0, nd = 0;
  for (int base = 0; base < n; base += 32) {
    const int t = base + lane;
    const bool valid = t < n;
    const bool removed = valid && m[t] != 0;
    const unsigned keep_bits = __ballot_sync(0xffffffffu, valid && !removed);
    const unsigned drop_bits = __ballot_sync(0xffffffffu, removed);
    const unsigned below = (1u << lane) - 1u;
    if (valid) {
      if (!removed) {
        const int pos = nk + __popc(keep_bits & below);
        if (pos < n_keep && keep_idx) keep_idx[(long long)b * n_keep + pos] = t;
        if (slot) slot[(long long)b * n + t] = pos;
      } else {
        const int pos = nd + __popc(drop_bits & below);
        if (pos < n_drop && drop_idx) drop_idx[(long long)b * n_drop + pos] = t;
        if (slot) slot[(long long)b * n + t] = pos;
      }
    }
    nk += __popc(keep_bits);
    nd += __popc(drop_bits);
  }
}

// ---------------------------------------------------------------------------------------
// row gather / scatter by index, 16-byte vectors, one warp per row
// ---------------------------------------------------------------------------------------
template <bool SCATTER>
__global__ void move_rows_kernel(const uint4* __restrict__ src, long long src_bstride, long long src_off,
                                 const int* __restrict__ idx, int B, int k, uint4* __restrict__ dst,
                                 long long dst_bstride, long long dst_off, int vec_per_row) {
  pdl_prologue();  // PDL: release the next launch, then wait for the previous kernel's results
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= (long long)B * k) return;
  const int b = (int)(row / k);
  const int i = (int)(row - (long long)b * k);
  const int t = idx[row];
  long long s_row, d_row;
  if (!SCATTER) {
    s_row = b * src_bstride + t;
    d_row = b * dst_bstride + dst_off + i;
  } else {
    s_row = b * src_bstride + src_off + i;
    d_row = b * dst_bstride + t;
  }
  const uint4* s = src + s_row * vec_per_row;
  uint4* d = dst + d_row * vec_per_row;
  for (int v = lane; v < vec_per_row; v += 32) d[v] = __ldg(s + v);
}

__global__ void embed_rows_kernel(const float4* __restrict__ a, long long a_bstride, long long a_off,
                                  const float4* __restrict__ rowv, const float4* __restrict__ table,
                                  const int* __restrict__ idx, int B, int k, int vec_per_row, float4* __restrict__ out,
                                  uint2* __restrict__ out16, long long out_bstride, long long out_off) {
  pdl_prologue();  // PDL: release the next launch, then wait for the previous kernel's results
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= (long long)B * k) return;
  const int b = (int)(row / k);
  const int i = (int)(row - (long long)b * k);
  const float4* t = table ? table + (long long)idx[row] * vec_per_row : nullptr;
  const float4* ap = a ? a + (b * a_bstride + a_off + i) * vec_per_row : nullptr;
  const long long o_row = (b * out_bstride + out_off + i) * vec_per_row;
  for (int v = lane; v < vec_per_row; v += 32) {
    float4 r = t ? __ldg(t + v) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (ap) {
      const float4 x = __ldg(ap + v);
      r.x += x.x, r.y += x.y, r.z += x.z, r.w += x.w;
    }
    if (rowv) {
      const float4 x = __ldg(rowv + v);
      r.x += x.x, r.y += x.y, r.z += x.z, r.w += x.w;
    }
    if (out) out[o_row + v] = r;
    if (out16) out16[o_row + v] = make_uint2(pack_bf16(r.x, r.y), pack_bf16(r.z, r.w));
  }
}

// out[d] += sum over (b, i<k) of X[b*bstride + off + i, d]   (fp32; token / bias style gradients)
__global__ void colsum_seg_f32_kernel(const float* __restrict__ X, long long bstride, long long off, int B, int k, int D,
                                      float* __restrict__ out, int rows_per_block) {
  pdl_prologue();  // PDL: release the next launch, then wait for the previous kernel's results
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= D) return;
  const long long total = (long long)B * k;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = min(r0 + rows_per_block, total);
  float acc = 0.f;
  for (long long r = r0; r < r1; ++r) {
    const long long b = r / k;
    const long long i = r - b * k;
    acc += __ldg(X + (b * bstride + off + i) * D + col);
  }
  atomicAdd(out + col, acc);
}

// dst = bf16(src * scale_host * (scale_dev ? scale_dev[g] : 1)), g = element / group (group == 0: one scalar, g = 0;
// otherwise group is a multiple of 4 elements: per-sample factors of stochastic depth)
__global__ void scale_cast_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n,
                                  const float* __restrict__ scale_dev, float scale, long long group) {
  pdl_prologue();  // PDL: release the next launch, then wait for the previous kernel's results
  const float sc0 = scale * ((scale_dev && group == 0) ? __ldg(scale_dev) : 1.0f);
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src) + i);
    const float sc = group > 0 ? sc0 * __ldg(scale_dev + (i << 2) / group) : sc0;
    reinterpret_cast<uint2*>(dst)[i] = make_uint2(pack_bf16(a.x * sc, a.y * sc), pack_bf16(a.z * sc, a.w * sc));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = (n4 << 2) + threadIdx.x;
    dst[i] = __float2bfloat16_rn(src[i] * sc0);  // (a ragged tail only exists in scalar mode: group % 4 == 0 divides n)
  }
}

// ---------------------------------------------------------------------------------------
// ScaleIntensity of a raw-dtype batch: 16 raw bytes per thread and step, fp32 out as 16-byte vectors
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void scale_intensity_kernel(const T* __restrict__ raw, const float* __restrict__ lo, const float* __restrict__ hi,
                                       float* __restrict__ out, long long n_per_sample) {
  pdl_prologue();  // PDL: release the next launch, then wait for the previous kernel's results
  constexpr int V = 16 / (int)sizeof(T);  // elements per 16-byte load
  const int b = blockIdx.y;
  const float l = __ldg(lo + b), span = __ldg(hi + b) - l;
  const float inv = span > 0.f ? 1.0f / span : 0.f;
  const T* src = raw + (long long)b * n_per_sample;
  float* dst = out + (long long)b * n_per_sample;
  const long long nv = n_per_sample / V;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(src) + i);
    const T* e = reinterpret_cast<const T*>(&u);
    float4* o = reinterpret_cast<float4*>(dst + i * V);
#pragma unroll
    for (int k = 0; k < V / 4; ++k)
      o[k] = make_float4((static_cast<float>(e[4 * k]) - l) * inv, (static_cast<float>(e[4 * k + 1]) - l) * inv,
                         (static_cast<float>(e[4 * k + 2]) - l) * inv, (static_cast<float>(e[4 * k + 3]) - l) * inv);
  }
  if (blockIdx.x == 0) {  // ragged tail of a sample whose size is not a multiple of V
    for (long long i = nv * V + threadIdx.x; i < n_per_sample; i += blockDim.x)
      dst[i] = (static_cast<float>(src[i]) - l) * inv;
  }
}

// ---------------------------------------------------------------------------------------
// patchify / unpatchify: one thread per IMAGE element (image side is the contiguous stream)
// ---------------------------------------------------------------------------------------
template <typename T, bool INVERSE>
__global__ void patchify_kernel(const T* __restrict__ src, T* __restrict__ dst, NdGeom g, long long n_per_batch,
                                long long total) {
  pdl_prologue();  // PDL: release the next launch, then wait for the previous kernel's results
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const long long b = e / n_per_batch;
    long long r = e - b * n_per_batch;  // offset inside (C, S1..Sn)
    int coord[MAX_ND];
#pragma unroll
    for (int a = MAX_ND - 1; a >= 0; --a) {
      if (a < g.ndim) {
        coord[a] = (int)(r % g.spatial[a]);
        r /= g.spatial[a];
      }
    }
    const int c = (int)r;
    int tok = 0, off = 0;
#pragma unroll
    for (int a = 0; a < MAX_ND; ++a) {
      if (a < g.ndim) {
        tok = tok * g.grid[a] + coord[a] / g.patch[a];
        off = off * g.patch[a] + coord[a] % g.patch[a];
      }
    }
    const long long t = (b * g.n_tok + tok) * (long long)g.E + (long long)off * g.C + c;
    if constexpr (!INVERSE)
      dst[t] = src[e];
    else
      dst[e] = src[t];
  }
}

// ---------------------------------------------------------------------------------------
// gather patches of selected tokens from a strided source; one thread per output element
// ---------------------------------------------------------------------------------------
template <typename TS, typename TR, bool SCATTER>
__global__ void patches_kernel(TS* __restrict__ img, TR* __restrict__ rows, NdGeom g, const int* __restrict__ idx,
                               int k, int chan_last, long long total, int accumulate) {
  pdl_prologue();  // PDL: release the next launch, then wait for the previous kernel's results
  const long long stride = (long long)gridDim.x * blockDim.x;
  int pprod = 1;
#pragma unroll
  for (int a = 0; a < MAX_ND; ++a)
    if (a < g.ndim) pprod *= g.patch[a];
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const long long row = e / g.E;
    const int el = (int)(e - row * g.E);
    const int b = (int)(row / k);
    int tok = idx ? idx[row] : (int)(row - (long long)b * k);
    int c, off;
    if (chan_last) {
      c = el % g.C;
      off = el / g.C;
    } else {
      c = el / pprod;
      off = el - c * pprod;
    }
    long long s = b * g.sb + c * g.sc;
#pragma unroll
    for (int a = MAX_ND - 1; a >= 0; --a) {
      if (a < g.ndim) {
        const int ga = tok % g.grid[a];
        tok /= g.grid[a];
        const int oa = off % g.patch[a];
        off /= g.patch[a];
        s += (long long)(ga * g.patch[a] + oa) * g.sstride[a];
      }
    }
    if constexpr (!SCATTER)
      rows[e] = static_cast<TR>(static_cast<float>(img[s]));
    else if (accumulate)  // patches are disjoint: exactly one thread touches each destination element per call
      img[s] = static_cast<TS>(static_cast<float>(img[s]) + static_cast<float>(rows[e]));
    else
      img[s] = static_cast<TS>(static_cast<float>(rows[e]));
  }
}

// ---------------------------------------------------------------------------------------
// column sums of a bf16 matrix (bias gradient): block = 8 warps x 64 columns, rows strided
// ---------------------------------------------------------------------------------------
__global__ void colsum_bf16_kernel(const bf16* __restrict__ X, long long ldx, int M, int N, float* __restrict__ out,
                                   int rows_per_block) {
  pdl_prologue();  // PDL: release the next launch, then wait for the previous kernel's results
  __shared__ float2 part[8][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col = blockIdx.x * 64 + lane * 2;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(r0 + rows_per_block, M);
  float2 acc = make_float2(0.f, 0.f);
  if (col < N) {
    for (int r = r0 + warp; r < r1; r += 8) {
      const float2 v = unpack_bf16(__ldg(reinterpret_cast<const uint32_t*>(X + (long long)r * ldx + col)));
      acc.x += v.x, acc.y += v.y;
    }
  }
  part[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && col < N) {
    float2 s = part[0][lane];
#pragma unroll
    for (int w = 1; w < 8; ++w) s.x += part[w][lane].x, s.y += part[w][lane].y;
    atomicAdd(out + col, s.x);
    atomicAdd(out + col + 1, s.y);
  }
}

// NeoX-style rotary embedding over the first `ro` of `d` channels (cinema/rotary.py:12-50):
//   y[..., j]        = x[j] cos_j - x[j + ro/2] sin_j      j < ro/2
//   y[..., j + ro/2] = x[j + ro/2] cos_j + x[j] sin_j
// rows are (b, token, head); cos / sin are (n_tokens, ro/2) fp32 tables.  sin_sign = -1 gives the transpose (backward).
template <typename T>
__global__ void rope_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ cos_t,
                            const float* __restrict__ sin_t, long long rows, int n_tokens, int H, int d, int ro,
                            float sin_sign) {
  pdl_prologue();  // PDL: release the next launch, then wait for the previous kernel's results
  const int half = ro >> 1;
  const long long total = rows * d;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const long long row = e / d;
    const int c = (int)(e - row * d);
    const float v = static_cast<float>(x[e]);
    if (c >= ro) {
      y[e] = static_cast<T>(v);
      continue;
    }
    const int tok = (int)((row / H) % n_tokens);
    const int j = c < half ? c : c - half;
    const float cs = __ldg(cos_t + (long long)tok * half + j);
    const float sn = __ldg(sin_t + (long long)tok * half + j) * sin_sign;
    const float other = static_cast<float>(c < half ? x[e + half] : x[e - half]);
    y[e] = static_cast<T>(c < half ? v * cs - other * sn : v * cs + other * sn);
  }
}

int fill_geom(NdGeom& g, int C, int ndim, const int* spatial_or_grid, bool is_grid, const int* patch) {
  CB_CHECK_ARG(ndim >= 1 && ndim <= MAX_ND, "patch geometry: ndim %d not in 1..4", ndim);
  g.ndim = ndim, g.C = C, g.n_tok = 1, g.E = C;
  for (int a = 0; a < MAX_ND; ++a) g.spatial[a] = g.patch[a] = g.grid[a] = 1, g.sstride[a] = 0;
  for (int a = 0; a < ndim; ++a) {
    CB_CHECK_ARG(patch[a] > 0, "patch size must be positive");
    g.patch[a] = patch[a];
    if (is_grid) {
      g.grid[a] = spatial_or_grid[a];
      g.spatial[a] = g.grid[a] * patch[a];
    } else {
      g.spatial[a] = spatial_or_grid[a];
      CB_CHECK_ARG(g.spatial[a] % patch[a] == 0, "Input size (%d) cannot be divided by patch size (%d).", g.spatial[a],
                   patch[a]);
      g.grid[a] = g.spatial[a] / patch[a];
    }
    g.n_tok *= g.grid[a];
    g.E *= patch[a];
  }
  long long s = 1;  // contiguous default strides
  for (int a = ndim - 1; a >= 0; --a) g.sstride[a] = s, s *= g.spatial[a];
  g.sc = s, g.sb = s * C;
  return 0;
}

// ---------------------------------------------------------------------------------------
// Token-major patch rows (what the native conv stem produces and consumes): the source is a stack of per-token blocks
// [T][P positions][C channels] fp32, channel contiguous, and a patch row is a permutation of (part of) one block:
//   gather : rows[t * R + r, :] (bf16) = cast(block_t[perm])        R = prod(grid) rows per token
//   scatter: block_t (fp32) (+)= rows[t * R + r, :][perm^-1]
// One WARP per token, the block staged through shared memory by 1-D bulk copies (cp.async.bulk -> UBLKCP; per-warp
// double buffering on two mbarriers): the global side moves whole contiguous 2-4 KB blocks at full sector efficiency,
// the permutation (channel-last <-> channel-first order inside a row, sub-patch rows) happens between shared memory
// and registers, rows leave as 16-byte vectors; the scatter converts into a second staging tile in block order and
// hands it to the bulk engine as ONE fp32 reduce-add (or store) per token -- no read-modify-write of the gradient
// block by the SM.  Replaces the element-per-thread generic kernel (eight 64-bit divisions per element, 2-byte
// stores: 0.36-0.45 TB/s) for every stem / fusion / patch-embedding gather of the MAE step
// (cinema/vit.py:67-142 patchify + cinema/mae/mae.py:550 gather, restated on visible tokens).
// ---------------------------------------------------------------------------------------
constexpr int TOK_MAX_BLOCK = 2048;  // elements per token block (8 KB fp32)
constexpr int TOK_WARPS = 4;

struct TokGeom {
  int T;              // tokens
  int blk;            // elements per token block == R * E
  int E_shift;        // log2(E): elements per row
  int C_shift;        // log2(C)
  int pp_shift;       // log2(prod(patch))
  int chan_last;      // row element order: (position, channel) or (channel, position)
  long long tok_stride;  // source elements between consecutive token blocks (>= blk)
  short row_org[16];  // block offset (elements) of row r's first position, channel 0
  short pos_off[64];  // block offset of patch position `off` relative to its row origin
};

template <typename TSRC>
__device__ __forceinline__ int tok_src_index(const TokGeom& g, int o) {
  const int r = o >> g.E_shift;
  const int el = o & ((1 << g.E_shift) - 1);
  int c, off;
  if (g.chan_last) {
    off = el >> g.C_shift;
    c = el & ((1 << g.C_shift) - 1);
  } else {
    c = el >> g.pp_shift;
    off = el & ((1 << g.pp_shift) - 1);
  }
  return g.row_org[r] + g.pos_off[off] + c;
}

__global__ void __launch_bounds__(TOK_WARPS * 32)
tok_gather_kernel(const float* __restrict__ src, bf16* __restrict__ rows, const __grid_constant__ TokGeom g) {
  pdl_prologue();  // PDL: release the next launch, then wait for the previous kernel's results
  extern __shared__ __align__(128) uint8_t tok_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* slot[2] = {reinterpret_cast<float*>(tok_smem) + (warp * 2 + 0) * TOK_MAX_BLOCK,
                    reinterpret_cast<float*>(tok_smem) + (warp * 2 + 1) * TOK_MAX_BLOCK};
  uint64_t* bars = reinterpret_cast<uint64_t*>(tok_smem + TOK_WARPS * 2 * TOK_MAX_BLOCK * 4) + warp * 2;
  if (lane == 0) {
    mbar_init(&bars[0], 1), mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  __syncwarp();
  const int stride = gridDim.x * TOK_WARPS;
  int t = blockIdx.x * TOK_WARPS + warp;
  const uint32_t bytes = (uint32_t)g.blk * 4u;
  if (t < g.T && lane == 0) {
    mbar_expect_tx(&bars[0], bytes);
    bulk_load_1d(slot[0], src + (long long)t * g.tok_stride, bytes, &bars[0]);
  }
  for (int it = 0; t < g.T; t += stride, ++it) {
    const int cur = it & 1;
    if (t + stride < g.T && lane == 0) {  // prefetch the next token of this warp into the other slot
      mbar_expect_tx(&bars[cur ^ 1], bytes);
      bulk_load_1d(slot[cur ^ 1], src + (long long)(t + stride) * g.tok_stride, bytes, &bars[cur ^ 1]);
    }
    mbar_wait(&bars[cur], (it >> 1) & 1);
    const float* blk = slot[cur];
    bf16* out = rows + (long long)t * g.blk;
    for (int o = lane * 8; o < g.blk; o += 256) {
      float v[8];
      if (g.chan_last) {  // eight consecutive channels of one position: contiguous in the block
        const float4* s4 = reinterpret_cast<const float4*>(blk + tok_src_index<float>(g, o));
        const float4 a = s4[0], b = s4[1];
        v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = blk[tok_src_index<float>(g, o + j)];
      }
      *reinterpret_cast<uint4*>(out + o) =
          make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    }
    __syncwarp();  // every lane is done with this slot before it is refilled two iterations later
  }
}

template <typename TROW>
__global__ void __launch_bounds__(TOK_WARPS * 32)
tok_scatter_kernel(const TROW* __restrict__ rows, float* __restrict__ dst, const __grid_constant__ TokGeom g,
                   int accumulate) {
  pdl_prologue();  // PDL: release the next launch, then wait for the previous kernel's results
  extern __shared__ __align__(128) uint8_t tok_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // per warp: two row slots (raw row bytes) + one fp32 staging tile in block order
  uint8_t* base = tok_smem + warp * (3 * TOK_MAX_BLOCK * 4);
  TROW* slot[2] = {reinterpret_cast<TROW*>(base), reinterpret_cast<TROW*>(base + TOK_MAX_BLOCK * 4)};
  float* stage = reinterpret_cast<float*>(base + 2 * TOK_MAX_BLOCK * 4);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tok_smem + TOK_WARPS * 3 * TOK_MAX_BLOCK * 4) + warp * 2;
  if (lane == 0) {
    mbar_init(&bars[0], 1), mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  __syncwarp();
  const int stride = gridDim.x * TOK_WARPS;
  int t = blockIdx.x * TOK_WARPS + warp;
  const uint32_t in_bytes = (uint32_t)g.blk * (uint32_t)sizeof(TROW);
  const uint32_t out_bytes = (uint32_t)g.blk * 4u;
  if (t < g.T && lane == 0) {
    mbar_expect_tx(&bars[0], in_bytes);
    bulk_load_1d(slot[0], rows + (long long)t * g.blk, in_bytes, &bars[0]);
  }
  for (int it = 0; t < g.T; t += stride, ++it) {
    const int cur = it & 1;
    if (t + stride < g.T && lane == 0) {
      mbar_expect_tx(&bars[cur ^ 1], in_bytes);
      bulk_load_1d(slot[cur ^ 1], rows + (long long)(t + stride) * g.blk, in_bytes, &bars[cur ^ 1]);
    }
    mbar_wait(&bars[cur], (it >> 1) & 1);
    if (lane == 0) bulk_wait_group_read0();  // the previous token's bulk write has finished reading the staging tile
    __syncwarp();
    const TROW* rw = slot[cur];
    for (int o = lane * 8; o < g.blk; o += 256) {
      float v[8];
      if constexpr (sizeof(TROW) == 2) {
        const uint4 u = *reinterpret_cast<const uint4*>(rw + o);
        const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
        v[0] = a.x, v[1] = a.y, v[2] = b.x, v[3] = b.y, v[4] = c.x, v[5] = c.y, v[6] = d.x, v[7] = d.y;
      } else {
        const float4 a = reinterpret_cast<const float4*>(rw + o)[0], b = reinterpret_cast<const float4*>(rw + o)[1];
        v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
      }
      if (g.chan_last) {
        float4* d4 = reinterpret_cast<float4*>(stage + tok_src_index<float>(g, o));
        d4[0] = make_float4(v[0], v[1], v[2], v[3]);
        d4[1] = make_float4(v[4], v[5], v[6], v[7]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) stage[tok_src_index<float>(g, o + j)] = v[j];
      }
    }
    fence_proxy_async_smem();  // staging tile (generic-proxy writes) -> visible to the bulk engine
    __syncwarp();
    if (lane == 0) {
      float* d = dst + (long long)t * g.tok_stride;
      if (accumulate) bulk_reduce_add_f32_1d(d, stage, out_bytes);
      else bulk_store_1d(d, stage, out_bytes);
      bulk_commit_group();
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // writes complete before the CTA retires
}

// Is (g, idx) the token-major case?  Fills tg when it is.
bool token_major_geom(const NdGeom& g, int B, const int* idx, int chan_last, const void* src, TokGeom& tg) {
  if (idx != nullptr || g.sc != 1 || g.ndim < 1) return false;
  int R = 1, pp = 1;
  for (int a = 0; a < g.ndim; ++a) R *= g.grid[a], pp *= g.patch[a];
  const int E = pp * g.C, blk = R * E;
  auto pow2 = [](int x) { return x > 0 && (x & (x - 1)) == 0; };
  if (!pow2(E) || !pow2(g.C) || !pow2(pp) || g.C % 8 != 0 || R > 16 || pp > 64 || blk > TOK_MAX_BLOCK || blk % 256 != 0)
    return false;
  if (g.sb < blk || (g.sb * 4) % 16 != 0 || ((uintptr_t)src & 15) != 0) return false;
  auto lg = [](int x) { int s = 0; while ((1 << s) < x) ++s; return s; };
  tg.T = B, tg.blk = blk, tg.E_shift = lg(E), tg.C_shift = lg(g.C), tg.pp_shift = lg(pp), tg.chan_last = chan_last;
  tg.tok_stride = g.sb;
  long long max_off = 0;
  for (int r = 0; r < R; ++r) {  // row r = grid coordinate (row-major over the grid dims)
    int rem = r;
    long long o = 0;
    for (int a = g.ndim - 1; a >= 0; --a) {
      o += (long long)(rem % g.grid[a]) * g.patch[a] * g.sstride[a];
      rem /= g.grid[a];
    }
    if (o > 32767) return false;
    tg.row_org[r] = (short)o;
    max_off = o > max_off ? o : max_off;
  }
  long long max_pos = 0;
  for (int f = 0; f < pp; ++f) {  // patch position (row-major over the patch dims)
    int rem = f;
    long long o = 0;
    for (int a = g.ndim - 1; a >= 0; --a) {
      o += (long long)(rem % g.patch[a]) * g.sstride[a];
      rem /= g.patch[a];
    }
    if (o > 32767) return false;
    tg.pos_off[f] = (short)o;
    max_pos = o > max_pos ? o : max_pos;
  }
  // the addressed elements must stay inside one contiguous block of blk elements, and position offsets must be
  // multiples of 8 elements for the vector path (they are multiples of C)
  if (max_off + max_pos + g.C > blk) return false;
  for (int f = 0; f < pp; ++f)
    if (tg.pos_off[f] % 4 != 0) return false;
  for (int r = 0; r < R; ++r)
    if (tg.row_org[r] % 4 != 0) return false;
  return true;
}

inline int tok_grid(int T) {
  const int want = (T + TOK_WARPS - 1) / TOK_WARPS;
  const int cap = cb_sm_count() * 4;
  return want < cap ? want : cap;
}

// ---------------------------------------------------------------------------------------
// RandZoom(keep_size) + ScaleIntensity + SpatialPad("end") of a raw-dtype batch on the device (cinema/mae/pretrain.py:
// 163-199): the zoom is MONAI's `Zoom` = torch `interpolate(scale_factor = z, align_corners = False)` of the UNPADDED frame
// (trilinear for SAX, bicubic with A = -0.75 for LAX) centre-padded with zeros / centre-cropped back to the frame's own
// size; ScaleIntensity then takes min / max over that zoomed frame, and the model-size padding is zero after scaling.
// Pass 1 resamples into fp32 and reduces the per-sample min / max (monotone uint keys, atomicMin / atomicMax); pass 2
// scales in place.  z == 1 reproduces the input exactly (all interpolation weights are 0 / 1).
// ---------------------------------------------------------------------------------------
struct ZoomGeom {
  int nd;    // 2 (bicubic) or 3 (trilinear)
  int S[3];  // model input size per axis (row-major, S[2] = 1 when nd == 2): the stride space of raw and out
};

__device__ __forceinline__ unsigned f32_order_key(float v) {  // monotone float -> uint
  const unsigned u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float f32_from_key(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

struct ZoomAxis {
  int E, O, half;
  bool crop;
};
__device__ __forceinline__ ZoomAxis zoom_axis(int extent, float z) {
  ZoomAxis a;
  a.E = extent;
  a.O = (int)floor((double)extent * (double)z);  // torch: output size = floor(input size * scale_factor), in double
  const int diff = a.E - a.O;
  a.crop = diff < 0;
  a.half = (diff < 0 ? -diff : diff) / 2;  // MONAI Zoom keep_size: pad [half, diff - half] or slice [half, half + E)
  return a;
}

__device__ __forceinline__ float cubic_w1(float x) { return ((-0.75f + 2.f) * x - (-0.75f + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cubic_w2(float x) { return ((-0.75f * x + 3.75f) * x - 6.f) * x + 3.f; }

template <typename T, bool CUBIC>
__global__ void __launch_bounds__(256)
zoom_resample_kernel(const T* __restrict__ raw, const int* __restrict__ extent, const float* __restrict__ zoom,
                     float* __restrict__ out, unsigned* __restrict__ keys, ZoomGeom g, int B) {
  pdl_prologue();
  const int b = blockIdx.y;
  const long long n = (long long)g.S[0] * g.S[1] * g.S[2];
  const T* src = raw + (long long)b * n;
  float* dst = out + (long long)b * n;
  const float z = __ldg(zoom + b);
  const float rs = (float)(1.0 / (double)z);  // area_pixel_compute_scale with a given scale factor
  ZoomAxis ax[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) ax[a] = zoom_axis(a < g.nd ? __ldg(extent + b * 3 + a) : 1, a < g.nd ? z : 1.f);
  float vmin = INFINITY, vmax = -INFINITY;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int c[3];
    long long r = i;
    c[2] = (int)(r % g.S[2]), r /= g.S[2];
    c[1] = (int)(r % g.S[1]);
    c[0] = (int)(r / g.S[1]);
    float v = 0.f;
    if (c[0] < ax[0].E && c[1] < ax[1].E && c[2] < ax[2].E) {  // inside the frame (the rest is SpatialPad: 0 after scaling)
      bool ok = true;
      float real[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const int o = ax[a].crop ? c[a] + ax[a].half : c[a] - ax[a].half;  // coordinate in the zoomed (O-sized) image
        ok = ok && o >= 0 && o < ax[a].O;
        real[a] = a < g.nd ? rs * ((float)o + 0.5f) - 0.5f : 0.f;
      }
      if (ok) {
        if constexpr (!CUBIC) {
          int i0[3], i1[3];
          float l1[3];
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            const float x = fmaxf(real[a], 0.f);
            i0[a] = min((int)x, ax[a].E - 1);
            i1[a] = i0[a] + (i0[a] < ax[a].E - 1 ? 1 : 0);
            l1[a] = fminf(fmaxf(x - (float)i0[a], 0.f), 1.f);
          }
          auto at = [&](int p, int q, int s) { return static_cast<float>(src[((long long)p * g.S[1] + q) * g.S[2] + s]); };
          const float w0 = 1.f - l1[2], w1 = l1[2];
          const float a00 = w0 * at(i0[0], i0[1], i0[2]) + w1 * at(i0[0], i0[1], i1[2]);
          const float a01 = w0 * at(i0[0], i1[1], i0[2]) + w1 * at(i0[0], i1[1], i1[2]);
          const float a10 = w0 * at(i1[0], i0[1], i0[2]) + w1 * at(i1[0], i0[1], i1[2]);
          const float a11 = w0 * at(i1[0], i1[1], i0[2]) + w1 * at(i1[0], i1[1], i1[2]);
          v = (1.f - l1[0]) * ((1.f - l1[1]) * a00 + l1[1] * a01) + l1[0] * ((1.f - l1[1]) * a10 + l1[1] * a11);
        } else {
          const float fy = floorf(real[0]), fx = floorf(real[1]);
          const int iy = (int)fy, ix = (int)fx;
          const float ty = real[0] - fy, tx = real[1] - fx;
          const float wy[4] = {cubic_w2(ty + 1.f), cubic_w1(ty), cubic_w1(1.f - ty), cubic_w2(2.f - ty)};
          const float wx[4] = {cubic_w2(tx + 1.f), cubic_w1(tx), cubic_w1(1.f - tx), cubic_w2(2.f - tx)};
          v = 0.f;
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const int yy = min(max(iy - 1 + p, 0), ax[0].E - 1);
            float row = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int xx = min(max(ix - 1 + q, 0), ax[1].E - 1);
              row += wx[q] * static_cast<float>(src[(long long)yy * g.S[1] + xx]);
            }
            v += wy[p] * row;
          }
        }
      }
      vmin = fminf(vmin, v), vmax = fmaxf(vmax, v);
    }
    dst[i] = v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
    vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
  }
  __shared__ float smin[8], smax[8];
  if ((threadIdx.x & 31) == 0) smin[threadIdx.x >> 5] = vmin, smax[threadIdx.x >> 5] = vmax;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) vmin = fminf(vmin, smin[w]), vmax = fmaxf(vmax, smax[w]);
    if (vmin <= vmax) {  // (a block that saw no voxel of the frame contributes nothing)
      atomicMin(keys + b, f32_order_key(vmin));
      atomicMax(keys + B + b, f32_order_key(vmax));
    }
  }
}

__global__ void __launch_bounds__(256)
zoom_scale_kernel(float* __restrict__ out, const int* __restrict__ extent, const unsigned* __restrict__ keys, ZoomGeom g, int B) {
  pdl_prologue();
  const int b = blockIdx.y;
  const long long n = (long long)g.S[0] * g.S[1] * g.S[2];
  float* dst = out + (long long)b * n;
  const float lo = f32_from_key(keys[b]), span = f32_from_key(keys[B + b]) - lo;
  const float inv = span > 0.f ? 1.0f / span : 0.f;
  int E[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) E[a] = a < g.nd ? __ldg(extent + b * 3 + a) : 1;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    long long r = i;
    const int c2 = (int)(r % g.S[2]);
    r /= g.S[2];
    const int c1 = (int)(r % g.S[1]), c0 = (int)(r / g.S[1]);
    if (c0 < E[0] && c1 < E[1] && c2 < E[2]) dst[i] = (dst[i] - lo) * inv;
  }
}

inline int blocks_for(long long work, int threads) {
  long long b = (work + threads - 1) / threads;
  const long long cap = (long long)cb_sm_count() * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

extern "C" int cb_scale_intensity(const void* raw, int raw_dtype, const float* lo, const float* hi, float* out, int B,
                                  long long n_per_sample, void* stream) {
  if (B <= 0 || n_per_sample <= 0) return 0;
  const int esz = raw_dtype == CB_DT_U8 ? 1 : (raw_dtype == CB_DT_F32 ? 4 : 2);
  CB_CHECK_ARG(raw_dtype == CB_DT_U8 || raw_dtype == CB_DT_I16 || raw_dtype == CB_DT_U16 || raw_dtype == CB_DT_F32,
               "scale_intensity: raw dtype %d not supported", raw_dtype);
  CB_CHECK_ARG(((uintptr_t)raw & 15) == 0 && ((uintptr_t)out & 15) == 0 && (n_per_sample * esz) % 16 == 0,
               "scale_intensity: samples must start on 16-byte boundaries");
  const int per_sample = blocks_for(n_per_sample * esz / 16 + 1, 256);
  dim3 grid(per_sample < 64 ? per_sample : 64, B);  // 64 x B blocks of 256 threads cover a 16-sample batch several times over
  cudaStream_t s = (cudaStream_t)stream;
  switch (raw_dtype) {
    case CB_DT_U8: cb_launch(scale_intensity_kernel<uint8_t>, grid, 256, 0, s, (const uint8_t*)raw, lo, hi, out, n_per_sample); break;
    case CB_DT_I16: cb_launch(scale_intensity_kernel<int16_t>, grid, 256, 0, s, (const int16_t*)raw, lo, hi, out, n_per_sample); break;
    case CB_DT_U16: cb_launch(scale_intensity_kernel<uint16_t>, grid, 256, 0, s, (const uint16_t*)raw, lo, hi, out, n_per_sample); break;
    default: cb_launch(scale_intensity_kernel<float>, grid, 256, 0, s, (const float*)raw, lo, hi, out, n_per_sample); break;
  }
  CB_LAUNCH_CHECK();
  return 0;
}

extern "C" int cb_zoom_intensity(const void* raw, int raw_dtype, const int* extent, const float* zoom, float* out,
                                 void* keys_ws, int B, int nd, const int* size, int cubic, void* stream) {
  if (B <= 0) return 0;
  CB_CHECK_ARG(raw_dtype == CB_DT_U8 || raw_dtype == CB_DT_I16 || raw_dtype == CB_DT_U16 || raw_dtype == CB_DT_F32,
               "zoom_intensity: raw dtype %d not supported", raw_dtype);
  CB_CHECK_ARG((nd == 2 && cubic) || (nd == 3 && !cubic), "zoom_intensity: bicubic is 2-D, trilinear is 3-D (nd=%d cubic=%d)", nd, cubic);
  CB_CHECK_ARG(extent != nullptr && zoom != nullptr && keys_ws != nullptr && size != nullptr, "zoom_intensity: null argument");
  ZoomGeom g;
  g.nd = nd;
  for (int a = 0; a < 3; ++a) g.S[a] = a < nd ? size[a] : 1;
  CB_CHECK_ARG(g.S[0] > 0 && g.S[1] > 0 && g.S[2] > 0, "zoom_intensity: bad size");
  const long long n = (long long)g.S[0] * g.S[1] * g.S[2];
  cudaStream_t s = (cudaStream_t)stream;
  unsigned* keys = reinterpret_cast<unsigned*>(keys_ws);
  CB_CUDA(cudaMemsetAsync(keys, 0xff, sizeof(unsigned) * B, s));      // running minima
  CB_CUDA(cudaMemsetAsync(keys + B, 0x00, sizeof(unsigned) * B, s));  // running maxima
  const int per_sample = blocks_for(n, 256);
  dim3 grid(per_sample < 96 ? per_sample : 96, B);
  switch (raw_dtype) {
#define CB_ZOOM_CASE(DT, T)                                                                                             \
  case DT:                                                                                                              \
    if (cubic) cb_launch(zoom_resample_kernel<T, true>, grid, 256, 0, s, (const T*)raw, extent, zoom, out, keys, g, B);  \
    else cb_launch(zoom_resample_kernel<T, false>, grid, 256, 0, s, (const T*)raw, extent, zoom, out, keys, g, B);       \
    break;
    CB_ZOOM_CASE(CB_DT_U8, uint8_t)
    CB_ZOOM_CASE(CB_DT_I16, int16_t)
    CB_ZOOM_CASE(CB_DT_U16, uint16_t)
    default:
      if (cubic) cb_launch(zoom_resample_kernel<float, true>, grid, 256, 0, s, (const float*)raw, extent, zoom, out, keys, g, B);
      else cb_launch(zoom_resample_kernel<float, false>, grid, 256, 0, s, (const float*)raw, extent, zoom, out, keys, g, B);
      break;
#undef CB_ZOOM_CASE
  }
  CB_LAUNCH_CHECK();
  cb_launch(zoom_scale_kernel, grid, 256, 0, s, out, extent, (const unsigned*)keys, g, B);
  CB_LAUNCH_CHECK();
  return 0;
}

extern "C" int cb_cast_f32_bf16(const float* src, void* dst, long long n, void* stream) {
  if (n <= 0) return 0;
  CB_CHECK_ARG(((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0, "cast: buffers must be 16-byte aligned");
  cb_launch(cast_f32_bf16_kernel, blocks_for(n / 8 + 1, 256), 256, 0, (cudaStream_t)stream, src, (bf16*)dst, n);
  CB_LAUNCH_CHECK();
  return 0;
}

extern "C" int cb_mask_to_index(const unsigned char* mask, int B, int n, int n_keep, int* keep_idx, int* drop_idx,
                                int* slot, void* stream) {
  CB_CHECK_ARG(B > 0 && n > 0 && n_keep >= 0 && n_keep <= n, "mask_to_index: bad sizes B=%d n=%d keep=%d", B, n, n_keep);
  cb_launch(mask_to_index_kernel, (B + 3) / 4, 128, 0, (cudaStream_t)stream, mask, B, n, n_keep, keep_idx, drop_idx, slot);
  CB_LAUNCH_CHECK();
  return 0;
}

extern "C" int cb_gather_rows(const void* src, long long src_bstride, const int* idx, int B, int k, void* out,
                              long long out_bstride, long long out_off, long long row_bytes, void* stream) {
  if (B <= 0 || k <= 0) return 0;
  CB_CHECK_ARG(row_bytes > 0 && row_bytes % 16 == 0, "gather_rows: row_bytes %lld must be a multiple of 16", row_bytes);
  const long long rows = (long long)B * k;
  cb_launch(move_rows_kernel<false>, (unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream, 
      (const uint4*)src, src_bstride, 0, idx, B, k, (uint4*)out, out_bstride, out_off, (int)(row_bytes / 16));
  CB_LAUNCH_CHECK();
  return 0;
}

extern "C" int cb_scatter_rows(const void* src, long long src_bstride, long long src_off, const int* idx, int B, int k,
                               void* dst, long long dst_bstride, long long row_bytes, void* stream) {
  if (B <= 0 || k <= 0) return 0;
  CB_CHECK_ARG(row_bytes > 0 && row_bytes % 16 == 0, "scatter_rows: row_bytes %lld must be a multiple of 16", row_bytes);
  const long long rows = (long long)B * k;
  cb_launch(move_rows_kernel<true>, (unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream, 
      (const uint4*)src, src_bstride, src_off, idx, B, k, (uint4*)dst, dst_bstride, 0, (int)(row_bytes / 16));
  CB_LAUNCH_CHECK();
  return 0;
}

extern "C" int cb_embed_rows_f32(const float* a, long long a_bstride, long long a_off, const float* row,
                                 const float* table, const int* idx, int B, int k, int D, float* out, void* out16,
                                 long long out_bstride, long long out_off, void* stream) {
  if (B <= 0 || k <= 0) return 0;
  CB_CHECK_ARG(D > 0 && D % 4 == 0, "embed_rows: D=%d must be a multiple of 4", D);
  CB_CHECK_ARG(out != nullptr || out16 != nullptr, "embed_rows: no output given");
  const long long rows = (long long)B * k;
  cb_launch(embed_rows_kernel, (unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream, 
      (const float4*)a, a_bstride, a_off, (const float4*)row, (const float4*)table, idx, B, k, D / 4, (float4*)out,
      (uint2*)out16, out_bstride, out_off);
  CB_LAUNCH_CHECK();
  return 0;
}

extern "C" int cb_patchify(const void* src, void* dst, int B, int C, int ndim, const int* spatial, const int* patch,
                           int elem_bytes, int inverse, void* stream) {
  NdGeom g;
  if (int rc = fill_geom(g, C, ndim, spatial, false, patch)) return rc;
  CB_CHECK_ARG(elem_bytes == 2 || elem_bytes == 4, "patchify: element size %d unsupported", elem_bytes);
  const long long per_batch = g.sb;
  const long long total = per_batch * B;
  if (total <= 0) return 0;
  const int blocks = blocks_for(total, 256);
  cudaStream_t s = (cudaStream_t)stream;
  if (elem_bytes == 4) {
    if (!inverse)
      cb_launch(patchify_kernel<float, false>, blocks, 256, 0, s, (const float*)src, (float*)dst, g, per_batch, total);
    else
      cb_launch(patchify_kernel<float, true>, blocks, 256, 0, s, (const float*)src, (float*)dst, g, per_batch, total);
  } else {
    if (!inverse)
      cb_launch(patchify_kernel<uint16_t, false>, blocks, 256, 0, s, (const uint16_t*)src, (uint16_t*)dst, g, per_batch, total);
    else
      cb_launch(patchify_kernel<uint16_t, true>, blocks, 256, 0, s, (const uint16_t*)src, (uint16_t*)dst, g, per_batch, total);
  }
  CB_LAUNCH_CHECK();
  return 0;
}

static int patches_common(NdGeom& g, long long sb, long long sc, const long long* sstride, int C, int ndim,
                          const int* grid, const int* patch) {
  if (int rc = fill_geom(g, C, ndim, grid, true, patch)) return rc;
  g.sb = sb, g.sc = sc;
  for (int a = 0; a < ndim; ++a) g.sstride[a] = sstride[a];
  return 0;
}

extern "C" int cb_gather_patches(const void* src, int src_dtype, long long sb, long long sc, const long long* sstride,
                                 int B, int C, int ndim, const int* grid, const int* patch, const int* idx, int k,
                                 int chan_last, void* out, void* stream) {
  NdGeom g;
  if (int rc = patches_common(g, sb, sc, sstride, C, ndim, grid, patch)) return rc;
  if (idx == nullptr) k = g.n_tok;
  const long long total = (long long)B * k * g.E;
  if (total <= 0) return 0;
  const int blocks = blocks_for(total, 256);
  cudaStream_t s = (cudaStream_t)stream;
  TokGeom tg;
  if (src_dtype == CB_DT_F32 && ((uintptr_t)out & 15) == 0 && token_major_geom(g, B, idx, chan_last, src, tg)) {
    constexpr int smem = TOK_WARPS * 2 * TOK_MAX_BLOCK * 4 + TOK_WARPS * 2 * 8;
    static bool attr_set = false;
    if (!attr_set) {
      CB_CUDA(cudaFuncSetAttribute(tok_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      attr_set = true;
    }
    cb_launch(tok_gather_kernel, tok_grid(B), TOK_WARPS * 32, smem, s, (const float*)src, (bf16*)out, tg);
    CB_LAUNCH_CHECK();
    return 0;
  }
  if (src_dtype == CB_DT_F32)
    cb_launch(patches_kernel<const float, bf16, false>, blocks, 256, 0, s, (const float*)src, (bf16*)out, g, idx, k, chan_last, total, 0);
  else
    cb_launch(patches_kernel<const bf16, bf16, false>, blocks, 256, 0, s, (const bf16*)src, (bf16*)out, g, idx, k, chan_last, total, 0);
  CB_LAUNCH_CHECK();
  return 0;
}

extern "C" int cb_scatter_patches(const void* rows, int rows_dtype, void* dst, int dst_dtype, long long sb, long long sc,
                                  const long long* sstride, int B, int C, int ndim, const int* grid, const int* patch,
                                  const int* idx, int k, int chan_last, int accumulate, void* stream) {
  NdGeom g;
  if (int rc = patches_common(g, sb, sc, sstride, C, ndim, grid, patch)) return rc;
  if (idx == nullptr) k = g.n_tok;
  const long long total = (long long)B * k * g.E;
  if (total <= 0) return 0;
  const int blocks = blocks_for(total, 256);
  cudaStream_t s = (cudaStream_t)stream;
  TokGeom tg;
  if (dst_dtype == CB_DT_F32 && ((uintptr_t)rows & 15) == 0 && token_major_geom(g, B, idx, chan_last, dst, tg)) {
    constexpr int smem = TOK_WARPS * 3 * TOK_MAX_BLOCK * 4 + TOK_WARPS * 2 * 8;
    static bool attr_set = false;
    if (!attr_set) {
      CB_CUDA(cudaFuncSetAttribute(tok_scatter_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      CB_CUDA(cudaFuncSetAttribute(tok_scatter_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      attr_set = true;
    }
    if (rows_dtype == CB_DT_BF16)
      cb_launch(tok_scatter_kernel<bf16>, tok_grid(B), TOK_WARPS * 32, smem, s, (const bf16*)rows, (float*)dst, tg, accumulate);
    else
      cb_launch(tok_scatter_kernel<float>, tok_grid(B), TOK_WARPS * 32, smem, s, (const float*)rows, (float*)dst, tg, accumulate);
    CB_LAUNCH_CHECK();
    return 0;
  }
  if (rows_dtype == CB_DT_BF16 && dst_dtype == CB_DT_BF16)
    cb_launch(patches_kernel<bf16, const bf16, true>, blocks, 256, 0, s, (bf16*)dst, (const bf16*)rows, g, idx, k, chan_last, total, accumulate);
  else if (rows_dtype == CB_DT_BF16 && dst_dtype == CB_DT_F32)
    cb_launch(patches_kernel<float, const bf16, true>, blocks, 256, 0, s, (float*)dst, (const bf16*)rows, g, idx, k, chan_last, total, accumulate);
  else if (rows_dtype == CB_DT_F32 && dst_dtype == CB_DT_F32)
    cb_launch(patches_kernel<float, const float, true>, blocks, 256, 0, s, (float*)dst, (const float*)rows, g, idx, k, chan_last, total, accumulate);
  else
    cb_launch(patches_kernel<bf16, const float, true>, blocks, 256, 0, s, (bf16*)dst, (const float*)rows, g, idx, k, chan_last, total, accumulate);
  CB_LAUNCH_CHECK();
  return 0;
}

extern "C" int cb_colsum_bf16(const void* X, long long ldx, int M, int N, float* out, void* stream) {
  if (M <= 0 || N <= 0) return 0;
  CB_CHECK_ARG(N % 2 == 0 && ldx % 2 == 0, "colsum: N and ldx must be even");
  const int col_blocks = (N + 63) / 64;
  int row_blocks = (cb_sm_count() * 8 + col_blocks - 1) / col_blocks;
  if (row_blocks > (M + 63) / 64) row_blocks = (M + 63) / 64;
  if (row_blocks < 1) row_blocks = 1;
  const int rows_per_block = (M + row_blocks - 1) / row_blocks;
  row_blocks = (M + rows_per_block - 1) / rows_per_block;
  cb_launch(colsum_bf16_kernel, dim3(col_blocks, row_blocks), 256, 0, (cudaStream_t)stream, (const bf16*)X, ldx, M, N, out,
                                                                                      rows_per_block);
  CB_LAUNCH_CHECK();
  return 0;
}

extern "C" int cb_colsum_seg_f32(const float* X, long long bstride_rows, long long off, int B, int k, int D, float* out,
                                 void* stream) {
  if (B <= 0 || k <= 0 || D <= 0) return 0;
  const long long total = (long long)B * k;
  const int col_blocks = (D + 127) / 128;
  long long row_blocks = ((long long)cb_sm_count() * 4 + col_blocks - 1) / col_blocks;
  if (row_blocks > (total + 15) / 16) row_blocks = (total + 15) / 16;
  if (row_blocks < 1) row_blocks = 1;
  const int rows_per_block = (int)((total + row_blocks - 1) / row_blocks);
  row_blocks = (total + rows_per_block - 1) / rows_per_block;
  cb_launch(colsum_seg_f32_kernel, dim3(col_blocks, (unsigned)row_blocks), 128, 0, (cudaStream_t)stream, 
      X, bstride_rows, off, B, k, D, out, rows_per_block);
  CB_LAUNCH_CHECK();
  return 0;
}

extern "C" int cb_scale_cast_bf16(const float* src, void* dst, long long n, const float* scale_dev, float scale,
                                  long long group, void* stream) {
  if (n <= 0) return 0;
  CB_CHECK_ARG(group == 0 || (scale_dev != nullptr && group % 4 == 0 && n % group == 0),
               "scale_cast: grouped mode needs device factors and a group size that is a multiple of 4 dividing n");
  CB_CHECK_ARG(((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 7) == 0, "scale_cast: buffers must be 16/8-byte aligned");
  cb_launch(scale_cast_kernel, blocks_for(n / 4 + 1, 256), 256, 0, (cudaStream_t)stream, src, (bf16*)dst, n, scale_dev, scale, group);
  CB_LAUNCH_CHECK();
  return 0;
}

extern "C" int cb_rope_apply(const void* x, void* y, int dtype, const float* cos_t, const float* sin_t, int B,
                             int n_tokens, int H, int d, int rotary_dim, int transpose, void* stream) {
  CB_CHECK_ARG(rotary_dim % 2 == 0 && rotary_dim >= 0, "rope: rotary dim %d must be even", rotary_dim);
  CB_CHECK_ARG(rotary_dim <= d, "Rotary dim %d is larger than the last dimension of x %d", rotary_dim, d);
  const long long rows = (long long)B * n_tokens * H;
  if (rows <= 0 || d <= 0) return 0;
  const float sgn = transpose ? -1.f : 1.f;
  const int blocks = blocks_for(rows * d, 256);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == CB_DT_F32)
    cb_launch(rope_kernel<float>, blocks, 256, 0, s, (const float*)x, (float*)y, cos_t, sin_t, rows, n_tokens, H, d, rotary_dim, sgn);
  else
    cb_launch(rope_kernel<bf16>, blocks, 256, 0, s, (const bf16*)x, (bf16*)y, cos_t, sin_t, rows, n_tokens, H, d, rotary_dim, sgn);
  CB_LAUNCH_CHECK();
  return 0;
}

// Row LayerNorm forward / backward, fp32 statistics, one warp per row with the row cached in
// registers (two-pass variance), 16-byte vector loads, warp-shuffle reductions.
// HBM-bound: forward moves 4 B in + 2 B out per element; backward 2 + 4 (+4) in, 4 (+2) out.
// Replaces ATen LayerNorm (fp32 under autocast) at cinema/vit.py:549,564,650,738,
// cinema/convvit.py:254,290 and ConvLayerNorm (cinema/conv.py:169-187) on channel-last rows.
#include "../../include/cinema_b200.h"
#include "common.cuh"

namespace {

// NV = number of float4 per lane; D <= NV * 128
template <int NV>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ x, long long ldx,
                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                     int M, int D, float eps, bf16* __restrict__ y16, long long ldy16,
                                                     float* __restrict__ y32, long long ldy32, float* __restrict__ mean_o,
                                                     float* __restrict__ rstd_o, int act) {
  pdl_prologue();  // PDL: release the next launch, then wait for the previous kernel's results
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int nvec = D >> 2;
  const float inv_d = 1.0f / (float)D;
  for (long long row = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < M;
       row += (long long)gridDim.x * warps_per_block) {
    const float4* xr = reinterpret_cast<const float4*>(x + row * ldx);
    float4 v[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + i * 32;
      v[i] = c < nvec ? __ldg(xr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
    const float mean = warp_sum(s) * inv_d;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) {
        const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
        ss += a * a + b * b + cc * cc + d * d;
      }
    }
    const float rstd = rsqrtf(warp_sum(ss) * inv_d + eps);
    if (lane == 0) {
      if (mean_o) mean_o[row] = mean;
      if (rstd_o) rstd_o[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c);
        const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c);
        float4 o;
        o.x = (v[i].x - mean) * rstd * g.x + b.x;
        o.y = (v[i].y - mean) * rstd * g.y + b.y;
        o.z = (v[i].z - mean) * rstd * g.z + b.z;
        o.w = (v[i].w - mean) * rstd * g.w + b.w;
        if (act) o.x = gelu_f(o.x), o.y = gelu_f(o.y), o.z = gelu_f(o.z), o.w = gelu_f(o.w);
        if (y32) reinterpret_cast<float4*>(y32 + row * ldy32)[c] = o;
        if (y16) reinterpret_cast<uint2*>(y16 + row * ldy16)[c] = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
      }
    }
  }
}

template <int NV, bool DY_BF16>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const void* __restrict__ dy_, long long lddy,
                                                     const float* __restrict__ x, long long ldx,
                                                     const float* __restrict__ mean_i, const float* __restrict__ rstd_i,
                                                     const float* __restrict__ gamma, const float* __restrict__ dres,
                                                     long long lddres, int M, int D, float* __restrict__ dx32,
                                                     long long lddx32, bf16* __restrict__ dx16, long long lddx16,
                                                     float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                     const float* __restrict__ beta_act) {
  pdl_prologue();  // PDL: release the next launch, then wait for the previous kernel's results
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int nvec = D >> 2;
  const float inv_d = 1.0f / (float)D;
  float4 dg[NV], db[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) dg[i] = db[i] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (long long row = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < M;
       row += (long long)gridDim.x * warps_per_block) {
    const float mean = mean_i[row], rstd = rstd_i[row];
    const float4* xr = reinterpret_cast<const float4*>(x + row * ldx);
    float4 xh[NV], dyg[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) {
        float4 d;
        if (DY_BF16) {
          const uint2 u = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(dy_) + row * lddy) + c);
          const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
          d = make_float4(a.x, a.y, b.x, b.y);
        } else {
          d = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy_) + row * lddy) + c);
        }
        const float4 xv = __ldg(xr + c);
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c);
        xh[i] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
        if (beta_act != nullptr) {  // the forward applied GELU to the LN output: dy <- dy * GELU'(xhat * gamma + beta)
          const float4 bb = __ldg(reinterpret_cast<const float4*>(beta_act) + c);
          d.x *= gelu_grad_f(fmaf(xh[i].x, g.x, bb.x)), d.y *= gelu_grad_f(fmaf(xh[i].y, g.y, bb.y));
          d.z *= gelu_grad_f(fmaf(xh[i].z, g.z, bb.z)), d.w *= gelu_grad_f(fmaf(xh[i].w, g.w, bb.w));
        }
        dyg[i] = make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w);
        s1 += dyg[i].x + dyg[i].y + dyg[i].z + dyg[i].w;
        s2 += dyg[i].x * xh[i].x + dyg[i].y * xh[i].y + dyg[i].z * xh[i].z + dyg[i].w * xh[i].w;
        dg[i].x += d.x * xh[i].x, dg[i].y += d.y * xh[i].y, dg[i].z += d.z * xh[i].z, dg[i].w += d.w * xh[i].w;
        db[i].x += d.x, db[i].y += d.y, db[i].z += d.z, db[i].w += d.w;
      } else {
        xh[i] = dyg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    const float m1 = warp_sum(s1) * inv_d;
    const float m2 = warp_sum(s2) * inv_d;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) {
        float4 o;
        o.x = rstd * (dyg[i].x - m1 - xh[i].x * m2);
        o.y = rstd * (dyg[i].y - m1 - xh[i].y * m2);
        o.z = rstd * (dyg[i].z - m1 - xh[i].z * m2);
        o.w = rstd * (dyg[i].w - m1 - xh[i].w * m2);
        if (dres) {
          const float4 r = __ldg(reinterpret_cast<const float4*>(dres + row * lddres) + c);
          o.x += r.x, o.y += r.y, o.z += r.z, o.w += r.w;
        }
        if (dx32) reinterpret_cast<float4*>(dx32 + row * lddx32)[c] = o;
        if (dx16) reinterpret_cast<uint2*>(dx16 + row * lddx16)[c] = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
      }
    }
  }

  if (dgamma == nullptr && dbeta == nullptr) return;
  // block reduction of the per-warp partial dgamma / dbeta, then one atomic per column per block
  extern __shared__ float sred[];  // [2][D]
  float* sg = sred;
  float* sb = sred + D;
  for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) sred[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      atomicAdd(sg + 4 * c + 0, dg[i].x), atomicAdd(sg + 4 * c + 1, dg[i].y);
      atomicAdd(sg + 4 * c + 2, dg[i].z), atomicAdd(sg + 4 * c + 3, dg[i].w);
      atomicAdd(sb + 4 * c + 0, db[i].x), atomicAdd(sb + 4 * c + 1, db[i].y);
      atomicAdd(sb + 4 * c + 2, db[i].z), atomicAdd(sb + 4 * c + 3, db[i].w);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    if (dgamma) atomicAdd(dgamma + i, sg[i]);
    if (dbeta) atomicAdd(dbeta + i, sb[i]);
  }
}


// ------------------------------------------------------------------------------------------------------------
// Backward, TMA-staged.  The register-cached kernel above is latency-bound (two dependent load phases per row and
// ~130 live registers leave ~16 warps x 4.5 KB in flight per SM); here every warp runs a private ring of row
// slots in shared memory that lane 0 fills with 1-D bulk copies (cp.async.bulk, completion on a per-slot
// mbarrier), so ~150 KB per SM are in flight without holding a register, and the row is re-read from shared memory
// for the second pass.  Outputs go straight to global memory as 16-byte vectors.
// Optionally emits dxsum[c] += sum_rows bf16(dx[row, c]): the bias gradient of the Linear layer that consumes dx
// as its output gradient (replaces a separate column-sum pass over dx16).
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

constexpr int LNB_WARPS = 8;

template <int NV, bool DY_BF16>
__global__ void __launch_bounds__(LNB_WARPS * 32, 1)
ln_bwd_tma_kernel(const void* __restrict__ dy_, long long lddy, const float* __restrict__ x, long long ldx,
                  const float* __restrict__ mean_i, const float* __restrict__ rstd_i, const float* __restrict__ gamma,
                  const float* __restrict__ dres, long long lddres, int M, int D, float* __restrict__ dx32,
                  long long lddx32, bf16* __restrict__ dx16, long long lddx16, float* __restrict__ dgamma,
                  float* __restrict__ dbeta, float* __restrict__ dxsum, const float* __restrict__ beta_act, int stages,
                  int slot_bytes) {
  pdl_prologue();  // PDL: release the next launch, then wait for the previous kernel's results
  extern __shared__ __align__(128) uint8_t lsm[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nvec = D >> 2;
  const float inv_d = 1.0f / (float)D;
  const uint32_t x_bytes = (uint32_t)D * 4u;
  const uint32_t dy_bytes = (uint32_t)D * (DY_BF16 ? 2u : 4u);
  const bool has_res = dres != nullptr;
  uint8_t* ring = lsm + (size_t)warp * stages * slot_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(lsm + (size_t)LNB_WARPS * stages * slot_bytes) + warp * stages;
  if (lane == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(&bars[s], 1);
    mbar_fence_init();
  }
  __syncwarp();
  const long long first = (long long)blockIdx.x * LNB_WARPS + warp;
  const long long stride = (long long)gridDim.x * LNB_WARPS;
  auto issue = [&](long long row, int s) {  // lane 0 only
    uint8_t* slot = ring + (size_t)s * slot_bytes;
    mbar_expect_tx(&bars[s], x_bytes + dy_bytes + (has_res ? x_bytes : 0u));
    bulk_load_1d(slot, x + row * ldx, x_bytes, &bars[s]);
    if (has_res) bulk_load_1d(slot + x_bytes, dres + row * lddres, x_bytes, &bars[s]);
    bulk_load_1d(slot + 2 * x_bytes,
                 DY_BF16 ? (const void*)(reinterpret_cast<const bf16*>(dy_) + row * lddy)
                         : (const void*)(reinterpret_cast<const float*>(dy_) + row * lddy),
                 dy_bytes, &bars[s]);
  };
  if (lane == 0) {
    for (int s = 0; s < stages; ++s) {
      const long long row = first + (long long)s * stride;
      if (row < M) issue(row, s);
    }
  }
  float4 dg[NV], db[NV], dxs[NV], gm[NV], bt[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    dg[i] = db[i] = dxs[i] = bt[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int c = lane + i * 32;
    gm[i] = c < nvec ? __ldg(reinterpret_cast<const float4*>(gamma) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (beta_act != nullptr && c < nvec) bt[i] = __ldg(reinterpret_cast<const float4*>(beta_act) + c);
  }

  int s = 0;
  uint32_t phase = 0;
  for (long long row = first; row < M; row += stride) {
    const float mean = __ldg(mean_i + row), rstd = __ldg(rstd_i + row);
    const float nmr = -mean * rstd;
    mbar_wait(&bars[s], phase);
    const uint8_t* slot = ring + (size_t)s * slot_bytes;
    const float4* xs = reinterpret_cast<const float4*>(slot);
    const float4* rs = reinterpret_cast<const float4*>(slot + x_bytes);
    const uint8_t* dys = slot + 2 * x_bytes;
    float4 xh[NV], dyg[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + i * 32;
      xh[i] = dyg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < nvec) {
        float4 d;
        if (DY_BF16) {
          const uint2 u = reinterpret_cast<const uint2*>(dys)[c];
          const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
          d = make_float4(a.x, a.y, b.x, b.y);
        } else {
          d = reinterpret_cast<const float4*>(dys)[c];
        }
        const float4 xv = xs[c];
        const float4 g = gm[i];
        xh[i] = make_float4(fmaf(xv.x, rstd, nmr), fmaf(xv.y, rstd, nmr), fmaf(xv.z, rstd, nmr), fmaf(xv.w, rstd, nmr));
        if (beta_act != nullptr) {  // the forward applied GELU to the LN output: dy <- dy * GELU'(xhat * gamma + beta)
          d.x *= gelu_grad_f(fmaf(xh[i].x, g.x, bt[i].x)), d.y *= gelu_grad_f(fmaf(xh[i].y, g.y, bt[i].y));
          d.z *= gelu_grad_f(fmaf(xh[i].z, g.z, bt[i].z)), d.w *= gelu_grad_f(fmaf(xh[i].w, g.w, bt[i].w));
        }
        dyg[i] = make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w);
        s1 += (dyg[i].x + dyg[i].y) + (dyg[i].z + dyg[i].w);
        s2 = fmaf(dyg[i].x, xh[i].x, fmaf(dyg[i].y, xh[i].y, fmaf(dyg[i].z, xh[i].z, fmaf(dyg[i].w, xh[i].w, s2))));
        dg[i].x = fmaf(d.x, xh[i].x, dg[i].x), dg[i].y = fmaf(d.y, xh[i].y, dg[i].y);
        dg[i].z = fmaf(d.z, xh[i].z, dg[i].z), dg[i].w = fmaf(d.w, xh[i].w, dg[i].w);
        db[i].x += d.x, db[i].y += d.y, db[i].z += d.z, db[i].w += d.w;
      }
    }
    const float m1r = warp_sum(s1) * inv_d * rstd;
    const float m2r = -warp_sum(s2) * inv_d * rstd;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) {
        float4 o;  // rstd * (dyg - m1 - xh * m2)
        o.x = fmaf(xh[i].x, m2r, fmaf(dyg[i].x, rstd, -m1r));
        o.y = fmaf(xh[i].y, m2r, fmaf(dyg[i].y, rstd, -m1r));
        o.z = fmaf(xh[i].z, m2r, fmaf(dyg[i].z, rstd, -m1r));
        o.w = fmaf(xh[i].w, m2r, fmaf(dyg[i].w, rstd, -m1r));
        if (has_res) {
          const float4 r = rs[c];
          o.x += r.x, o.y += r.y, o.z += r.z, o.w += r.w;
        }
        if (dx32) reinterpret_cast<float4*>(dx32 + row * lddx32)[c] = o;
        const uint2 pk = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
        if (dx16) reinterpret_cast<uint2*>(dx16 + row * lddx16)[c] = pk;
        if (dxsum != nullptr) {
          const float2 a = unpack_bf16(pk.x), b = unpack_bf16(pk.y);
          dxs[i].x += a.x, dxs[i].y += a.y, dxs[i].z += b.x, dxs[i].w += b.y;
        }
      }
    }
    __syncwarp();  // every lane is done with the slot
    if (lane == 0) {
      const long long nrow = row + (long long)stages * stride;
      if (nrow < M) {
        fence_proxy_async_smem();
        issue(nrow, s);
      }
    }
    if (++s == stages) s = 0, phase ^= 1;
  }

  if (dgamma == nullptr && dbeta == nullptr && dxsum == nullptr) return;
  // block reduction of the per-warp partial sums in the (now idle) ring, then one atomic per column per block
  __syncthreads();
  float* sred = reinterpret_cast<float*>(lsm);  // [3][D]
  for (int i = threadIdx.x; i < 3 * D; i += blockDim.x) sred[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      float* sg = sred + 4 * c;
      atomicAdd(sg + 0, dg[i].x), atomicAdd(sg + 1, dg[i].y), atomicAdd(sg + 2, dg[i].z), atomicAdd(sg + 3, dg[i].w);
      float* sb = sred + D + 4 * c;
      atomicAdd(sb + 0, db[i].x), atomicAdd(sb + 1, db[i].y), atomicAdd(sb + 2, db[i].z), atomicAdd(sb + 3, db[i].w);
      if (dxsum != nullptr) {
        float* sx = sred + 2 * D + 4 * c;
        atomicAdd(sx + 0, dxs[i].x), atomicAdd(sx + 1, dxs[i].y), atomicAdd(sx + 2, dxs[i].z), atomicAdd(sx + 3, dxs[i].w);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    if (dgamma) atomicAdd(dgamma + i, sred[i]);
    if (dbeta) atomicAdd(dbeta + i, sred[D + i]);
    if (dxsum) atomicAdd(dxsum + i, sred[2 * D + i]);
  }
}


// ------------------------------------------------------------------------------------------------------------
// Short rows (the ConvMAE stem: D = 64 / 128 channels, ~150 k rows): LPR = D / 4 lanes own one row (one float4
// each), a warp works on 32 / LPR rows at once and keeps UNR independent row groups in flight, reductions are
// LPR-wide shuffles.  The per-column sums (dgamma, dbeta, dxsum) of a lane's four columns live in registers.
// ------------------------------------------------------------------------------------------------------------
template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int LPR, int UNR>
__global__ void __launch_bounds__(256) ln_fwd_small_kernel(const float* __restrict__ x, long long ldx,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           int M, float eps, bf16* __restrict__ y16, long long ldy16,
                                                           float* __restrict__ y32, long long ldy32,
                                                           float* __restrict__ mean_o, float* __restrict__ rstd_o, int act) {
  pdl_prologue();  // PDL: release the next launch, then wait for the previous kernel's results
  constexpr int RPW = 32 / LPR;  // rows per warp pass
  const int lane = threadIdx.x & 31;
  const int c = lane % LPR;      // float4 column of this lane
  const int rsub = lane / LPR;
  const float inv_d = 1.0f / (float)(4 * LPR);
  const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c);
  const float4 bb = __ldg(reinterpret_cast<const float4*>(beta) + c);
  const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long base = wid * (RPW * UNR); base < M; base += nw * (RPW * UNR)) {
    float4 v[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long row = base + u * RPW + rsub;
      v[u] = row < M ? __ldg(reinterpret_cast<const float4*>(x + row * ldx) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long row = base + u * RPW + rsub;
      const float mean = group_sum<LPR>((v[u].x + v[u].y) + (v[u].z + v[u].w)) * inv_d;
      const float a0 = v[u].x - mean, a1 = v[u].y - mean, a2 = v[u].z - mean, a3 = v[u].w - mean;
      const float rstd = rsqrtf(group_sum<LPR>(a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3) * inv_d + eps);
      if (row < M) {
        if (c == 0) {
          if (mean_o) mean_o[row] = mean;
          if (rstd_o) rstd_o[row] = rstd;
        }
        float4 o = make_float4(fmaf(a0 * rstd, g.x, bb.x), fmaf(a1 * rstd, g.y, bb.y), fmaf(a2 * rstd, g.z, bb.z),
                               fmaf(a3 * rstd, g.w, bb.w));
        if (act) o.x = gelu_f(o.x), o.y = gelu_f(o.y), o.z = gelu_f(o.z), o.w = gelu_f(o.w);
        if (y32) reinterpret_cast<float4*>(y32 + row * ldy32)[c] = o;
        if (y16) reinterpret_cast<uint2*>(y16 + row * ldy16)[c] = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
      }
    }
  }
}

template <int LPR, int UNR, bool DY_BF16>
__global__ void __launch_bounds__(256, 2)  // <= 128 registers: two blocks (16 warps) per SM; at 151 registers only one was resident
ln_bwd_small_kernel(const void* __restrict__ dy_, long long lddy, const float* __restrict__ x, long long ldx,
                    const float* __restrict__ mean_i, const float* __restrict__ rstd_i, const float* __restrict__ gamma,
                    const float* __restrict__ dres, long long lddres, int M, float* __restrict__ dx32, long long lddx32,
                    bf16* __restrict__ dx16, long long lddx16, float* __restrict__ dgamma, float* __restrict__ dbeta,
                    float* __restrict__ dxsum, const float* __restrict__ beta_act) {
  pdl_prologue();  // PDL: release the next launch, then wait for the previous kernel's results
  constexpr int RPW = 32 / LPR;
  constexpr int D = 4 * LPR;
  const int lane = threadIdx.x & 31;
  const int c = lane % LPR;
  const int rsub = lane / LPR;
  const float inv_d = 1.0f / (float)D;
  const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c);
  float4 bt = make_float4(0.f, 0.f, 0.f, 0.f);
  if (beta_act != nullptr) bt = __ldg(reinterpret_cast<const float4*>(beta_act) + c);
  float4 dg = make_float4(0.f, 0.f, 0.f, 0.f), db = dg, dxs = dg;
  const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long base = wid * (RPW * UNR); base < M; base += nw * (RPW * UNR)) {
    float4 xv[UNR], d[UNR], rs[UNR];
    float mean[UNR], rstd[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long row = base + u * RPW + rsub;
      xv[u] = d[u] = rs[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      mean[u] = 0.f, rstd[u] = 0.f;
      if (row < M) {
        xv[u] = __ldg(reinterpret_cast<const float4*>(x + row * ldx) + c);
        if (DY_BF16) {
          const uint2 w = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(dy_) + row * lddy) + c);
          const float2 a = unpack_bf16(w.x), b = unpack_bf16(w.y);
          d[u] = make_float4(a.x, a.y, b.x, b.y);
        } else {
          d[u] = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy_) + row * lddy) + c);
        }
        if (dres) rs[u] = __ldg(reinterpret_cast<const float4*>(dres + row * lddres) + c);
        mean[u] = __ldg(mean_i + row), rstd[u] = __ldg(rstd_i + row);
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long row = base + u * RPW + rsub;
      const float nmr = -mean[u] * rstd[u];
      const float4 xh = make_float4(fmaf(xv[u].x, rstd[u], nmr), fmaf(xv[u].y, rstd[u], nmr), fmaf(xv[u].z, rstd[u], nmr),
                                    fmaf(xv[u].w, rstd[u], nmr));
      float4 dd = d[u];
      if (beta_act != nullptr) {  // the forward applied GELU to the LN output
        dd.x *= gelu_grad_f(fmaf(xh.x, g.x, bt.x)), dd.y *= gelu_grad_f(fmaf(xh.y, g.y, bt.y));
        dd.z *= gelu_grad_f(fmaf(xh.z, g.z, bt.z)), dd.w *= gelu_grad_f(fmaf(xh.w, g.w, bt.w));
      }
      const float4 dyg = make_float4(dd.x * g.x, dd.y * g.y, dd.z * g.z, dd.w * g.w);
      const float m1r = group_sum<LPR>((dyg.x + dyg.y) + (dyg.z + dyg.w)) * inv_d * rstd[u];
      const float m2r = -group_sum<LPR>(fmaf(dyg.x, xh.x, fmaf(dyg.y, xh.y, fmaf(dyg.z, xh.z, dyg.w * xh.w)))) * inv_d * rstd[u];
      if (row < M) {
        dg.x = fmaf(dd.x, xh.x, dg.x), dg.y = fmaf(dd.y, xh.y, dg.y), dg.z = fmaf(dd.z, xh.z, dg.z), dg.w = fmaf(dd.w, xh.w, dg.w);
        db.x += dd.x, db.y += dd.y, db.z += dd.z, db.w += dd.w;
        float4 o;
        o.x = fmaf(xh.x, m2r, fmaf(dyg.x, rstd[u], -m1r)) + rs[u].x;
        o.y = fmaf(xh.y, m2r, fmaf(dyg.y, rstd[u], -m1r)) + rs[u].y;
        o.z = fmaf(xh.z, m2r, fmaf(dyg.z, rstd[u], -m1r)) + rs[u].z;
        o.w = fmaf(xh.w, m2r, fmaf(dyg.w, rstd[u], -m1r)) + rs[u].w;
        if (dx32) reinterpret_cast<float4*>(dx32 + row * lddx32)[c] = o;
        const uint2 pk = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
        if (dx16) reinterpret_cast<uint2*>(dx16 + row * lddx16)[c] = pk;
        const float2 a = unpack_bf16(pk.x), b = unpack_bf16(pk.y);
        dxs.x += a.x, dxs.y += a.y, dxs.z += b.x, dxs.w += b.y;
      }
    }
  }
  if (dgamma == nullptr && dbeta == nullptr && dxsum == nullptr) return;
  __shared__ float sred[3 * D];
  for (int i = threadIdx.x; i < 3 * D; i += blockDim.x) sred[i] = 0.f;
  __syncthreads();
  atomicAdd(&sred[4 * c + 0], dg.x), atomicAdd(&sred[4 * c + 1], dg.y), atomicAdd(&sred[4 * c + 2], dg.z), atomicAdd(&sred[4 * c + 3], dg.w);
  atomicAdd(&sred[D + 4 * c + 0], db.x), atomicAdd(&sred[D + 4 * c + 1], db.y), atomicAdd(&sred[D + 4 * c + 2], db.z), atomicAdd(&sred[D + 4 * c + 3], db.w);
  if (dxsum != nullptr) {
    atomicAdd(&sred[2 * D + 4 * c + 0], dxs.x), atomicAdd(&sred[2 * D + 4 * c + 1], dxs.y);
    atomicAdd(&sred[2 * D + 4 * c + 2], dxs.z), atomicAdd(&sred[2 * D + 4 * c + 3], dxs.w);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    if (dgamma) atomicAdd(dgamma + i, sred[i]);
    if (dbeta) atomicAdd(dbeta + i, sred[D + i]);
    if (dxsum) atomicAdd(dxsum + i, sred[2 * D + i]);
  }
}

inline int pick_nv(int D) {
  const int need = (D / 4 + 31) / 32;
  const int opts[] = {1, 2, 4, 6, 8, 12, 16};
  for (int o : opts)
    if (need <= o) return o;
  return -1;
}

}  // namespace

#define LN_DISPATCH(NVVAL, CALL) \
  case NVVAL: {                  \
    constexpr int NV = NVVAL;    \
    CALL;                        \
  } break;

extern "C" int cb_layernorm_fwd(const float* x, long long ldx, const float* gamma, const float* beta, int M, int D,
                                float eps, void* y16, long long ldy16, float* y32, long long ldy32, float* mean,
                                float* rstd, int act, void* stream) {
  if (M <= 0) return 0;
  CB_CHECK_ARG(D > 0 && D % 4 == 0 && D <= 2048, "layernorm: D=%d must be a multiple of 4 and <= 2048", D);
  CB_CHECK_ARG(ldx % 4 == 0 && (y16 == nullptr || ldy16 % 4 == 0) && (y32 == nullptr || ldy32 % 4 == 0),
               "layernorm: row pitches must be multiples of 4 elements");
  const int nv = pick_nv(D);
  const int blocks = (int)min((long long)(M + 7) / 8, (long long)cb_sm_count() * 8);
  cudaStream_t s = (cudaStream_t)stream;
  if (D == 16 || D == 32 || D == 64 || D == 128) {  // short rows: several rows per warp pass
    constexpr int UNR = 4;
    const long long rows_per_block = 8ll * (128 / D) * UNR;
    const int sb = (int)min((M + rows_per_block - 1) / rows_per_block, (long long)cb_sm_count() * 8);
#define LN_FWD_SMALL(L) \
  cb_launch(ln_fwd_small_kernel<L, UNR>, sb, 256, 0, s, x, ldx, gamma, beta, M, eps, (bf16*)y16, ldy16, y32, ldy32, mean, rstd, act)
    if (D == 16) LN_FWD_SMALL(4);
    else if (D == 32) LN_FWD_SMALL(8);
    else if (D == 64) LN_FWD_SMALL(16);
    else LN_FWD_SMALL(32);
#undef LN_FWD_SMALL
    CB_LAUNCH_CHECK();
    return 0;
  }
  switch (nv) {
    LN_DISPATCH(1, (cb_launch(ln_fwd_kernel<NV>, blocks, 256, 0, s, x, ldx, gamma, beta, M, D, eps, (bf16*)y16, ldy16, y32, ldy32, mean, rstd, act)))
    LN_DISPATCH(2, (cb_launch(ln_fwd_kernel<NV>, blocks, 256, 0, s, x, ldx, gamma, beta, M, D, eps, (bf16*)y16, ldy16, y32, ldy32, mean, rstd, act)))
    LN_DISPATCH(4, (cb_launch(ln_fwd_kernel<NV>, blocks, 256, 0, s, x, ldx, gamma, beta, M, D, eps, (bf16*)y16, ldy16, y32, ldy32, mean, rstd, act)))
    LN_DISPATCH(6, (cb_launch(ln_fwd_kernel<NV>, blocks, 256, 0, s, x, ldx, gamma, beta, M, D, eps, (bf16*)y16, ldy16, y32, ldy32, mean, rstd, act)))
    LN_DISPATCH(8, (cb_launch(ln_fwd_kernel<NV>, blocks, 256, 0, s, x, ldx, gamma, beta, M, D, eps, (bf16*)y16, ldy16, y32, ldy32, mean, rstd, act)))
    LN_DISPATCH(12, (cb_launch(ln_fwd_kernel<NV>, blocks, 256, 0, s, x, ldx, gamma, beta, M, D, eps, (bf16*)y16, ldy16, y32, ldy32, mean, rstd, act)))
    LN_DISPATCH(16, (cb_launch(ln_fwd_kernel<NV>, blocks, 256, 0, s, x, ldx, gamma, beta, M, D, eps, (bf16*)y16, ldy16, y32, ldy32, mean, rstd, act)))
    default: CB_CHECK_ARG(false, "layernorm: unsupported D=%d", D);
  }
  CB_LAUNCH_CHECK();
  return 0;
}

extern "C" int cb_layernorm_bwd(const void* dy, long long lddy, int dy_dtype, const float* x, long long ldx,
                                const float* mean, const float* rstd, const float* gamma, const float* dres,
                                long long lddres, int M, int D, float* dx32, long long lddx32, void* dx16,
                                long long lddx16, float* dgamma, float* dbeta, float* dxsum, const float* beta_act,
                                void* stream) {
  if (M <= 0) return 0;
  CB_CHECK_ARG(D > 0 && D % 4 == 0 && D <= 2048, "layernorm_bwd: D=%d must be a multiple of 4 and <= 2048", D);
  CB_CHECK_ARG(lddy % 4 == 0 && ldx % 4 == 0, "layernorm_bwd: row pitches must be multiples of 4 elements");
  const int nv = pick_nv(D);
  cudaStream_t s = (cudaStream_t)stream;
  const bool bf = dy_dtype == CB_DT_BF16;
  // TMA-staged path: 16-byte aligned rows for the bulk copies
  const size_t dy_row = (size_t)D * (bf ? 2 : 4);
  const bool aligned = dy_row % 16 == 0 && ((size_t)lddy * (bf ? 2 : 4)) % 16 == 0 && ((uintptr_t)dy & 15) == 0 &&
                       ((uintptr_t)x & 15) == 0 && (dres == nullptr || (((uintptr_t)dres & 15) == 0 && lddres % 4 == 0));
  if (D == 16 || D == 32 || D == 64 || D == 128) {  // short rows (stem): several rows per warp pass, fused column sums
    constexpr int UNR = 4;
    const long long rows_per_block = 8ll * (128 / D) * UNR;
    const int sb = (int)min((M + rows_per_block - 1) / rows_per_block, (long long)cb_sm_count() * 4);
#define LN_BWD_SMALL(L)                                                                                             \
  do {                                                                                                              \
    if (bf)                                                                                                         \
      cb_launch(ln_bwd_small_kernel<L, UNR, true>, sb, 256, 0, s, dy, lddy, x, ldx, mean, rstd, gamma, dres, lddres, M, dx32, \
                                                           lddx32, (bf16*)dx16, lddx16, dgamma, dbeta, dxsum, beta_act); \
    else                                                                                                            \
      cb_launch(ln_bwd_small_kernel<L, UNR, false>, sb, 256, 0, s, dy, lddy, x, ldx, mean, rstd, gamma, dres, lddres, M, dx32, \
                                                            lddx32, (bf16*)dx16, lddx16, dgamma, dbeta, dxsum, beta_act); \
  } while (0)
    if (D == 16) LN_BWD_SMALL(4);
    else if (D == 32) LN_BWD_SMALL(8);
    else if (D == 64) LN_BWD_SMALL(16);
    else LN_BWD_SMALL(32);
#undef LN_BWD_SMALL
    CB_LAUNCH_CHECK();
    return 0;
  }
  // mid-length rows go through the TMA-staged kernel; very long rows (per-lane accumulators would spill) and unaligned
  // ones through the register-cached kernel
  if (aligned && nv > 0 && D >= 256 && D <= 1024) {
    const int slot_bytes = (int)(((size_t)D * 8 + dy_row + 127) / 128 * 128);
    int stages = (int)((size_t)200 * 1024 / ((size_t)LNB_WARPS * slot_bytes));
    if (stages > 8) stages = 8;
    const size_t red_bytes = 3 * (size_t)D * sizeof(float);
    size_t smem = (size_t)LNB_WARPS * stages * slot_bytes + LNB_WARPS * 8 * sizeof(uint64_t);
    if (stages >= 1 && (size_t)LNB_WARPS * stages * slot_bytes >= red_bytes) {
      const long long row_groups = ((long long)M + LNB_WARPS - 1) / LNB_WARPS;
      const int blocks = (int)min(row_groups, (long long)cb_sm_count());
#define LN_TMA_CALL(BF)                                                                                               \
  do {                                                                                                                \
    auto kern = ln_bwd_tma_kernel<NV, BF>;                                                                            \
    static size_t configured = 0;                                                                                     \
    if (smem > configured) {                                                                                          \
      CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                    \
      configured = smem;                                                                                              \
    }                                                                                                                 \
    cb_launch(kern, blocks, LNB_WARPS * 32, smem, s, dy, lddy, x, ldx, mean, rstd, gamma, dres, lddres, M, D, dx32, lddx32,  \
                                              (bf16*)dx16, lddx16, dgamma, dbeta, dxsum, beta_act, stages, slot_bytes); \
  } while (0)
      switch (nv) {
        LN_DISPATCH(1, if (bf) LN_TMA_CALL(true); else LN_TMA_CALL(false))
        LN_DISPATCH(2, if (bf) LN_TMA_CALL(true); else LN_TMA_CALL(false))
        LN_DISPATCH(4, if (bf) LN_TMA_CALL(true); else LN_TMA_CALL(false))
        LN_DISPATCH(6, if (bf) LN_TMA_CALL(true); else LN_TMA_CALL(false))
        LN_DISPATCH(8, if (bf) LN_TMA_CALL(true); else LN_TMA_CALL(false))
        LN_DISPATCH(12, if (bf) LN_TMA_CALL(true); else LN_TMA_CALL(false))
        LN_DISPATCH(16, if (bf) LN_TMA_CALL(true); else LN_TMA_CALL(false))
        default: CB_CHECK_ARG(false, "layernorm_bwd: unsupported D=%d", D);
      }
#undef LN_TMA_CALL
      CB_LAUNCH_CHECK();
      return 0;
    }
  }
  CB_CHECK_ARG(dxsum == nullptr, "layernorm_bwd: dxsum needs D in {16, 32, 64, 128} or 256 <= D <= 1024 with 16-byte aligned rows (TMA-staged path)");
  // few, fat blocks: every block ends with 2*D global atomics for dgamma / dbeta
  const int blocks = (int)min((long long)(M + 7) / 8, (long long)cb_sm_count() * 2);
  const size_t smem = 2 * (size_t)D * sizeof(float);
#define LN_BWD_CALL(BF)                                                                                              \
  cb_launch(ln_bwd_kernel<NV, BF>, blocks, 256, smem, s, dy, lddy, x, ldx, mean, rstd, gamma, dres, lddres, M, D, dx32, lddx32, \
                                                  (bf16*)dx16, lddx16, dgamma, dbeta, beta_act)
  switch (nv) {
    LN_DISPATCH(1, if (bf) LN_BWD_CALL(true); else LN_BWD_CALL(false))
    LN_DISPATCH(2, if (bf) LN_BWD_CALL(true); else LN_BWD_CALL(false))
    LN_DISPATCH(4, if (bf) LN_BWD_CALL(true); else LN_BWD_CALL(false))
    LN_DISPATCH(6, if (bf) LN_BWD_CALL(true); else LN_BWD_CALL(false))
    LN_DISPATCH(8, if (bf) LN_BWD_CALL(true); else LN_BWD_CALL(false))
    LN_DISPATCH(12, if (bf) LN_BWD_CALL(true); else LN_BWD_CALL(false))
    LN_DISPATCH(16, if (bf) LN_BWD_CALL(true); else LN_BWD_CALL(false))
    default: CB_CHECK_ARG(false, "layernorm_bwd: unsupported D=%d", D);
  }
#undef LN_BWD_CALL
  CB_LAUNCH_CHECK();
  return 0;
}

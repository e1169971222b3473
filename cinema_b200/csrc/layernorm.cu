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
  switch (nv) {
    LN_DISPATCH(1, (ln_fwd_kernel<NV><<<blocks, 256, 0, s>>>(x, ldx, gamma, beta, M, D, eps, (bf16*)y16, ldy16, y32, ldy32, mean, rstd, act)))
    LN_DISPATCH(2, (ln_fwd_kernel<NV><<<blocks, 256, 0, s>>>(x, ldx, gamma, beta, M, D, eps, (bf16*)y16, ldy16, y32, ldy32, mean, rstd, act)))
    LN_DISPATCH(4, (ln_fwd_kernel<NV><<<blocks, 256, 0, s>>>(x, ldx, gamma, beta, M, D, eps, (bf16*)y16, ldy16, y32, ldy32, mean, rstd, act)))
    LN_DISPATCH(6, (ln_fwd_kernel<NV><<<blocks, 256, 0, s>>>(x, ldx, gamma, beta, M, D, eps, (bf16*)y16, ldy16, y32, ldy32, mean, rstd, act)))
    LN_DISPATCH(8, (ln_fwd_kernel<NV><<<blocks, 256, 0, s>>>(x, ldx, gamma, beta, M, D, eps, (bf16*)y16, ldy16, y32, ldy32, mean, rstd, act)))
    LN_DISPATCH(12, (ln_fwd_kernel<NV><<<blocks, 256, 0, s>>>(x, ldx, gamma, beta, M, D, eps, (bf16*)y16, ldy16, y32, ldy32, mean, rstd, act)))
    LN_DISPATCH(16, (ln_fwd_kernel<NV><<<blocks, 256, 0, s>>>(x, ldx, gamma, beta, M, D, eps, (bf16*)y16, ldy16, y32, ldy32, mean, rstd, act)))
    default: CB_CHECK_ARG(false, "layernorm: unsupported D=%d", D);
  }
  CB_LAUNCH_CHECK();
  return 0;
}

extern "C" int cb_layernorm_bwd(const void* dy, long long lddy, int dy_dtype, const float* x, long long ldx,
                                const float* mean, const float* rstd, const float* gamma, const float* dres,
                                long long lddres, int M, int D, float* dx32, long long lddx32, void* dx16,
                                long long lddx16, float* dgamma, float* dbeta, const float* beta_act, void* stream) {
  if (M <= 0) return 0;
  CB_CHECK_ARG(D > 0 && D % 4 == 0 && D <= 2048, "layernorm_bwd: D=%d must be a multiple of 4 and <= 2048", D);
  CB_CHECK_ARG(lddy % 4 == 0 && ldx % 4 == 0, "layernorm_bwd: row pitches must be multiples of 4 elements");
  const int nv = pick_nv(D);
  // few, fat blocks: every block ends with 2*D global atomics for dgamma / dbeta
  const int blocks = (int)min((long long)(M + 7) / 8, (long long)cb_sm_count() * 2);
  const size_t smem = 2 * (size_t)D * sizeof(float);
  cudaStream_t s = (cudaStream_t)stream;
#define LN_BWD_CALL(BF)                                                                                              \
  ln_bwd_kernel<NV, BF><<<blocks, 256, smem, s>>>(dy, lddy, x, ldx, mean, rstd, gamma, dres, lddres, M, D, dx32, lddx32, \
                                                  (bf16*)dx16, lddx16, dgamma, dbeta, beta_act)
  const bool bf = dy_dtype == CB_DT_BF16;
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

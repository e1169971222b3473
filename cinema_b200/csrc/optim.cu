// Optimiser step over the flat parameter arena: global gradient norm, clipping, AdamW and the refresh
// of the bf16 weight shadow in ONE pass (reference: GradScaler.unscale_ + clip_grad_norm_(5.0) + AdamW,
// cinema/optim.py:173-226 and cinema/mae/pretrain.py:365-367, i.e. ~10 multi-tensor passes over
// ~600 tensors).  HBM-bound: 16 B read + 14 B written per parameter.
// Step-dependent scalars (lr, bias corrections) are read from device memory so that the launch
// can sit inside a CUDA graph.
#include "../../include/cinema_b200.h"
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  pdl_prologue();  // PDL: release the next launch, then wait for the previous kernel's results
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const float v = x[(n4 << 2) + threadIdx.x];
    acc += v * v;
  }
  __shared__ float part[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float s = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.f;
    s = warp_sum(s);
    if (threadIdx.x == 0) atomicAdd(out, s);
  }
}

// hyper = {lr, 1 - beta1^t, 1 - beta2^t}
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             bf16* __restrict__ p16, long long n, const float* __restrict__ hyper, float beta1, float beta2, float eps,
             float wd, const float* __restrict__ gnorm_sq, float max_norm, float grad_scale) {
  pdl_prologue();  // PDL: release the next launch, then wait for the previous kernel's results
  const float lr = __ldg(hyper), bc1 = __ldg(hyper + 1), bc2 = __ldg(hyper + 2);
  float gs = grad_scale;
  if (gnorm_sq != nullptr) {
    const float norm = sqrtf(__ldg(gnorm_sq)) * grad_scale;
    if (!isfinite(norm)) return;  // GradScaler semantics: skip the step on inf / nan gradients
    if (max_norm > 0.f) gs *= fminf(1.0f, max_norm / (norm + 1e-6f));
  }
  const float decay = 1.0f - lr * wd;
  const float step = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pp = &pv.x;
    const float* gp = &gv.x;
    float* mp = &mv.x;
    float* vp = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = gp[k] * gs;
      mp[k] = beta1 * mp[k] + (1.0f - beta1) * gk;
      vp[k] = beta2 * vp[k] + (1.0f - beta2) * gk * gk;
      const float denom = sqrtf(vp[k]) * inv_sqrt_bc2 + eps;
      pp[k] = pp[k] * decay - step * (mp[k] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (p16) reinterpret_cast<uint2*>(p16)[i] = make_uint2(pack_bf16(pv.x, pv.y), pack_bf16(pv.z, pv.w));
  }
}

inline int grid_for(long long n4) {
  long long b = (n4 + 255) / 256;
  const long long cap = (long long)cb_sm_count() * 8;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

extern "C" int cb_sumsq_f32(const float* x, long long n, float* out, void* stream) {
  if (n <= 0) return 0;
  CB_CHECK_ARG(((uintptr_t)x & 15) == 0, "sumsq: buffer must be 16-byte aligned");
  cb_launch(sumsq_kernel, grid_for(n >> 2), 256, 0, (cudaStream_t)stream, x, n, out);
  CB_LAUNCH_CHECK();
  return 0;
}

extern "C" int cb_adamw_flat(float* p, const float* g, float* m, float* v, void* p16, long long n, const float* hyper,
                             float beta1, float beta2, float eps, float weight_decay, const float* gnorm_sq,
                             float max_norm, float grad_scale, void* stream) {
  if (n <= 0) return 0;
  CB_CHECK_ARG(n % 4 == 0, "adamw: segment length %lld must be a multiple of 4", n);
  CB_CHECK_ARG((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0 && ((uintptr_t)p16 & 7) == 0,
               "adamw: buffers must be 16-byte aligned");
  cb_launch(adamw_kernel, grid_for(n >> 2), 256, 0, (cudaStream_t)stream, p, g, m, v, (bf16*)p16, n, hyper, beta1, beta2, eps,
                                                                  weight_decay, gnorm_sq, max_norm, grad_scale);
  CB_LAUNCH_CHECK();
  return 0;
}

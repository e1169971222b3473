// Segmentation loss of the fine-tuning scripts (cinema/segmentation/train.py:77-103): softmax cross-entropy
// (ignore_index = -1, mean over the labelled voxels) + MONAI DiceLoss(include_background=False, softmax=True) (per sample and
// foreground class: 1 - (2 sum p y + 1e-5) / (sum p + sum y + 1e-5), mean over (batch, class); ignored voxels count as
// background), over channel-first logits (B, C, *spatial).  The stock path runs ~12 elementwise / reduction kernels over the
// (B, C, S) probabilities and their one-hot targets; here the forward is ONE pass (per-voxel softmax in registers, block
// reductions, one fp32 atomic per accumulator and block) plus a one-block finalize that also prepares the backward's
// per-(sample, class) coefficients, and the backward is ONE pass that recomputes the softmax.  HBM-bound: C logits + one
// label per voxel in, C gradient values out.
#include "../../include/cinema_b200.h"
#include "common.cuh"

namespace {

constexpr int MAX_C = 8;

template <typename T>
__device__ __forceinline__ float ld_logit(const T* p, long long i) {
  return static_cast<float>(p[i]);
}
template <>
__device__ __forceinline__ float ld_logit<bf16>(const bf16* p, long long i) {
  return __bfloat162float(p[i]);
}
template <typename T>
__device__ __forceinline__ void st_grad(T* p, long long i, float v) {
  p[i] = static_cast<T>(v);
}
template <>
__device__ __forceinline__ void st_grad<bf16>(bf16* p, long long i, float v) {
  p[i] = __float2bfloat16(v);
}

__device__ __forceinline__ int load_label(const void* labels, int dtype, long long i) {
  switch (dtype) {
    case 0: return (int)reinterpret_cast<const long long*>(labels)[i];
    case 1: return reinterpret_cast<const int*>(labels)[i];
    case 2: return (int)reinterpret_cast<const short*>(labels)[i];
    default: return (int)reinterpret_cast<const unsigned char*>(labels)[i];
  }
}

// acc layout (fp32): [B][C][3] = {sum p y, sum p, sum y}, then [2] = {sum of -log p_y over labelled voxels, their count}
template <typename T>
__global__ void __launch_bounds__(256)
seg_loss_fwd_kernel(const T* __restrict__ logits, const void* __restrict__ labels, int label_dtype, int B, int C, long long S,
                    float* __restrict__ acc) {
  pdl_prologue();
  const int b = blockIdx.y;
  const T* lg = logits + (long long)b * C * S;
  float inter[MAX_C], psum[MAX_C], ysum[MAX_C];
#pragma unroll
  for (int c = 0; c < MAX_C; ++c) inter[c] = psum[c] = ysum[c] = 0.f;
  float ce = 0.f, cnt = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < S; s += stride) {
    float l[MAX_C];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < MAX_C; ++c)
      if (c < C) l[c] = ld_logit<T>(lg, (long long)c * S + s), m = fmaxf(m, l[c]);
    const int lab = load_label(labels, label_dtype, (long long)b * S + s);
    const int y = lab < 0 ? 0 : lab;  // labels.clamp(min=0) for the one-hot target of the Dice term
    float z = 0.f, ly = 0.f;
#pragma unroll
    for (int c = 0; c < MAX_C; ++c) {
      if (c < C) {
        const float d = l[c] - m;
        if (c == y) ly = d;
        l[c] = __expf(d), z += l[c];
      }
    }
    const float inv = 1.0f / z;
    if (lab >= 0) ce += logf(z) - ly, cnt += 1.f;  // -log softmax(l)[y]
#pragma unroll
    for (int c = 0; c < MAX_C; ++c) {
      if (c < C) {
        const float p = l[c] * inv;
        psum[c] += p;
        if (c == y) inter[c] += p, ysum[c] += 1.f;
      }
    }
  }
  // block reduction: warp shuffles, then one atomic per accumulator and warp
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int c = 0; c < MAX_C; ++c) {
    if (c < C) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        inter[c] += __shfl_xor_sync(0xffffffffu, inter[c], o);
        psum[c] += __shfl_xor_sync(0xffffffffu, psum[c], o);
        ysum[c] += __shfl_xor_sync(0xffffffffu, ysum[c], o);
      }
      if (lane == 0) {
        float* a = acc + ((long long)b * C + c) * 3;
        atomicAdd(a, inter[c]), atomicAdd(a + 1, psum[c]), atomicAdd(a + 2, ysum[c]);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ce += __shfl_xor_sync(0xffffffffu, ce, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if (lane == 0) {
    float* a = acc + (long long)B * C * 3;
    atomicAdd(a, ce), atomicAdd(a + 1, cnt);
  }
}

// out[0] = loss, out[1] = cross-entropy, out[2] = mean Dice loss; coef[B][C][2] = {a, b} with d dice / d p[b, c, s] =
// a * y + b, coef[B * C * 2] = 1 / labelled voxels (0 if none)
__global__ void seg_loss_finalize_kernel(const float* __restrict__ acc, int B, int C, float* __restrict__ out,
                                         float* __restrict__ coef) {
  pdl_prologue();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float eps = 1e-5f;
  const float n = (float)(B * (C - 1));
  float dice = 0.f;
  for (int b = 0; b < B; ++b) {
    for (int c = 0; c < C; ++c) {
      const float* a = acc + ((long long)b * C + c) * 3;
      float ca = 0.f, cb = 0.f;
      if (c > 0) {  // include_background = False
        const float num = 2.f * a[0] + eps, den = a[1] + a[2] + eps;
        dice += 1.f - num / den;
        ca = -2.f / (n * den);
        cb = num / (n * den * den);
      }
      coef[((long long)b * C + c) * 2] = ca;
      coef[((long long)b * C + c) * 2 + 1] = cb;
    }
  }
  dice = n > 0.f ? dice / n : 0.f;
  const float ce_sum = acc[(long long)B * C * 3], cnt = acc[(long long)B * C * 3 + 1];
  const float inv_cnt = cnt > 0.f ? 1.f / cnt : 0.f;
  const float ce = cnt > 0.f ? ce_sum * inv_cnt : nanf("");  // F.cross_entropy over zero labelled voxels is nan
  coef[(long long)B * C * 2] = inv_cnt;
  out[0] = dice + ce, out[1] = ce, out[2] = dice;
}

// dlogits[b, k, s] = gout * ( p_k (g_k - sum_c p_c g_c) + [labelled] (p_k - [k == y]) / n_labelled ),  g_c = a_bc [c == y] + b_bc
template <typename T>
__global__ void __launch_bounds__(256)
seg_loss_bwd_kernel(const T* __restrict__ logits, const void* __restrict__ labels, int label_dtype, int B, int C, long long S,
                    const float* __restrict__ coef, const float* __restrict__ gout, T* __restrict__ dlogits) {
  pdl_prologue();
  const int b = blockIdx.y;
  const T* lg = logits + (long long)b * C * S;
  T* dl = dlogits + (long long)b * C * S;
  float ca[MAX_C], cb[MAX_C];
#pragma unroll
  for (int c = 0; c < MAX_C; ++c)
    if (c < C) ca[c] = __ldg(coef + ((long long)b * C + c) * 2), cb[c] = __ldg(coef + ((long long)b * C + c) * 2 + 1);
  const float inv_cnt = __ldg(coef + (long long)B * C * 2);
  const float go = __ldg(gout);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < S; s += stride) {
    float p[MAX_C];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < MAX_C; ++c)
      if (c < C) p[c] = ld_logit<T>(lg, (long long)c * S + s), m = fmaxf(m, p[c]);
    float z = 0.f;
#pragma unroll
    for (int c = 0; c < MAX_C; ++c)
      if (c < C) p[c] = __expf(p[c] - m), z += p[c];
    const float inv = 1.0f / z;
    const int lab = load_label(labels, label_dtype, (long long)b * S + s);
    const int y = lab < 0 ? 0 : lab;
    const float w_ce = lab >= 0 ? inv_cnt : 0.f;
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < MAX_C; ++c)
      if (c < C) p[c] *= inv, dot += p[c] * (ca[c] * (c == y ? 1.f : 0.f) + cb[c]);
#pragma unroll
    for (int c = 0; c < MAX_C; ++c) {
      if (c < C) {
        const float hot = c == y ? 1.f : 0.f;
        const float g = ca[c] * hot + cb[c];
        st_grad<T>(dl, (long long)c * S + s, go * (p[c] * (g - dot) + w_ce * (p[c] - hot)));
      }
    }
  }
}

inline int seg_blocks(long long S) {
  long long b = (S + 255) / 256;
  const long long cap = (long long)cb_sm_count() * 2;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

extern "C" int cb_seg_loss_fwd(const void* logits, int logits_dtype, const void* labels, int label_dtype, int B, int C,
                               long long S, float* acc, float* out, float* coef, void* stream) {
  CB_CHECK_ARG(B > 0 && S > 0 && C >= 2 && C <= MAX_C, "seg_loss: need 2 <= C <= %d classes, B, S > 0 (C=%d)", MAX_C, C);
  CB_CHECK_ARG(logits_dtype == CB_DT_F32 || logits_dtype == CB_DT_BF16, "seg_loss: logits must be fp32 or bf16");
  CB_CHECK_ARG(label_dtype >= 0 && label_dtype <= 3, "seg_loss: label dtype code %d (0 int64, 1 int32, 2 int16, 3 uint8)", label_dtype);
  cudaStream_t s = (cudaStream_t)stream;
  CB_CUDA(cudaMemsetAsync(acc, 0, sizeof(float) * ((size_t)B * C * 3 + 2), s));
  dim3 grid(seg_blocks(S), B);
  if (logits_dtype == CB_DT_F32)
    cb_launch(seg_loss_fwd_kernel<float>, grid, 256, 0, s, (const float*)logits, labels, label_dtype, B, C, S, acc);
  else
    cb_launch(seg_loss_fwd_kernel<bf16>, grid, 256, 0, s, (const bf16*)logits, labels, label_dtype, B, C, S, acc);
  CB_LAUNCH_CHECK();
  cb_launch(seg_loss_finalize_kernel, 1, 32, 0, s, (const float*)acc, B, C, out, coef);
  CB_LAUNCH_CHECK();
  return 0;
}

extern "C" int cb_seg_loss_bwd(const void* logits, int logits_dtype, const void* labels, int label_dtype, int B, int C,
                               long long S, const float* coef, const float* grad_out, void* dlogits, void* stream) {
  CB_CHECK_ARG(B > 0 && S > 0 && C >= 2 && C <= MAX_C, "seg_loss: need 2 <= C <= %d classes, B, S > 0 (C=%d)", MAX_C, C);
  CB_CHECK_ARG(logits_dtype == CB_DT_F32 || logits_dtype == CB_DT_BF16, "seg_loss: logits must be fp32 or bf16");
  CB_CHECK_ARG(label_dtype >= 0 && label_dtype <= 3, "seg_loss: label dtype code %d", label_dtype);
  cudaStream_t s = (cudaStream_t)stream;
  dim3 grid(seg_blocks(S), B);
  if (logits_dtype == CB_DT_F32)
    cb_launch(seg_loss_bwd_kernel<float>, grid, 256, 0, s, (const float*)logits, labels, label_dtype, B, C, S, coef, grad_out,
              (float*)dlogits);
  else
    cb_launch(seg_loss_bwd_kernel<bf16>, grid, 256, 0, s, (const bf16*)logits, labels, label_dtype, B, C, S, coef, grad_out,
              (bf16*)dlogits);
  CB_LAUNCH_CHECK();
  return 0;
}

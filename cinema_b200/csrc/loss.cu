// Masked-pixel MSE of the MAE (cinema/mae/mae.py:107-152) fused with the target patchify
// (cinema/mae/mae.py:597): the image is read in place, one warp per patch; per-patch mean and
// unbiased std (the logged metrics, :129-137) come from the same pass, and the squared error is
// only evaluated on masked patches (:140-143).  Warp-shuffle reductions, one fp32 atomic per
// warp per accumulator.  HBM-bound: 4 B/pixel + 4 B/pred element in, 4 B/pred element out.
#include "../../include/cinema_b200.h"
#include "common.cuh"

namespace {

constexpr int MAX_ND = 4;

struct LossGeom {
  int ndim, C;
  int spatial[MAX_ND], patch[MAX_ND], grid[MAX_ND];
  long long sstride[MAX_ND];
  long long sc, sb;
  int n_tok, P;  // P = prod(patch); E = P * C
};

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  // valid for any sign: compare as ordered ints
  if (v >= 0.f)
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else
    atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

__global__ void __launch_bounds__(256) masked_mse_kernel(const float* __restrict__ image, LossGeom g, int B,
                                                         const unsigned char* __restrict__ mask,
                                                         const int* __restrict__ slot, const float* __restrict__ pred,
                                                         int n_drop, int norm_target, float eps, float* __restrict__ acc,
                                                         float* __restrict__ diff) {
  pdl_prologue();  // PDL: release the next launch, then wait for the previous kernel's results
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int E = g.P * g.C;
  const long long total = (long long)B * g.n_tok;
  float sq_acc = 0.f, mean_acc = 0.f, std_acc = 0.f, tmax = -INFINITY, pmax = -INFINITY;

  for (long long pt = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); pt < total;
       pt += (long long)gridDim.x * warps_per_block) {
    const int b = (int)(pt / g.n_tok);
    int tok = (int)(pt - (long long)b * g.n_tok);
    // origin of the patch inside the image
    long long base = b * g.sb;
    {
      int t = tok;
#pragma unroll
      for (int a = MAX_ND - 1; a >= 0; --a) {
        if (a < g.ndim) {
          base += (long long)((t % g.grid[a]) * g.patch[a]) * g.sstride[a];
          t /= g.grid[a];
        }
      }
    }
    // pass 1: mean.  token element e = off * C + c  (channel fastest, cinema/vit.py:112)
    float s = 0.f;
    for (int e = lane; e < E; e += 32) {
      const int c = e % g.C;
      int off = e / g.C;
      long long o = base + c * g.sc;
#pragma unroll
      for (int a = MAX_ND - 1; a >= 0; --a) {
        if (a < g.ndim) {
          o += (long long)(off % g.patch[a]) * g.sstride[a];
          off /= g.patch[a];
        }
      }
      s += __ldg(image + o);
    }
    const float mean = warp_sum(s) / (float)E;
    // pass 2: unbiased variance (torch.var default), data is L1/L2 resident by now
    float ss = 0.f;
    for (int e = lane; e < E; e += 32) {
      const int c = e % g.C;
      int off = e / g.C;
      long long o = base + c * g.sc;
#pragma unroll
      for (int a = MAX_ND - 1; a >= 0; --a) {
        if (a < g.ndim) {
          o += (long long)(off % g.patch[a]) * g.sstride[a];
          off /= g.patch[a];
        }
      }
      const float d = __ldg(image + o) - mean;
      ss += d * d;
    }
    const float var = warp_sum(ss) / (float)(E - 1);
    const float stdv = sqrtf(var);
    mean_acc += mean;
    std_acc += stdv;

    if (mask[pt] != 0) {
      const int j = slot[pt];
      const float* pr = pred + ((long long)b * n_drop + j) * E;
      float* df = diff ? diff + ((long long)b * n_drop + j) * E : nullptr;
      const float inv = 1.0f / (stdv + eps);
      for (int e = lane; e < E; e += 32) {
        const int c = e % g.C;
        int off = e / g.C;
        long long o = base + c * g.sc;
#pragma unroll
        for (int a = MAX_ND - 1; a >= 0; --a) {
          if (a < g.ndim) {
            o += (long long)(off % g.patch[a]) * g.sstride[a];
            off /= g.patch[a];
          }
        }
        float t = __ldg(image + o);
        if (norm_target) {
          t = (t - mean) * inv;
          tmax = fmaxf(tmax, t);
        }
        const float p = __ldg(pr + e);
        pmax = fmaxf(pmax, p);
        const float d = p - t;
        if (df) df[e] = d;
        sq_acc += d * d;
      }
    }
  }
  sq_acc = warp_sum(sq_acc);
  if (lane == 0) {
    atomicAdd(acc + 0, sq_acc);
    atomicAdd(acc + 1, mean_acc);  // mean / std accumulators are identical on all lanes
    atomicAdd(acc + 2, std_acc);
  }
  if (norm_target) {
    tmax = warp_max(tmax);
    pmax = warp_max(pmax);
    if (lane == 0) {
      if (tmax > -INFINITY) atomic_max_float(acc + 3, tmax);
      if (pmax > -INFINITY) atomic_max_float(acc + 4, pmax);
    }
  }
}

struct FinalizeArgs {
  int V;
  float sq_count[8];     // B * n_drop * E per view
  float patch_count[8];  // B * n_tok per view
};

// One thread: per-view mean squared error and metrics from the accumulators, the view-averaged loss
// over the views whose loss is finite (cinema/mae/mae.py:604-608, without the host sync), and the
// d loss / d pred scale of every view.
__global__ void mae_loss_finalize_kernel(const float* __restrict__ acc, FinalizeArgs a, float* __restrict__ out,
                                         float* __restrict__ scales) {
  pdl_prologue();  // PDL: release the next launch, then wait for the previous kernel's results
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float sum = 0.f;
  int n_fin = 0;
  for (int v = 0; v < a.V; ++v) {
    const float* ac = acc + v * 8;
    const float mse = ac[0] / a.sq_count[v];
    out[1 + v * 5 + 0] = mse;
    out[1 + v * 5 + 1] = ac[1] / a.patch_count[v];
    out[1 + v * 5 + 2] = ac[2] / a.patch_count[v];
    out[1 + v * 5 + 3] = ac[3];
    out[1 + v * 5 + 4] = ac[4];
    if (isfinite(mse)) sum += mse, ++n_fin;
  }
  out[0] = n_fin > 0 ? sum / (float)n_fin : __int_as_float(0x7fc00000);
  for (int v = 0; v < a.V; ++v) {
    const float mse = out[1 + v * 5 + 0];
    scales[v] = (isfinite(mse) && n_fin > 0) ? 2.0f / (a.sq_count[v] * (float)n_fin) : 0.f;
  }
}

}  // namespace

extern "C" int cb_mae_loss_finalize(const float* acc, int n_views, const float* sq_count, const float* patch_count,
                                    float* out, float* scales, void* stream) {
  CB_CHECK_ARG(n_views >= 1 && n_views <= 8, "loss_finalize: n_views %d not in 1..8", n_views);
  FinalizeArgs a;
  a.V = n_views;
  for (int v = 0; v < 8; ++v) a.sq_count[v] = a.patch_count[v] = 1.f;
  for (int v = 0; v < n_views; ++v) a.sq_count[v] = sq_count[v], a.patch_count[v] = patch_count[v];
  cb_launch(mae_loss_finalize_kernel, 1, 32, 0, (cudaStream_t)stream, acc, a, out, scales);
  CB_LAUNCH_CHECK();
  return 0;
}

extern "C" int cb_masked_mse_fwd(const float* image, int B, int C, int ndim, const int* spatial, const int* patch,
                                 const unsigned char* mask, const int* slot, const float* pred, int n_drop,
                                 int norm_target, float eps, float* acc, float* diff, void* stream) {
  CB_CHECK_ARG(ndim >= 1 && ndim <= MAX_ND, "masked_mse: ndim %d not in 1..4", ndim);
  LossGeom g;
  g.ndim = ndim, g.C = C, g.n_tok = 1, g.P = 1;
  for (int a = 0; a < MAX_ND; ++a) g.spatial[a] = g.patch[a] = g.grid[a] = 1, g.sstride[a] = 0;
  for (int a = 0; a < ndim; ++a) {
    CB_CHECK_ARG(patch[a] > 0 && spatial[a] % patch[a] == 0, "Input size (%d) cannot be divided by patch size (%d).",
                 spatial[a], patch[a]);
    g.spatial[a] = spatial[a], g.patch[a] = patch[a], g.grid[a] = spatial[a] / patch[a];
    g.n_tok *= g.grid[a], g.P *= patch[a];
  }
  long long s = 1;
  for (int a = ndim - 1; a >= 0; --a) g.sstride[a] = s, s *= g.spatial[a];
  g.sc = s, g.sb = s * C;
  CB_CHECK_ARG(g.P * C > 1, "masked_mse: patches of a single element have no unbiased variance");
  const long long total = (long long)B * g.n_tok;
  if (total <= 0) return 0;
  const int blocks = (int)min((total + 7) / 8, (long long)cb_sm_count() * 8);
  cb_launch(masked_mse_kernel, blocks, 256, 0, (cudaStream_t)stream, image, g, B, mask, slot, pred, n_drop, norm_target, eps,
                                                              acc, diff);
  CB_LAUNCH_CHECK();
  return 0;
}

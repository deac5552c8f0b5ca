// RGB-D loss glue of the tracker and the mapper, value + gradients in one launch.
//
// What it replaces (paths under /root/reference/src):
//   Mapper.py:553-562   depth_mask = gt_depth > 0;  loss = |gt_depth - depth|[depth_mask].sum()
//                       (+ w_color_loss * |gt_color - color|.sum() in the colour stage);  loss.backward()
//   Tracker.py:180-196  tmp = |gt_depth - depth| / sqrt(uncertainty.detach() + 1e-10)
//                       mask = (tmp < 10 * tmp.median()) & (gt_depth > 0)        (handle_dynamic; else gt_depth > 0)
//                       loss = tmp[mask].sum() + w_color_loss * |gt_color - color|[mask].sum();  loss.backward()
// ~25 eager launches each (abs, sqrt, median's sort, boolean-mask gathers with their host sync, sums and all the autograd
// transposes).  Everything is float64 as in the reference (depth / uncertainty are float64; float32 operands promote),
// the colour gradient is rounded to float32 where autograd casts it back to colour's dtype.
//
// One CTA: the batch is a few hundred to a few thousand rays.  The median (torch.median = the lower middle element) is an
// exact radix select over the bit patterns of the non-negative float64 residuals: 8 passes of an 8-bit histogram.
#include "ens_common.cuh"

namespace ens {

constexpr int RL_THREADS = 1024;

struct RgbdLossArgs {
  const double *depth, *var;
  const float *color, *gt_depth;
  const void *gt_color;
  int gt_color_f64;
  long long n;
  double w_color;
  int use_color, tracker, handle_dynamic;
  double *loss, *g_depth, *tmp;
  float *g_color;
};

__device__ __forceinline__ double block_sum(double v, double *sred) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x < 32) {
    s = threadIdx.x < RL_THREADS / 32 ? sred[threadIdx.x] : 0.0;
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  }
  return s;   // valid in warp 0
}

__device__ __forceinline__ double sgn(double x) { return (x > 0.0) - (x < 0.0); }

__global__ void __launch_bounds__(RL_THREADS) rgbd_loss_kernel(const RgbdLossArgs a) {
  __shared__ double sred[RL_THREADS / 32];
  __shared__ unsigned int hist[256];
  __shared__ unsigned long long s_prefix;
  __shared__ long long s_rank;
  __shared__ double s_thr;
  const long long n = a.n;

  double thr = 0.0;
  const bool gate = a.tracker && a.handle_dynamic && n > 0;
  if (a.tracker) {
    for (long long i = threadIdx.x; i < n; i += RL_THREADS)
      a.tmp[i] = fabs((double)a.gt_depth[i] - a.depth[i]) / sqrt(a.var[i] + 1e-10);
    __syncthreads();
  }
  if (gate) {
    // k-th smallest (k = (n-1)/2, zero-based) of tmp: non-negative doubles order like their bit patterns
    if (threadIdx.x == 0) { s_prefix = 0ull; s_rank = (n - 1) / 2; }
    for (int pass = 7; pass >= 0; --pass) {
      if (threadIdx.x < 256) hist[threadIdx.x] = 0u;
      __syncthreads();
      const unsigned long long prefix = s_prefix;
      const unsigned long long himask = pass == 7 ? 0ull : (~0ull << (8 * (pass + 1)));
      for (long long i = threadIdx.x; i < n; i += RL_THREADS) {
        const unsigned long long b = (unsigned long long)__double_as_longlong(a.tmp[i]);
        if ((b & himask) == prefix) atomicAdd(&hist[(b >> (8 * pass)) & 255ull], 1u);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        long long r = s_rank;
        int bin = 0;
        for (; bin < 256; ++bin) {
          if (r < (long long)hist[bin]) break;
          r -= hist[bin];
        }
        s_rank = r;
        s_prefix = prefix | ((unsigned long long)bin << (8 * pass));
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) s_thr = 10.0 * __longlong_as_double((long long)s_prefix);
    __syncthreads();
    thr = s_thr;
  }

  double part = 0.0;
  for (long long i = threadIdx.x; i < n; i += RL_THREADS) {
    const double gd = (double)a.gt_depth[i];
    const double diff = gd - a.depth[i];
    bool m = gd > 0.0;
    double gdep = 0.0;
    if (a.tracker) {
      const double t = a.tmp[i];
      if (gate) m = m && (t < thr);
      if (m) { part += t; gdep = -sgn(diff) / sqrt(a.var[i] + 1e-10); }
    } else if (m) {
      part += fabs(diff);
      gdep = -sgn(diff);
    }
    if (a.g_depth) a.g_depth[i] = gdep;
    if (a.use_color) {
      const bool mc = a.tracker ? m : true;                     // the mapper's colour term is not masked (Mapper.py:557)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const double gc = a.gt_color_f64 ? reinterpret_cast<const double *>(a.gt_color)[i * 3 + c]
                                         : (double)reinterpret_cast<const float *>(a.gt_color)[i * 3 + c];
        const double dc = gc - (double)a.color[i * 3 + c];
        if (mc) part += a.w_color * fabs(dc);
        if (a.g_color) a.g_color[i * 3 + c] = mc ? (float)(-a.w_color * sgn(dc)) : 0.f;
      }
    } else if (a.g_color) {
      a.g_color[i * 3 + 0] = 0.f; a.g_color[i * 3 + 1] = 0.f; a.g_color[i * 3 + 2] = 0.f;
    }
  }
  const double total = block_sum(part, sred);
  if (threadIdx.x == 0) a.loss[0] = total;
}

}  // namespace ens

using namespace ens;

extern "C" int ens_rgbd_loss(int tracker, const double *depth, const double *uncertainty, const float *color,
                             const float *gt_depth, const void *gt_color, int gt_color_is_f64, int64_t n, int use_color,
                             double w_color, int handle_dynamic, double *loss, double *g_depth, float *g_color,
                             double *workspace, ens_stream_t stream) {
  if (n < 0 || !loss) return ENS_EINVAL;
  if (n > 0 && (!depth || !gt_depth)) return ENS_EINVAL;
  if (n > 0 && use_color && (!color || !gt_color)) return ENS_EINVAL;
  if (n > 0 && tracker && (!uncertainty || !workspace)) return ENS_EINVAL;
  RgbdLossArgs a{};
  a.depth = depth; a.var = uncertainty; a.color = color; a.gt_depth = gt_depth; a.gt_color = gt_color;
  a.gt_color_f64 = gt_color_is_f64 ? 1 : 0;
  a.n = n; a.w_color = w_color; a.use_color = use_color ? 1 : 0; a.tracker = tracker ? 1 : 0;
  a.handle_dynamic = handle_dynamic ? 1 : 0;
  a.loss = loss; a.g_depth = g_depth; a.g_color = g_color; a.tmp = workspace;
  rgbd_loss_kernel<<<1, RL_THREADS, 0, (cudaStream_t)stream>>>(a);   // n == 0: writes loss = 0
  ENS_CHECK_CUDA();
  return ENS_OK;
}

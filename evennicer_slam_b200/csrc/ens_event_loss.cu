// Blurred-L2 event loss, value and gradient in one launch.
//
// What it replaces (paths under /root/reference/src): Tracker.py:204-224 and Mapper.py:593-615 --
//     loss  = ((gt - pred)**2).sum()
//     loss += sum_k w_k * ((gaussian_blur(gt, ks_k) - gaussian_blur(pred, ks_k))**2).sum()
//     loss *= balancer ; loss.backward()
// i.e. per Gaussian kernel two torchvision blurs (reflect pad + depthwise conv2d with the (ks x ks) outer-product kernel),
// a subtraction, a square, a sum, and the autograd transposes of all of them: ~30 launches on two 102 x 180 x 2 images.
// Blurring is linear, so blur(gt) - blur(pred) = blur(gt - pred): the kernel blurs the difference d once per kernel size.
//
//   forward   B_k(q)  = sum_t G_k(t) dpad(q + t),  dpad(p) = d(reflect(p)) on the image padded by r_k = ks_k / 2
//   loss      = balancer * ( sum d^2 + sum_k w_k sum_q B_k(q)^2 )
//   gradient  dL/dpred(q) = -balancer * ( 2 d(q) + sum_k 2 w_k sum_{p : reflect(p) = q} A_k(p) ),
//             A_k(p) = sum_t G_k(t) B_k(p - t)   (B_k = 0 outside the image) -- the transpose of conv2d, and the fold of
//             the padded border back onto the pixels it mirrors -- the transpose of the reflect pad.
//
// One CTA per 16 x 16 pixel tile and channel; the difference tile (with a 2 r halo) and the blurred tile (r halo) live in
// shared memory.  Launch-latency-bound at the reference's sizes (36 720 values): the point is one launch instead of thirty.
#include "ens_common.cuh"

namespace ens {

constexpr int EV_T = 16;          // tile edge
constexpr int EV_RMAX = 7;        // kernel_size <= 15
constexpr int EV_KMAX = 4;        // Gaussian kernels per call
constexpr int EV_DW = EV_T + 4 * EV_RMAX;
constexpr int EV_BW = EV_T + 2 * EV_RMAX;

struct EventLossArgs {
  const float *gt, *pred;         // [H][W][C]
  float *g_pred;                  // [H][W][C] or nullptr
  double *loss;                   // [2 + n_k]: total (scaled by balancer), unblurred sum d^2, blurred sums per kernel
  int H, W, C, n_k;
  int ks[EV_KMAX];
  float w[EV_KMAX];
  float k1d[EV_KMAX][2 * EV_RMAX + 1];
  float balancer;
};

__device__ __forceinline__ int reflect_idx(int p, int n) {     // torch 'reflect': -1 -> 1, n -> n-2 (valid for |overshoot| < n)
  if (p < 0) p = -p;
  if (p >= n) p = 2 * (n - 1) - p;
  return p;
}

__global__ void __launch_bounds__(EV_T * EV_T) event_loss_kernel(const EventLossArgs a) {
  __shared__ float sD[EV_DW * EV_DW];
  __shared__ float sB[EV_BW * EV_BW];
  __shared__ double sRed[EV_T * EV_T / 32];
  const int tx = threadIdx.x % EV_T, ty = threadIdx.x / EV_T;
  const int x0 = blockIdx.x * EV_T, y0 = blockIdx.y * EV_T, c = blockIdx.z;
  const int H = a.H, W = a.W, C = a.C;
  const int qx = x0 + tx, qy = y0 + ty;
  const bool inside = qx < W && qy < H;
  const float d_q = inside ? a.gt[((size_t)qy * W + qx) * C + c] - a.pred[((size_t)qy * W + qx) * C + c] : 0.f;
  float grad = 2.f * d_q;                                       // unblurred term
  double part[1 + EV_KMAX];
  part[0] = (double)d_q * (double)d_q;

  for (int k = 0; k < a.n_k; ++k) {
    const int r = a.ks[k] / 2;
    const float *g1 = a.k1d[k];
    // difference over the tile + 2r halo, in padded coordinates; beyond the r-wide pad nothing is ever read
    const int dw = EV_T + 4 * r;
    __syncthreads();
    for (int i = threadIdx.x; i < dw * dw; i += EV_T * EV_T) {
      const int py = y0 - 2 * r + i / dw, px = x0 - 2 * r + i % dw;
      float v = 0.f;
      if (py >= -r && py < H + r && px >= -r && px < W + r) {
        const int sy = reflect_idx(py, H), sx = reflect_idx(px, W);
        const size_t o = ((size_t)sy * W + sx) * C + c;
        v = a.gt[o] - a.pred[o];
      }
      sD[i] = v;
    }
    __syncthreads();
    // blurred difference over the tile + r halo (zero outside the image: conv2d's output domain)
    const int bw = EV_T + 2 * r;
    for (int i = threadIdx.x; i < bw * bw; i += EV_T * EV_T) {
      const int by = i / bw, bx = i % bw;
      const int py = y0 - r + by, px = x0 - r + bx;
      float acc = 0.f;
      if (py >= 0 && py < H && px >= 0 && px < W) {
        for (int v = 0; v <= 2 * r; ++v)
          for (int u = 0; u <= 2 * r; ++u)
            acc = fmaf(g1[v] * g1[u], sD[(by + v) * dw + bx + u], acc);     // sD origin is 2r before the tile
      }
      sB[i] = acc;
    }
    __syncthreads();
    const float b_q = inside ? sB[(ty + r) * bw + tx + r] : 0.f;
    part[1 + k] = (double)b_q * (double)b_q;
    if (inside) {
      // transpose of conv2d at q and at the padded positions that mirror onto q
      int ys[3], xs[3], ny = 0, nx = 0;
      ys[ny++] = qy;
      if (qy >= 1 && qy <= r) ys[ny++] = -qy;
      if (qy <= H - 2 && qy >= H - 1 - r) ys[ny++] = 2 * (H - 1) - qy;
      xs[nx++] = qx;
      if (qx >= 1 && qx <= r) xs[nx++] = -qx;
      if (qx <= W - 2 && qx >= W - 1 - r) xs[nx++] = 2 * (W - 1) - qx;
      float acc = 0.f;
      for (int iy = 0; iy < ny; ++iy)
        for (int ix = 0; ix < nx; ++ix) {
          const int py = ys[iy], px = xs[ix];
          for (int v = 0; v <= 2 * r; ++v) {
            const int by = py - (v - r);                        // B row read by tap v
            if (by < 0 || by >= H) continue;
            const int lby = by - (y0 - r);
            if (lby < 0 || lby >= bw) continue;                 // cannot happen for H, W > 2r (kept as a guard)
            for (int u = 0; u <= 2 * r; ++u) {
              const int bx = px - (u - r);
              if (bx < 0 || bx >= W) continue;
              const int lbx = bx - (x0 - r);
              if (lbx < 0 || lbx >= bw) continue;
              acc = fmaf(g1[v] * g1[u], sB[lby * bw + lbx], acc);
            }
          }
        }
      grad = fmaf(2.f * a.w[k], acc, grad);
    }
  }
  if (inside && a.g_pred) a.g_pred[((size_t)qy * W + qx) * C + c] = -a.balancer * grad;

  // loss parts: warp shuffle, then one atomic per CTA and part
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double total = 0.0;
  for (int j = 0; j <= a.n_k; ++j) {
    double v = part[j];
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) sRed[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int wv = 0; wv < EV_T * EV_T / 32; ++wv) s += sRed[wv];
      atomicAdd(&a.loss[1 + j], s);
      total += (j == 0 ? 1.0 : (double)a.w[j - 1]) * s;
    }
  }
  if (threadIdx.x == 0) atomicAdd(&a.loss[0], (double)a.balancer * total);
}

}  // namespace ens

using namespace ens;

extern "C" int ens_event_loss(const float *gt, const float *pred, int H, int W, int C, const int *kernel_sizes_host,
                              const float *kernels1d_host, const float *kernel_weights_host, int n_kernels,
                              float balancer, double *loss_parts, float *g_pred, ens_stream_t stream) {
  if (!gt || !pred || !loss_parts || H < 0 || W < 0 || C < 0 || n_kernels < 0) return ENS_EINVAL;
  if (n_kernels > EV_KMAX) return ENS_EUNSUPPORTED;
  if (n_kernels > 0 && (!kernel_sizes_host || !kernels1d_host || !kernel_weights_host)) return ENS_EINVAL;
  EventLossArgs a{};
  a.gt = gt; a.pred = pred; a.g_pred = g_pred; a.loss = loss_parts;
  a.H = H; a.W = W; a.C = C; a.n_k = n_kernels; a.balancer = balancer;
  int off = 0;
  for (int k = 0; k < n_kernels; ++k) {
    const int ks = kernel_sizes_host[k];
    if (ks < 1 || (ks & 1) == 0) return ENS_EINVAL;             // torchvision: odd, positive
    if (ks > 2 * EV_RMAX + 1) return ENS_EUNSUPPORTED;
    if (ks / 2 >= H || ks / 2 >= W) return ENS_ESHAPE;          // torch reflect pad needs pad < dim
    a.ks[k] = ks; a.w[k] = kernel_weights_host[k];
    for (int i = 0; i < ks; ++i) a.k1d[k][i] = kernels1d_host[off + i];
    off += ks;
  }
  cudaStream_t st = (cudaStream_t)stream;
  ENS_CUDA_CALL(cudaMemsetAsync(loss_parts, 0, sizeof(double) * (2 + n_kernels), st));
  if (H == 0 || W == 0 || C == 0) return ENS_OK;
  dim3 grid((W + EV_T - 1) / EV_T, (H + EV_T - 1) / EV_T, C);
  event_loss_kernel<<<grid, EV_T * EV_T, 0, st>>>(a);
  ENS_CHECK_CUDA();
  return ENS_OK;
}

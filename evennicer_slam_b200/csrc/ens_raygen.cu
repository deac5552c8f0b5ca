// Pixel draw gather + ray generation (common.get_samples / get_rays / get_rays_rescale) and its
// backward into c2w.  float32 with IEEE round-to-nearest per op and NO fma contraction, so the rays
// are bit-identical to the reference's eager `dirs * c2w[:3,:3]` / `torch.sum(-1)` (common.py:80-88).
#include "ens_common.cuh"

namespace ens {

struct Cam { float fx, fy, cx, cy; };

__device__ __forceinline__ void make_ray(float i, float j, Cam cam, const float *__restrict__ c2w, int ld,
                                         float *__restrict__ ro, float *__restrict__ rd) {
  // dirs = [(i-cx)/fx, -(j-cy)/fy, -1]                                   (common.py:82-83)
  const float d0 = __fdiv_rn(__fsub_rn(i, cam.cx), cam.fx);
  const float d1 = -__fdiv_rn(__fsub_rn(j, cam.cy), cam.fy);
  const float d2 = -1.0f;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    // torch.sum(dirs * c2w[:3,:3], -1): three rounded products summed left to right from 0
    const float p0 = __fmul_rn(d0, c2w[r * ld + 0]);
    const float p1 = __fmul_rn(d1, c2w[r * ld + 1]);
    const float p2 = __fmul_rn(d2, c2w[r * ld + 2]);
    rd[r] = __fadd_rn(__fadd_rn(p0, p1), p2);
    ro[r] = c2w[r * ld + 3];
  }
}

template <bool COLOR_F64>
__global__ void __launch_bounds__(256) sample_rays_kernel(const int64_t *__restrict__ indices, int64_t n, int H0, int W0,
                                                          int Wc, int W, Cam cam, const float *__restrict__ c2w, int ld,
                                                          const float *__restrict__ depth, const void *__restrict__ color,
                                                          float *__restrict__ pix_i, float *__restrict__ pix_j,
                                                          float *__restrict__ rays_o, float *__restrict__ rays_d,
                                                          float *__restrict__ out_depth, void *__restrict__ out_color) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int64_t idx = indices[t];
  const int row = (int)(idx / Wc), col = (int)(idx % Wc);
  const int px = W0 + col, py = H0 + row;         // i = W0+col, j = H0+row (common.py:136-140)
  const float fi = (float)px, fj = (float)py;
  if (pix_i) pix_i[t] = fi;
  if (pix_j) pix_j[t] = fj;
  float ro[3], rd[3];
  make_ray(fi, fj, cam, c2w, ld, ro, rd);
#pragma unroll
  for (int r = 0; r < 3; ++r) { rays_o[t * 3 + r] = ro[r]; rays_d[t * 3 + r] = rd[r]; }
  const int64_t pix = (int64_t)py * W + px;
  if (out_depth) out_depth[t] = depth[pix];
  if (out_color) {
    if (COLOR_F64) {
      const double *c = (const double *)color; double *o = (double *)out_color;
      o[t * 3 + 0] = c[pix * 3 + 0]; o[t * 3 + 1] = c[pix * 3 + 1]; o[t * 3 + 2] = c[pix * 3 + 2];
    } else {
      const float *c = (const float *)color; float *o = (float *)out_color;
      o[t * 3 + 0] = c[pix * 3 + 0]; o[t * 3 + 1] = c[pix * 3 + 1]; o[t * 3 + 2] = c[pix * 3 + 2];
    }
  }
}

__global__ void __launch_bounds__(256) lattice_rays_kernel(const float *__restrict__ lin_w, int nW,
                                                           const float *__restrict__ lin_h, int nH, Cam cam,
                                                           const float *__restrict__ c2w, int ld,
                                                           float *__restrict__ rays_o, float *__restrict__ rays_d) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (nH > 0 ? (int64_t)nW * nH : (int64_t)nW)) return;
  // nH == 0: "pairs" mode, ray t has pixel (lin_w[t], lin_h[t])  (get_rays_from_uv, common.py:74-89)
  const int r = nH > 0 ? (int)(t / nW) : (int)t, c = nH > 0 ? (int)(t % nW) : (int)t;
  float ro[3], rd[3];
  make_ray(lin_w[c], lin_h[r], cam, c2w, ld, ro, rd);
#pragma unroll
  for (int k = 0; k < 3; ++k) { rays_o[t * 3 + k] = ro[k]; rays_d[t * 3 + k] = rd[k]; }
}

// g_c2w[r][k] += sum_n g_d[n][r] * dirs[n][k] (k<3);  g_c2w[r][3] += sum_n g_o[n][r]     
// Block reduce in float64-free float32: 12 partial sums per thread -> warp shuffle -> one atomic per warp.
__global__ void __launch_bounds__(256) rays_bwd_kernel(const float *__restrict__ pix_i, const float *__restrict__ pix_j,
                                                       int64_t n, int nW, Cam cam, const float *__restrict__ g_o,
                                                       const float *__restrict__ g_d, float *__restrict__ g_c2w) {
  float acc[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) acc[k] = 0.f;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    float fi, fj;
    if (nW > 0) { fi = pix_i[t % nW]; fj = pix_j[t / nW]; } else { fi = pix_i[t]; fj = pix_j[t]; }
    const float d0 = __fdiv_rn(__fsub_rn(fi, cam.cx), cam.fx);
    const float d1 = -__fdiv_rn(__fsub_rn(fj, cam.cy), cam.fy);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float gd = g_d ? g_d[t * 3 + r] : 0.f;
      acc[r * 4 + 0] = fmaf(gd, d0, acc[r * 4 + 0]);
      acc[r * 4 + 1] = fmaf(gd, d1, acc[r * 4 + 1]);
      acc[r * 4 + 2] -= gd;
      acc[r * 4 + 3] += g_o ? g_o[t * 3 + r] : 0.f;
    }
  }
#pragma unroll
  for (int k = 0; k < 12; ++k) {
    float v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&g_c2w[k], v);
  }
}

// quad2rotation + get_camera_from_tensor (common.py:189-228) for n camera tensors [qw,qx,qy,qz,tx,ty,tz],
// and its backward.  One thread per camera; float32 with the reference's expression order.
__global__ void pose_fwd_kernel(const float *__restrict__ cam, int n, float *__restrict__ c2w) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const float qr = cam[t * 7 + 0], qi = cam[t * 7 + 1], qj = cam[t * 7 + 2], qk = cam[t * 7 + 3];
  const float nn = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(qr, qr), __fmul_rn(qi, qi)), __fmul_rn(qj, qj)), __fmul_rn(qk, qk));
  const float s = __fdiv_rn(2.0f, nn);
  float *o = c2w + t * 12;
  o[0] = __fsub_rn(1.f, __fmul_rn(s, __fadd_rn(__fmul_rn(qj, qj), __fmul_rn(qk, qk))));
  o[1] = __fmul_rn(s, __fsub_rn(__fmul_rn(qi, qj), __fmul_rn(qk, qr)));
  o[2] = __fmul_rn(s, __fadd_rn(__fmul_rn(qi, qk), __fmul_rn(qj, qr)));
  o[3] = cam[t * 7 + 4];
  o[4] = __fmul_rn(s, __fadd_rn(__fmul_rn(qi, qj), __fmul_rn(qk, qr)));
  o[5] = __fsub_rn(1.f, __fmul_rn(s, __fadd_rn(__fmul_rn(qi, qi), __fmul_rn(qk, qk))));
  o[6] = __fmul_rn(s, __fsub_rn(__fmul_rn(qj, qk), __fmul_rn(qi, qr)));
  o[7] = cam[t * 7 + 5];
  o[8] = __fmul_rn(s, __fsub_rn(__fmul_rn(qi, qk), __fmul_rn(qj, qr)));
  o[9] = __fmul_rn(s, __fadd_rn(__fmul_rn(qj, qk), __fmul_rn(qi, qr)));
  o[10] = __fsub_rn(1.f, __fmul_rn(s, __fadd_rn(__fmul_rn(qi, qi), __fmul_rn(qj, qj))));
  o[11] = cam[t * 7 + 6];
}

__global__ void pose_bwd_kernel(const float *__restrict__ cam, int n, const float *__restrict__ g, float *__restrict__ gcam) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const float qr = cam[t * 7 + 0], qi = cam[t * 7 + 1], qj = cam[t * 7 + 2], qk = cam[t * 7 + 3];
  const float nn = qr * qr + qi * qi + qj * qj + qk * qk;
  const float s = 2.0f / nn;
  const float *G = g + t * 12;
  // R = I*[1..] + s*M(q): dL/ds = sum G_ab * M_ab ; dL/dq = s * dM/dq contributions + dL/ds * ds/dq, ds/dq = -s^2 q... (= -2*2q/nn^2)
  const float m00 = -(qj * qj + qk * qk), m01 = qi * qj - qk * qr, m02 = qi * qk + qj * qr;
  const float m10 = qi * qj + qk * qr, m11 = -(qi * qi + qk * qk), m12 = qj * qk - qi * qr;
  const float m20 = qi * qk - qj * qr, m21 = qj * qk + qi * qr, m22 = -(qi * qi + qj * qj);
  const float gs = G[0] * m00 + G[1] * m01 + G[2] * m02 + G[4] * m10 + G[5] * m11 + G[6] * m12 + G[8] * m20 + G[9] * m21 + G[10] * m22;
  const float dsdq = -s * s;       // ds/dq_x = -2*2*q_x/nn^2 = -(s^2) * q_x
  float gr = s * (-G[1] * qk + G[2] * qj + G[4] * qk - G[6] * qi - G[8] * qj + G[9] * qi);
  float gi = s * (G[1] * qj + G[2] * qk + G[4] * qj - 2.f * G[5] * qi - G[6] * qr + G[8] * qk + G[9] * qr - 2.f * G[10] * qi);
  float gj = s * (-2.f * G[0] * qj + G[1] * qi + G[2] * qr + G[4] * qi + G[6] * qk - G[8] * qr + G[9] * qk - 2.f * G[10] * qj);
  float gk = s * (-2.f * G[0] * qk - G[1] * qr + G[2] * qi + G[4] * qr - 2.f * G[5] * qk + G[6] * qj + G[8] * qi + G[9] * qj);
  gr += gs * dsdq * qr; gi += gs * dsdq * qi; gj += gs * dsdq * qj; gk += gs * dsdq * qk;
  float *o = gcam + t * 7;
  o[0] = gr; o[1] = gi; o[2] = gj; o[3] = gk; o[4] = G[3]; o[5] = G[7]; o[6] = G[11];
}

}  // namespace ens

using namespace ens;

extern "C" int ens_pose_fwd(const float *cam_tensors, int n, float *c2w, ens_stream_t stream) {
  if (!cam_tensors || !c2w || n < 0) return ENS_EINVAL;
  if (n == 0) return ENS_OK;
  pose_fwd_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(cam_tensors, n, c2w);
  ENS_CHECK_CUDA();
  return ENS_OK;
}

extern "C" int ens_pose_bwd(const float *cam_tensors, int n, const float *g_c2w, float *g_cam, ens_stream_t stream) {
  if (!cam_tensors || !g_c2w || !g_cam || n < 0) return ENS_EINVAL;
  if (n == 0) return ENS_OK;
  pose_bwd_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(cam_tensors, n, g_c2w, g_cam);
  ENS_CHECK_CUDA();
  return ENS_OK;
}

extern "C" int ens_sample_rays(const int64_t *indices, int64_t n, int H0, int H1, int W0, int W1, int H, int W,
                               float fx, float fy, float cx, float cy, const float *c2w, int c2w_stride,
                               const float *depth, const void *color, int color_is_f64, float *pix_i, float *pix_j,
                               float *rays_o, float *rays_d, float *out_depth, void *out_color,
                               ens_stream_t stream) {
  if (!indices || !c2w || !rays_o || !rays_d || n < 0) return ENS_EINVAL;
  if (H0 < 0 || W0 < 0 || H1 > H || W1 > W || H1 <= H0 || W1 <= W0 || c2w_stride < 4) return ENS_ESHAPE;
  if ((out_depth && !depth) || (out_color && !color)) return ENS_EINVAL;
  if (n == 0) return ENS_OK;
  Cam cam{fx, fy, cx, cy};
  const unsigned g = (unsigned)((n + 255) / 256);
  cudaStream_t s = (cudaStream_t)stream;
  if (color_is_f64)
    sample_rays_kernel<true><<<g, 256, 0, s>>>(indices, n, H0, W0, W1 - W0, W, cam, c2w, c2w_stride, depth, color,
                                               pix_i, pix_j, rays_o, rays_d, out_depth, out_color);
  else
    sample_rays_kernel<false><<<g, 256, 0, s>>>(indices, n, H0, W0, W1 - W0, W, cam, c2w, c2w_stride, depth, color,
                                                pix_i, pix_j, rays_o, rays_d, out_depth, out_color);
  ENS_CHECK_CUDA();
  return ENS_OK;
}

extern "C" int ens_lattice_rays(const float *lin_w, int nW, const float *lin_h, int nH, float fx, float fy, float cx,
                                float cy, const float *c2w, int c2w_stride, float *rays_o, float *rays_d,
                                ens_stream_t stream) {
  if (!lin_w || !lin_h || !c2w || !rays_o || !rays_d) return ENS_EINVAL;
  if (nW <= 0 || nH < 0 || c2w_stride < 4) return ENS_ESHAPE;
  Cam cam{fx, fy, cx, cy};
  const int64_t n = nH > 0 ? (int64_t)nW * nH : (int64_t)nW;
  lattice_rays_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(lin_w, nW, lin_h, nH, cam, c2w,
                                                                                     c2w_stride, rays_o, rays_d);
  ENS_CHECK_CUDA();
  return ENS_OK;
}

extern "C" int ens_rays_bwd(const float *pix_i, const float *pix_j, int64_t n, int nW, float fx, float fy, float cx,
                            float cy, const float *g_rays_o, const float *g_rays_d, float *g_c2w,
                            ens_stream_t stream) {
  if (!pix_i || !pix_j || !g_c2w || n < 0) return ENS_EINVAL;
  if (n == 0) return ENS_OK;
  Cam cam{fx, fy, cx, cy};
  int64_t nb = (n + 255) / 256;
  if (nb > sm_count() * 4) nb = sm_count() * 4;
  rays_bwd_kernel<<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(pix_i, pix_j, n, nW, cam, g_rays_o, g_rays_d, g_c2w);
  ENS_CHECK_CUDA();
  return ENS_OK;
}

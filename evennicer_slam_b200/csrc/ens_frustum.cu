// Frustum feature selection and keyframe overlap of the mapper on the GPU.
// Replaces the numpy + cv2.remap host code of src/Mapper.py:115-186 (get_mask_from_c2w: 178 k - 759 k voxel centres per grid
// per mapped frame) and the per-keyframe projection loop of :222-241 (keyframe_selection_overlap).
// Arithmetic follows oracle/frustum_oracle.py operation by operation (rounding points forced with __fmul_rn / __fadd_rn /
// __dmul_rn / __dadd_rn so that nvcc contracts nothing into an FMA the host code does not have):
//   world -> camera  float32, products rounded one by one, (p0+p1)+(p2+p3)        (numpy float32 matmul)
//   K @ cam_cord     float64, u = fx*(-X) + cx*Z, v = fy*Y + cy*Z, z = Z + 1e-5;  uv = float32(u/z, v/z)
//   cv2.remap INTER_LINEAR, BORDER_CONSTANT 0: coordinates rounded to 1/32 pixel (round-half-even), int16 pixel index,
//                    weights (1-fy)(1-fx), (1-fy)fx, fy(1-fx), fy fx, float32 sum left to right
// HBM-bound trivia (a few MB); two launches because zero depths take the MAXIMUM of all interpolated depths (:166-168).
#include "ens_common.cuh"

namespace ens {

struct Cam6 { float H, W; double fx, fy, cx, cy; };
struct Mat34 { float m[12]; };

__device__ __forceinline__ void world_to_cam(const float *__restrict__ w, float x, float y, float z, float &X, float &Y, float &Z) {
  X = __fadd_rn(__fadd_rn(__fmul_rn(w[0], x), __fmul_rn(w[1], y)), __fadd_rn(__fmul_rn(w[2], z), w[3]));
  Y = __fadd_rn(__fadd_rn(__fmul_rn(w[4], x), __fmul_rn(w[5], y)), __fadd_rn(__fmul_rn(w[6], z), w[7]));
  Z = __fadd_rn(__fadd_rn(__fmul_rn(w[8], x), __fmul_rn(w[9], y)), __fadd_rn(__fmul_rn(w[10], z), w[11]));
}

__device__ __forceinline__ void project(const Cam6 &c, float X, float Y, float Z, float &u, float &v, double &z) {
  const double Xd = -(double)X, Yd = (double)Y, Zd = (double)Z;
  const double ud = __dadd_rn(__dmul_rn(c.fx, Xd), __dmul_rn(c.cx, Zd));
  const double vd = __dadd_rn(__dmul_rn(c.fy, Yd), __dmul_rn(c.cy, Zd));
  z = __dadd_rn(Zd, 1e-5);
  u = __double2float_rn(__ddiv_rn(ud, z));
  v = __double2float_rn(__ddiv_rn(vd, z));
}

// cvRound(x * 32) as x86 does it: round half to even; NaN and out-of-range give INT_MIN
__device__ __forceinline__ int cv_round32(float x) {
  const float s = __fmul_rn(x, 32.f);
  if (!(fabsf(s) < 2147483648.f)) return INT_MIN;
  return __float2int_rn(s);
}

__device__ __forceinline__ float remap_linear(const float *__restrict__ img, int H, int W, float u, float v) {
  const int sx = cv_round32(u), sy = cv_round32(v);
  int ix = sx >> 5, iy = sy >> 5;
  ix = max(-32768, min(32767, ix));
  iy = max(-32768, min(32767, iy));
  const float fx1 = (float)(sx & 31) * 0.03125f, fy1 = (float)(sy & 31) * 0.03125f;
  const float fx0 = 1.f - fx1, fy0 = 1.f - fy1;
  auto px = [&](int yy, int xx) -> float {
    return (xx >= 0 && xx < W && yy >= 0 && yy < H) ? __ldg(img + (size_t)yy * W + xx) : 0.f;
  };
  const float v0 = px(iy, ix), v1 = px(iy, ix + 1), v2 = px(iy + 1, ix), v3 = px(iy + 1, ix + 1);
  float acc = __fadd_rn(__fmul_rn(v0, __fmul_rn(fy0, fx0)), __fmul_rn(v1, __fmul_rn(fy0, fx1)));
  acc = __fadd_rn(acc, __fmul_rn(v2, __fmul_rn(fy1, fx0)));
  return __fadd_rn(acc, __fmul_rn(v3, __fmul_rn(fy1, fx1)));
}

// float max through integer atomics (any sign; NaN not handled -- a depth image has none)
__device__ __forceinline__ void atomic_max_float(float *addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned *>(addr), __float_as_uint(v));
}

// pass 1: interpolated depth per voxel centre (x slowest, as the reference's meshgrid) + their maximum
__global__ void __launch_bounds__(256) frustum_depth_kernel(Mat34 w2c, Cam6 cam, const float *__restrict__ xs, const float *__restrict__ ys,
                                                            const float *__restrict__ zs, int NX, int NY, int NZ,
                                                            const float *__restrict__ depth, int H, int W,
                                                            float *__restrict__ vox_depth, float *__restrict__ dmax) {
  const int64_t n = (int64_t)NX * NY * NZ;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float d = -INFINITY;
  if (t < n) {
    const int iz = (int)(t % NZ), iy = (int)((t / NZ) % NY), ix = (int)(t / ((int64_t)NZ * NY));
    float X, Y, Z, u, v;
    double z;
    world_to_cam(w2c.m, xs[ix], ys[iy], zs[iz], X, Y, Z);
    project(cam, X, Y, Z, u, v, z);
    d = remap_linear(depth, H, W, u, v);
    vox_depth[t] = d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) d = fmaxf(d, __shfl_xor_sync(0xffffffffu, d, o));
  if ((threadIdx.x & 31) == 0 && d > -INFINITY) atomic_max_float(dmax, d);
}

// pass 2: the selection.  mask_zyx != 0: mask[z][y][x] (the layout of the grids, what FrustumGridAdam takes), else [x][y][z]
// (the reference's return value).
__global__ void __launch_bounds__(256) frustum_mask_kernel(Mat34 w2c, Cam6 cam, float ox, float oy, float oz, const float *__restrict__ xs,
                                                           const float *__restrict__ ys, const float *__restrict__ zs, int NX, int NY, int NZ,
                                                           const float *__restrict__ vox_depth, const float *__restrict__ dmax,
                                                           int mask_zyx, uint8_t *__restrict__ mask) {
  const int64_t n = (int64_t)NX * NY * NZ;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  int ix, iy, iz;
  if (mask_zyx) { ix = (int)(t % NX); iy = (int)((t / NX) % NY); iz = (int)(t / ((int64_t)NX * NY)); }
  else { iz = (int)(t % NZ); iy = (int)((t / NZ) % NY); ix = (int)(t / ((int64_t)NZ * NY)); }
  const float x = xs[ix], y = ys[iy], zc = zs[iz];
  float X, Y, Z, u, v;
  double z;
  world_to_cam(w2c.m, x, y, zc, X, Y, Z);
  project(cam, X, Y, Z, u, v, z);
  float d = vox_depth[((int64_t)ix * NY + iy) * NZ + iz];
  if (d == 0.f) d = *dmax;
  bool m = (u < cam.W) && (u > 0.f) && (v < cam.H) && (v > 0.f);
  m = m && (0.0 <= -z) && (-z <= (double)__fadd_rn(d, 0.5f));
  const float dx = __fsub_rn(x, ox), dy = __fsub_rn(y, oy), dz = __fsub_rn(zc, oz);
  const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
  m = m || (d2 < 0.25f);
  mask[t] = m ? 1 : 0;
}

// one CTA per keyframe: how many of the current frame's sample points project inside its image (Mapper.py:222-241)
__global__ void __launch_bounds__(256) keyframe_overlap_kernel(const float *__restrict__ w2cs, Cam6 cam, float edge,
                                                               const float *__restrict__ verts, int n, int *__restrict__ counts) {
  __shared__ float w[12];
  __shared__ int total;
  if (threadIdx.x < 12) w[threadIdx.x] = w2cs[(size_t)blockIdx.x * 16 + threadIdx.x];
  if (threadIdx.x == 0) total = 0;
  __syncthreads();
  int c = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float X, Y, Z, u, v;
    double z;
    world_to_cam(w, verts[3 * i], verts[3 * i + 1], verts[3 * i + 2], X, Y, Z);
    project(cam, X, Y, Z, u, v, z);
    c += (u < cam.W - edge) && (u > edge) && (v < cam.H - edge) && (v > edge) && (z < 0.0);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&total, c);
  __syncthreads();
  if (threadIdx.x == 0) counts[blockIdx.x] = total;
}

}  // namespace ens

using namespace ens;

extern "C" int64_t ens_frustum_workspace_bytes(int NX, int NY, int NZ) {
  if (NX < 1 || NY < 1 || NZ < 1) return 0;
  return ((int64_t)NX * NY * NZ + 64) * (int64_t)sizeof(float);
}

extern "C" int ens_frustum_mask(const float *w2c_host, const float *cam_centre_host, const double *cam6_host, const float *xs,
                                const float *ys, const float *zs, int NX, int NY, int NZ, const float *depth, int H, int W,
                                int mask_zyx, uint8_t *mask, void *workspace, int64_t workspace_bytes, ens_stream_t stream) {
  if (!w2c_host || !cam_centre_host || !cam6_host || !xs || !ys || !zs || !depth || !mask || !workspace) return ENS_EINVAL;
  if (NX < 1 || NY < 1 || NZ < 1 || H < 1 || W < 1 || H >= 32767 || W >= 32767) return ENS_ESHAPE;
  if (workspace_bytes < ens_frustum_workspace_bytes(NX, NY, NZ)) return ENS_ESHAPE;
  Mat34 m;
  for (int i = 0; i < 12; ++i) m.m[i] = w2c_host[i];
  Cam6 c{(float)cam6_host[0], (float)cam6_host[1], cam6_host[2], cam6_host[3], cam6_host[4], cam6_host[5]};
  const int64_t n = (int64_t)NX * NY * NZ;
  float *vox_depth = reinterpret_cast<float *>(workspace);
  float *dmax = vox_depth + n;
  cudaStream_t s = (cudaStream_t)stream;
  // seed of the running maximum: cudaMemsetAsync takes a byte, and -inf (0xff800000) is not byte-uniform; 0xfefefefe is
  // -1.69e38, below any interpolated depth
  ENS_CUDA_CALL(cudaMemsetAsync(dmax, 0xfe, 4, s));
  const unsigned nb = (unsigned)((n + 255) / 256);
  frustum_depth_kernel<<<nb, 256, 0, s>>>(m, c, xs, ys, zs, NX, NY, NZ, depth, H, W, vox_depth, dmax);
  ENS_CHECK_CUDA();
  frustum_mask_kernel<<<nb, 256, 0, s>>>(m, c, cam_centre_host[0], cam_centre_host[1], cam_centre_host[2], xs, ys, zs, NX, NY, NZ,
                                         vox_depth, dmax, mask_zyx, mask);
  ENS_CHECK_CUDA();
  return ENS_OK;
}

extern "C" int ens_keyframe_overlap(const float *w2cs, int n_keyframes, const double *cam6_host, float edge, const float *vertices,
                                    int n_vertices, int *counts, ens_stream_t stream) {
  if (!w2cs || !cam6_host || !vertices || !counts) return ENS_EINVAL;
  if (n_keyframes < 0 || n_vertices < 0) return ENS_ESHAPE;
  if (n_keyframes == 0) return ENS_OK;
  Cam6 c{(float)cam6_host[0], (float)cam6_host[1], cam6_host[2], cam6_host[3], cam6_host[4], cam6_host[5]};
  keyframe_overlap_kernel<<<n_keyframes, 256, 0, (cudaStream_t)stream>>>(w2cs, c, edge, vertices, n_vertices, counts);
  ENS_CHECK_CUDA();
  return ENS_OK;
}

// Layout plumbing: grid layout conversion, decoder packing, batch depth maximum.
#include "ens_common.cuh"

namespace ens {

// [32][V] <-> [V][32] through a padded 32x32 shared tile; both sides fully coalesced.
// HBM-bound: 2 x 4 bytes per element.  Grid = ceil(V/32) CTAs of 32x8 threads.
__global__ void __launch_bounds__(256) grid_to_native_kernel(const float *__restrict__ src, float *__restrict__ dst,
                                                             int64_t V) {
  __shared__ float tile[32][33];
  const int64_t v0 = (int64_t)blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
  for (int c = ty; c < 32; c += 8) {
    int64_t v = v0 + tx;
    tile[c][tx] = (v < V) ? src[(int64_t)c * V + v] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    int64_t v = v0 + r;
    if (v < V) dst[v * 32 + tx] = tile[tx][r];
  }
}

template <bool ACC>
__global__ void __launch_bounds__(256) grid_from_native_kernel(const float *__restrict__ src, float *__restrict__ dst,
                                                               int64_t V) {
  __shared__ float tile[32][33];
  const int64_t v0 = (int64_t)blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    int64_t v = v0 + r;
    tile[r][tx] = (v < V) ? src[v * 32 + tx] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int c = ty; c < 32; c += 8) {
    int64_t v = v0 + tx;
    if (v < V) {
      float x = tile[tx][c];
      if (ACC) dst[(int64_t)c * V + v] += x; else dst[(int64_t)c * V + v] = x;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// decoder packing: one thread per packed float, pulling from the per-tensor pointers
// ---------------------------------------------------------------------------------------------
struct PtrTable { const float *p[24]; };

template <int CD, int NO>
__device__ __forceinline__ void pack_mlp_body(const PtrTable &t, float *__restrict__ out, int idx) {
  using P = MlpPack<CD>;
  float v = 0.f;
  // state_dict order: fc_c.i.weight (2i), fc_c.i.bias (2i+1), _B (10), pts.i.weight (11+2i), pts.i.bias (12+2i),
  // out.weight (21), out.bias (22)
  if (idx < 3 * EMBP) {
    int r = idx / EMBP, k = idx % EMBP;
    v = (k < EMB) ? t.p[10][r * EMB + k] : 0.f;
  } else if (idx >= P::off_Wo()) {
    int o = idx - P::off_Wo();
    if (o < 128) {
      int j = o / 4, n = o % 4;
      v = (n < NO) ? t.p[21][n * 32 + j] : 0.f;
    } else {
      int n = o - 128;
      v = (n < NO) ? t.p[22][n] : 0.f;
    }
  } else {
    int i = 0;
#pragma unroll
    for (int l = 1; l < 5; ++l) if (idx >= P::off_layer(l)) i = l;
    int o = idx - P::off_layer(i);
    const int K = P::K(i);
    if (o < K * 32) {                       // W_iT[k][j] = W_i[j][k]
      int k = o / 32, j = o % 32;
      v = t.p[11 + 2 * i][j * K + k];
    } else if (o < K * 32 + 32) {
      v = t.p[12 + 2 * i][o - K * 32];
    } else if (o < K * 32 + 32 + CD * 32) {  // Wc_iT[k][j] = Wc_i[j][k]
      int q = o - K * 32 - 32;
      int k = q / 32, j = q % 32;
      v = t.p[2 * i][j * CD + k];
    } else {
      v = t.p[2 * i + 1][o - K * 32 - 32 - CD * 32];
    }
  }
  out[idx] = v;
}

// mma layout (see MlpPackV2): dense [out][in], XOR-swizzled columns
template <int CD, int NO>
__device__ __forceinline__ void pack_mlp_v2_body(const PtrTable &t, float *__restrict__ out, int idx) {
  using P = MlpPackV2<CD>;
  float v = 0.f;
  if (idx < P::off_W0()) {
    int r = idx / EMBP, k = idx % EMBP;
    v = (k < EMB) ? t.p[10][r * EMB + k] : 0.f;
  } else if (idx < P::off_L(0)) {
    const bool is3 = idx >= P::off_W3e();
    int o = idx - (is3 ? P::off_W3e() : P::off_W0());
    int n = o / EMBP, k = swz(n, o % EMBP);           // physical column -> logical k (swz is an involution)
    if (k < EMB) v = is3 ? t.p[11 + 2 * 3][n * 125 + k] : t.p[11][n * EMB + k];
  } else if (idx < P::off_Wo()) {
    int o = idx - P::off_L(0);
    const int i = o / P::block_floats();
    o -= i * P::block_floats();
    if (o < P::in_b()) {
      int n = o / 32, k = swz(n, o % 32);
      if (i == 3) v = t.p[11 + 2 * 3][n * 125 + EMB + k];
      else if (i > 0) v = t.p[11 + 2 * i][n * 32 + k];
    } else if (o < P::in_Wc()) {
      v = t.p[12 + 2 * i][o - P::in_b()];
    } else if (o < P::in_bc()) {
      int q = o - P::in_Wc();
      int n = q / CD, k = swz(n, q % CD);
      v = t.p[2 * i][n * CD + k];
    } else {
      v = t.p[2 * i + 1][o - P::in_bc()];
    }
  } else {
    int o = idx - P::off_Wo();
    if (o < 128) {
      int n = o / 32, j = o % 32;
      v = (n < NO) ? t.p[21][n * 32 + j] : 0.f;
    } else {
      int n = o - 128;
      v = (n < NO) ? t.p[22][n] : 0.f;
    }
  }
  out[idx] = v;
}

// mma backward layout (see MlpPackV2B): transposed matrices, [in][32], XOR-swizzled columns
template <int CD, int NO>
__device__ __forceinline__ void pack_mlp_v2b_body(const PtrTable &t, float *__restrict__ out, int idx) {
  using P = MlpPackV2B;
  float v = 0.f;
  if (idx < P::off_Wo()) {
    int r = idx / EMBP, k = idx % EMBP;
    v = (k < EMB) ? t.p[10][r * EMB + k] : 0.f;
  } else if (idx < P::off_W0T()) {
    int o = idx - P::off_Wo();
    int n = o / 32, j = o % 32;
    v = (n < NO) ? t.p[21][n * 32 + j] : 0.f;
  } else if (idx < P::off_L(0)) {
    const bool is3 = idx >= P::off_W3eT();
    int o = idx - (is3 ? P::off_W3eT() : P::off_W0T());
    int k = o / 32, n = swz(k, o % 32);               // row k (input feature), logical column n (output)
    if (k < EMB) v = is3 ? t.p[11 + 2 * 3][n * 125 + k] : t.p[11][n * EMB + k];
  } else {
    int o = idx - P::off_L(0);
    const int i = o / 2048;
    o -= i * 2048;
    if (o < 1024) {
      int k = o / 32, n = swz(k, o % 32);
      if (i == 3) v = t.p[11 + 2 * 3][n * 125 + EMB + k];
      else if (i > 0) v = t.p[11 + 2 * i][n * 32 + k];
    } else {
      o -= 1024;
      int c = o / 32, n = swz(c, o % 32);
      v = t.p[2 * i][n * CD + c];
    }
  }
  out[idx] = v;
}

// tcgen05 layout (see MlpPackTC): canonical UMMA K-major core matrices, value and tf32 remainder; the feature
// matrices are pre-multiplied into the next block (float64 accumulation, rounded once)
template <int CD, int NO>
__device__ __forceinline__ void pack_mlp_tc_body(const PtrTable &t, float *__restrict__ out, int idx) {
  using P = MlpPackTC<CD>;
  // hidden part of pts_linears[i].weight, i = 1..4
  auto Wh = [&](int i, int n, int j) { return (i == 3) ? t.p[11 + 2 * 3][n * 125 + EMB + j] : t.p[11 + 2 * i][n * 32 + j]; };
  float v = 0.f;
  if (idx < 2 * P::TOT()) {
    const bool lo = idx >= P::TOT();
    int o = lo ? idx - P::TOT() : idx;
    int K, mat, i = 0;
    if (o < P::off_W3e()) { mat = 0; K = EMBP; }
    else if (o < P::off_Wh(1)) { mat = 1; K = EMBP; o -= P::off_W3e(); }
    else if (o < P::off_M(0)) { mat = 2; K = 32; o -= P::off_Wh(1); i = 1 + o / 1024; o %= 1024; }
    else { mat = 3; K = CD; o -= P::off_M(0); i = o / (32 * CD); o %= (32 * CD); }
    const int ni = o / ((K / 4) * 32), r1 = o % ((K / 4) * 32);
    const int ki = r1 / 32, r2 = r1 % 32;
    const int n = ni * 8 + r2 / 4, k = ki * 4 + r2 % 4;
    float w = 0.f;
    if (mat == 0) w = (k < EMB) ? t.p[11][n * EMB + k] : 0.f;
    else if (mat == 1) w = (k < EMB) ? t.p[11 + 2 * 3][n * 125 + k] : 0.f;
    else if (mat == 2) w = Wh(i, n, k);
    else {                                       // M_i[n][k] = sum_j Wh_{i+1}[n][j] Wc_i[j][k]
      double acc = 0.0;
      for (int j = 0; j < 32; ++j) acc += (double)Wh(i + 1, n, j) * (double)t.p[2 * i][j * CD + k];
      w = (float)acc;
    }
    const float hi = __uint_as_float(__float_as_uint(w) & 0xffffe000u);
    v = lo ? (w - hi) : w;
  } else {
    int o = idx - P::off_B();
    if (o < 3 * EMBP) { int r = o / EMBP, k = o % EMBP; v = (k < EMB) ? t.p[10][r * EMB + k] : 0.f; }
    else if (o < 3 * EMBP + 160) {               // b'_i
      o -= 3 * EMBP;
      const int i = o / 32, n = o % 32;
      double acc = (double)t.p[12 + 2 * i][n];
      if (i > 0) for (int j = 0; j < 32; ++j) acc += (double)Wh(i, n, j) * (double)t.p[2 * (i - 1) + 1][j];
      v = (float)acc;
    } else if (o < 3 * EMBP + 160 + 128) { o -= 3 * EMBP + 160; int n = o / 32, j = o % 32; v = (n < NO) ? t.p[21][n * 32 + j] : 0.f; }
    else if (o < 3 * EMBP + 160 + 128 + 4 * CD) {   // Mo[q][k] = sum_j Wo[q][j] Wc_4[j][k]
      o -= 3 * EMBP + 160 + 128;
      const int q = o / CD, k = o % CD;
      double acc = 0.0;
      if (q < NO) for (int j = 0; j < 32; ++j) acc += (double)t.p[21][q * 32 + j] * (double)t.p[2 * 4][j * CD + k];
      v = (float)acc;
    } else {                                     // bo'[q] = bo[q] + sum_j Wo[q][j] bc_4[j]
      const int q = o - (3 * EMBP + 160 + 128 + 4 * CD);
      double acc = 0.0;
      if (q < NO) { acc = (double)t.p[22][q]; for (int j = 0; j < 32; ++j) acc += (double)t.p[21][q * 32 + j] * (double)t.p[2 * 4 + 1][j]; }
      v = (float)acc;
    }
  }
  out[idx] = v;
}

// tcgen05 backward layout (see MlpPackTCB): transposed matrices, canonical K-major (K = 32 output units), value + remainder
template <int CD, int NO>
__device__ __forceinline__ void pack_mlp_tcb_body(const PtrTable &t, float *__restrict__ out, int idx) {
  using P = MlpPackTCB;
  auto Wh = [&](int i, int n, int j) { return (i == 3) ? t.p[11 + 2 * 3][n * 125 + EMB + j] : t.p[11 + 2 * i][n * 32 + j]; };
  float v = 0.f;
  if (idx < 2 * P::TOT()) {
    const bool lo = idx >= P::TOT();
    int o = lo ? idx - P::TOT() : idx;
    int mat = -1, i = 0;          // mat 0: WhT_i, 1: MT_i, 2: W0T, 3: W3eT
    if (o >= P::off_W0T()) { mat = 2; o -= P::off_W0T(); }
    else if (o >= P::off_W3eT() && o < P::off_G(2)) { mat = 3; o -= P::off_W3eT(); }
    else {
      for (int q = 1; q <= 4; ++q) {
        if (o >= P::off_WhT(q) && o < P::off_WhT(q) + 1024) { mat = 0; i = q; o -= P::off_WhT(q); break; }
        if (o >= P::off_MT(q - 1) && o < P::off_MT(q - 1) + 1024) { mat = 1; i = q - 1; o -= P::off_MT(q - 1); break; }
      }
    }
    // canonical [rows][K = 32]: row group of 8 = 256 floats, k group of 4 = 32 floats
    const int ri = o / 256, r1 = o % 256;
    const int ki = r1 / 32, r2 = r1 % 32;
    const int row = ri * 8 + r2 / 4, n = ki * 4 + r2 % 4;            // row = input index of the layer, n = output unit
    float w = 0.f;
    if (mat == 0) w = Wh(i, n, row);
    else if (mat == 1) {                                            // M_i[n][row] = sum_j Wh_{i+1}[n][j] Wc_i[j][row]
      double acc = 0.0;
      for (int j = 0; j < 32; ++j) acc += (double)Wh(i + 1, n, j) * (double)t.p[2 * i][j * CD + row];
      w = (float)acc;
    } else if (mat == 2) w = (row < EMB) ? t.p[11][n * EMB + row] : 0.f;
    else w = (row < EMB) ? t.p[11 + 2 * 3][n * 125 + row] : 0.f;
    const float hi = __uint_as_float(__float_as_uint(w) & 0xffffe000u);
    v = lo ? (w - hi) : w;
  } else {
    int o = idx - P::off_B();
    if (o < 3 * EMBP) { int r = o / EMBP, k = o % EMBP; v = (k < EMB) ? t.p[10][r * EMB + k] : 0.f; }
    else if (o < 3 * EMBP + 128) { o -= 3 * EMBP; int n = o / 32, j = o % 32; v = (n < NO) ? t.p[21][n * 32 + j] : 0.f; }
    else {                                                          // MoF[q][k] = sum_j Wo[q][j] Wc_4[j][k], k < 32
      o -= 3 * EMBP + 128;
      const int q = o / 32, k = o % 32;
      double acc = 0.0;
      if (q < NO) for (int j = 0; j < 32; ++j) acc += (double)t.p[21][q * 32 + j] * (double)t.p[2 * 4][j * CD + k];
      v = (float)acc;
    }
  }
  out[idx] = v;
}

// one launch packs all sections of a decoder blob (fma | mma forward | mma backward | tcgen05 | tcgen05 backward)
template <int CD, int NO>
__global__ void pack_mlp_all_kernel(PtrTable t, float *__restrict__ out) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < MlpPack<CD>::total()) { pack_mlp_body<CD, NO>(t, out, idx); return; }
  idx -= MlpPack<CD>::total();
  if (idx < MlpPackV2<CD>::total()) { pack_mlp_v2_body<CD, NO>(t, out + off_v2<CD>(), idx); return; }
  idx -= MlpPackV2<CD>::total();
  if (idx < MlpPackV2B::total()) { pack_mlp_v2b_body<CD, NO>(t, out + off_v2b<CD>(), idx); return; }
  idx -= MlpPackV2B::total();
  if (idx < MlpPackTC<CD>::total()) { pack_mlp_tc_body<CD, NO>(t, out + off_tc<CD>(), idx); return; }
  idx -= MlpPackTC<CD>::total();
  if (idx < MlpPackTCB::total()) pack_mlp_tcb_body<CD, NO>(t, out + off_tcb<CD>(), idx);
}

__global__ void pack_coarse_kernel(PtrTable t, float *__restrict__ out) {
  using P = CoarsePack;
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P::total()) return;
  float v = 0.f;
  // state_dict order: pts.i.weight (2i), pts.i.bias (2i+1), out.weight (10), out.bias (11)
  if (idx >= P::off_Wo()) {
    int o = idx - P::off_Wo();
    if (o < 128) {
      int j = o / 4, n = o % 4;
      v = (n == 0) ? t.p[10][j] : 0.f;
    } else {
      v = (o == 128) ? t.p[11][0] : 0.f;
    }
  } else {
    int i = 0;
#pragma unroll
    for (int l = 1; l < 5; ++l) if (idx >= P::off_W(l)) i = l;
    int o = idx - P::off_W(i);
    const int K = P::K(i);
    if (o < K * 32) {
      int k = o / 32, j = o % 32;
      v = t.p[2 * i][j * K + k];
    } else {
      v = t.p[2 * i + 1][o - K * 32];
    }
  }
  out[idx] = v;
}

// ---------------------------------------------------------------------------------------------
// max(gt_depth): one CTA, grid-stride.  NaN-propagating like torch.max.  n is at most a few 100k.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) depth_max_kernel(const float *__restrict__ d, int64_t n,
                                                         double *__restrict__ out) {
  __shared__ float red[32];
  __shared__ int nan_seen;
  if (threadIdx.x == 0) nan_seen = 0;
  __syncthreads();
  float m = -INFINITY;
  bool nan = false;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    float v = d[i];
    nan |= (v != v);
    m = fmaxf(m, v);
  }
  if (nan) nan_seen = 1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : -INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) {
      if (nan_seen) m = NAN;
      out[0] = (double)__fmul_rn(m, 1.2f);   // max(gt_depth*1.2): float32 product, monotone in gt_depth
      out[1] = (double)m;
    }
  }
}

}  // namespace ens

using namespace ens;

extern "C" int ens_grid_to_native(const float *ref_layout, float *native, int64_t n_vox, ens_stream_t stream) {
  if (!ref_layout || !native || n_vox < 0) return ENS_EINVAL;
  if (n_vox == 0) return ENS_OK;
  dim3 blk(32, 8);
  grid_to_native_kernel<<<(unsigned)((n_vox + 31) / 32), blk, 0, (cudaStream_t)stream>>>(ref_layout, native, n_vox);
  ENS_CHECK_CUDA();
  return ENS_OK;
}

extern "C" int ens_grid_from_native(const float *native, float *ref_layout, int64_t n_vox, int accumulate,
                                    ens_stream_t stream) {
  if (!ref_layout || !native || n_vox < 0) return ENS_EINVAL;
  if (n_vox == 0) return ENS_OK;
  dim3 blk(32, 8);
  unsigned g = (unsigned)((n_vox + 31) / 32);
  if (accumulate)
    grid_from_native_kernel<true><<<g, blk, 0, (cudaStream_t)stream>>>(native, ref_layout, n_vox);
  else
    grid_from_native_kernel<false><<<g, blk, 0, (cudaStream_t)stream>>>(native, ref_layout, n_vox);
  ENS_CHECK_CUDA();
  return ENS_OK;
}

extern "C" int ens_pack_decoder(int level, const float *const *tensors_host, int n_tensors, float *packed,
                                ens_stream_t stream) {
  if (!tensors_host || !packed) return ENS_EINVAL;
  if (level < 0 || level > 3) return ENS_EINVAL;
  const int need = (level == ENS_LEVEL_COARSE) ? 12 : 23;
  if (n_tensors != need) return ENS_ESHAPE;
  PtrTable t;
  for (int i = 0; i < 24; ++i) t.p[i] = (i < n_tensors) ? tensors_host[i] : nullptr;
  for (int i = 0; i < n_tensors; ++i) if (!t.p[i]) return ENS_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  auto nb = [](int total) { return (total + 255) / 256; };
  switch (level) {
    case ENS_LEVEL_COARSE: pack_coarse_kernel<<<nb(CoarsePack::total()), 256, 0, s>>>(t, packed); break;
    case ENS_LEVEL_MIDDLE:
      pack_mlp_all_kernel<32, 1><<<nb(packed_floats(level)), 256, 0, s>>>(t, packed);
      break;
    case ENS_LEVEL_FINE:
      pack_mlp_all_kernel<64, 1><<<nb(packed_floats(level)), 256, 0, s>>>(t, packed);
      break;
    default:
      pack_mlp_all_kernel<32, 4><<<nb(packed_floats(level)), 256, 0, s>>>(t, packed);
      break;
  }
  ENS_CHECK_CUDA();
  return ENS_OK;
}

extern "C" int ens_depth_max(const float *gt_depth, int64_t n, double *out, ens_stream_t stream) {
  if (!gt_depth || !out || n <= 0) return ENS_EINVAL;
  depth_max_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(gt_depth, n, out);
  ENS_CHECK_CUDA();
  return ENS_OK;
}

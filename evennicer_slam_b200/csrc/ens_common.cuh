// Shared device/host definitions for the fused renderer (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ens_render.h"

// Launch check: cudaGetLastError CLEARS a non-sticky error (bad launch configuration, too much shared memory), so a failed
// launch does not poison every later call of this library or of PyTorch; the message is kept for ens_last_error().
namespace ens {
void note_cuda_error(cudaError_t e, const char *file, int line);
int sm_count();   // SMs of the current device (cached per device; 148 on a B200), ens_capi.cu
}
#define ENS_CHECK_CUDA()                                   \
  do {                                                     \
    cudaError_t e__ = cudaGetLastError();                  \
    if (e__ != cudaSuccess) { ens::note_cuda_error(e__, __FILE__, __LINE__); return ENS_ECUDA; } \
  } while (0)

#define ENS_CUDA_CALL(x)                                  \
  do {                                                     \
    cudaError_t e__ = (x);                                 \
    if (e__ != cudaSuccess) { (void)cudaGetLastError(); ens::note_cuda_error(e__, __FILE__, __LINE__); return ENS_ECUDA; } \
  } while (0)

namespace ens {

constexpr int C = 32;      // feature channels per grid level
constexpr int HID = 32;    // hidden width of every decoder
constexpr int EMB = 93;    // Fourier features
constexpr int EMBP = 96;   // EMB padded to a multiple of 4 for the packed _B rows

// ---------------------------------------------------------------------------------------------
// Packed decoder blob (what the kernels stage into shared memory).  All matrices are stored
// TRANSPOSED, [in][out=32], so the 32 output weights of one input feature are one 128-byte row.
//   MLP (middle / fine / color), CD = c_dim (32 | 64):
//     B      [3][96]                      (Fourier matrix, k padded with zeros)
//     layer i=0..4:  W_iT [K_i][32], b_i [32], Wc_iT [CD][32], bc_i [32]     K = {93,32,32,125,32}
//     out:   WoT [32][4], bo [4]          (1-output decoders use column 0; the rest is zero)
//   MLP_no_xyz (coarse):
//     layer i=0..4:  W_iT [K_i][32], b_i [32]                               K = {32,32,32,64,32}
//     out:   WoT [32][4], bo [4]
// Layer 3's input is the skip concat: MLP rows 0..92 = embedding, 93..124 = h2 (decoder.py:198-199);
// MLP_no_xyz rows 0..31 = c, 32..63 = h2 (decoder.py:271-272).
// ---------------------------------------------------------------------------------------------
template <int CD>
struct MlpPack {
  __host__ __device__ static constexpr int K(int i) { return i == 0 ? 93 : (i == 3 ? 125 : 32); }
  __host__ __device__ static constexpr int layer_floats(int i) { return K(i) * 32 + 32 + CD * 32 + 32; }
  __host__ __device__ static constexpr int off_B() { return 0; }
  __host__ __device__ static constexpr int off_layer(int i) {
    int o = 3 * EMBP;
    for (int t = 0; t < i; ++t) o += layer_floats(t);
    return o;
  }
  __host__ __device__ static constexpr int off_W(int i) { return off_layer(i); }
  __host__ __device__ static constexpr int off_b(int i) { return off_layer(i) + K(i) * 32; }
  __host__ __device__ static constexpr int off_Wc(int i) { return off_b(i) + 32; }
  __host__ __device__ static constexpr int off_bc(int i) { return off_Wc(i) + CD * 32; }
  __host__ __device__ static constexpr int off_Wo() { return off_layer(5); }
  __host__ __device__ static constexpr int off_bo() { return off_Wo() + 128; }
  __host__ __device__ static constexpr int total() { return off_bo() + 4; }
};

// ---------------------------------------------------------------------------------------------
// Packed decoder blob, "mma" layout (tensor-core variant).  Matrices are stored as PyTorch keeps them,
// [out=32][in] with the input index contiguous, DENSE (no padding) and XOR-swizzled so that the
// m16n8k8 B-fragment loads (one 8-byte load of W[n][8kt+2t .. +1] per lane) are bank-conflict free:
//     element (n, k) of a [32][W] matrix lives at  n*W + (k ^ ((n & 3) << 3)).
//   B      [3][96]
//   W0     [32][96]   pts_linears.0.weight, k padded 93 -> 96 with zeros
//   W3e    [32][96]   pts_linears.3.weight[:, 0:93]   (embedding half of the skip layer)
//   block i=0..4:  Wh_i [32][32] (i=1,2,4: pts_linears.i.weight; i=3: pts_linears.3.weight[:, 93:125]; i=0: zeros)
//                  b_i  [32]     pts_linears.i.bias
//                  Wc_i [32][CD] fc_c.i.weight
//                  bc_i [32]     fc_c.i.bias
//   Wo     [4][32] output_linear.weight (rows >= NO zero), bo [4]
// ---------------------------------------------------------------------------------------------
template <int CD>
struct MlpPackV2 {
  __host__ __device__ static constexpr int off_B() { return 0; }
  __host__ __device__ static constexpr int off_W0() { return 3 * EMBP; }
  __host__ __device__ static constexpr int off_W3e() { return off_W0() + 32 * EMBP; }
  __host__ __device__ static constexpr int block_floats() { return 32 * 32 + 32 + 32 * CD + 32; }
  __host__ __device__ static constexpr int off_L(int i) { return off_W3e() + 32 * EMBP + i * block_floats(); }
  __host__ __device__ static constexpr int in_Wh() { return 0; }
  __host__ __device__ static constexpr int in_b() { return 32 * 32; }
  __host__ __device__ static constexpr int in_Wc() { return 32 * 32 + 32; }
  __host__ __device__ static constexpr int in_bc() { return 32 * 32 + 32 + 32 * CD; }
  __host__ __device__ static constexpr int off_Wo() { return off_L(5); }
  __host__ __device__ static constexpr int off_bo() { return off_Wo() + 128; }
  __host__ __device__ static constexpr int total() { return off_bo() + 4; }
};
__host__ __device__ constexpr int swz(int row, int k) { return k ^ ((row & 3) << 3); }

// ---------------------------------------------------------------------------------------------
// Packed decoder blob, "mma backward" layout: the TRANSPOSED matrices the data-gradient GEMMs read
// (g_x = g_u W), each stored [in][out=32] with the same XOR swizzle, so their B fragments are again one
// conflict-free 8-byte shared load per lane.
//   B      [3][96]
//   Wo     [4][32]
//   W0T    [96][32]   W0T[k][n]  = pts_linears.0.weight[n][k]       (rows k >= 93 zero)
//   W3eT   [96][32]   W3eT[k][n] = pts_linears.3.weight[n][k], k < 93
//   block i=0..4:  WhT_i [32][32]  WhT[k][n] = hidden part of pts_linears.i.weight[n][k]  (i = 0: zeros)
//                  WcT_i [32][32]  WcT[c][n] = fc_c.i.weight[n][c], c < 32   (only the first 32 feature
//                                  channels carry gradient: the fine decoder's middle half is no_grad)
// ---------------------------------------------------------------------------------------------
struct MlpPackV2B {
  __host__ __device__ static constexpr int off_B() { return 0; }
  __host__ __device__ static constexpr int off_Wo() { return 3 * EMBP; }
  __host__ __device__ static constexpr int off_W0T() { return off_Wo() + 128; }
  __host__ __device__ static constexpr int off_W3eT() { return off_W0T() + EMBP * 32; }
  __host__ __device__ static constexpr int off_L(int i) { return off_W3eT() + EMBP * 32 + i * 2048; }
  __host__ __device__ static constexpr int in_WhT() { return 0; }
  __host__ __device__ static constexpr int in_WcT() { return 1024; }
  __host__ __device__ static constexpr int total() { return off_L(5); }
};

struct CoarsePack {
  __host__ __device__ static constexpr int K(int i) { return i == 3 ? 64 : 32; }
  __host__ __device__ static constexpr int off_W(int i) {
    int o = 0;
    for (int t = 0; t < i; ++t) o += K(t) * 32 + 32;
    return o;
  }
  __host__ __device__ static constexpr int off_b(int i) { return off_W(i) + K(i) * 32; }
  __host__ __device__ static constexpr int off_Wo() { return off_W(5); }
  __host__ __device__ static constexpr int off_bo() { return off_Wo() + 128; }
  __host__ __device__ static constexpr int total() { return off_bo() + 4; }
};

// ---------------------------------------------------------------------------------------------
// Flat gradient buffer of one decoder = its tensors back to back in state_dict order, reference
// shapes (decoder.py:118-150):  fc_c.i.{weight [32][CD], bias [32]} i=0..4, embedder._B [3][93],
// pts_linears.i.{weight [32][K_i], bias [32]} i=0..4, output_linear.{weight [NO][32], bias [NO]}.
// ---------------------------------------------------------------------------------------------
template <int CD, int NO>
struct MlpGrad {
  __host__ __device__ static constexpr int K(int i) { return MlpPack<CD>::K(i); }
  __host__ __device__ static constexpr int off_Wc(int i) { return i * (32 * CD + 32); }
  __host__ __device__ static constexpr int off_bc(int i) { return off_Wc(i) + 32 * CD; }
  __host__ __device__ static constexpr int off_B() { return 5 * (32 * CD + 32); }
  __host__ __device__ static constexpr int off_W(int i) {
    int o = off_B() + 3 * EMB;
    for (int t = 0; t < i; ++t) o += 32 * K(t) + 32;
    return o;
  }
  __host__ __device__ static constexpr int off_b(int i) { return off_W(i) + 32 * K(i); }
  __host__ __device__ static constexpr int off_Wo() { return off_W(5); }
  __host__ __device__ static constexpr int off_bo() { return off_Wo() + NO * 32; }
  __host__ __device__ static constexpr int total() { return off_bo() + NO; }
};

struct CoarseGrad {
  __host__ __device__ static constexpr int K(int i) { return CoarsePack::K(i); }
  __host__ __device__ static constexpr int off_W(int i) {
    int o = 0;
    for (int t = 0; t < i; ++t) o += 32 * K(t) + 32;
    return o;
  }
  __host__ __device__ static constexpr int off_b(int i) { return off_W(i) + 32 * K(i); }
  __host__ __device__ static constexpr int off_Wo() { return off_W(5); }
  __host__ __device__ static constexpr int off_bo() { return off_Wo() + 32; }
  __host__ __device__ static constexpr int total() { return off_bo() + 1; }
};

// ---------------------------------------------------------------------------------------------
// Packed decoder blob, "tc" layout (tcgen05 / UMMA forward).  Every [32][K] matrix is stored in the canonical
// K-major, no-swizzle UMMA shared-memory layout (8-row x 16-byte core matrices):
//     element (n, k) at  (n/8)*(K/4)*32 + (k/4)*32 + (n%8)*4 + (k%4)   floats      (SBO = (K/4)*128 B, LBO = 128 B)
// twice: the value itself (the tensor core reads its top 19 bits = the TF32 "hi" part) and, TOT floats further,
// the remainder  lo = w - tf32_trunc(w)  for the 3xTF32 passes.
// The feature injection of block i is folded into block i+1 (h_i = relu(u_i) + Wc_i c + bc_i enters only through
// W_{i+1} h_i), so no separate feature accumulator is needed in TMEM:
//     u_{i+1} = W_{i+1} relu(u_i) + M_i c + b'_{i+1},   M_i = W_{i+1}[hidden] Wc_i,   b'_{i+1} = W_{i+1}[hidden] bc_i + b_{i+1}
//     out     = Wo relu(u_4) + Mo c + bo',              Mo  = Wo Wc_4,                bo' = Wo bc_4 + bo
//   matrices, in order: W0 [32][96], W3e [32][96], Wh_1, Wh_2, Wh_3 (hidden part), Wh_4 [32][32], M_0..M_3 [32][CD]
//   then: B [3][96], b'_i [5][32], Wo [4][32] (rows >= NO zero), Mo [4][CD], bo' [4]
// ---------------------------------------------------------------------------------------------
template <int CD>
struct MlpPackTC {
  __host__ __device__ static constexpr int off_W0() { return 0; }
  __host__ __device__ static constexpr int off_W3e() { return 32 * EMBP; }
  __host__ __device__ static constexpr int off_Wh(int i) { return 2 * 32 * EMBP + (i - 1) * 1024; }   // i = 1..4
  __host__ __device__ static constexpr int off_M(int i) { return 2 * 32 * EMBP + 4 * 1024 + i * 32 * CD; }   // i = 0..3
  __host__ __device__ static constexpr int TOT() { return 2 * 32 * EMBP + 4 * 1024 + 4 * 32 * CD; }
  __host__ __device__ static constexpr int off_B() { return 2 * TOT(); }
  __host__ __device__ static constexpr int off_b(int i) { return off_B() + 3 * EMBP + 32 * i; }
  __host__ __device__ static constexpr int off_Wo() { return off_B() + 3 * EMBP + 160; }
  __host__ __device__ static constexpr int off_Mo() { return off_Wo() + 128; }
  __host__ __device__ static constexpr int off_bo() { return off_Mo() + 4 * CD; }
  __host__ __device__ static constexpr int total() { return off_bo() + 4; }
};
__host__ __device__ constexpr int canon_off(int n, int k, int K) { return (n / 8) * (K / 4) * 32 + (k / 4) * 32 + (n % 8) * 4 + (k % 4); }

// ---------------------------------------------------------------------------------------------
// Packed decoder blob, "tc backward" layout (tcgen05 data-gradient GEMMs  g_x = g_u W).  The TRANSPOSED matrices in the
// same canonical K-major UMMA layout as MlpPackTC (K = 32 = the output units of the block), value then TF32 remainder
// (TOT floats further).  The feature gradient uses the folded matrices of the forward:  g_c = sum_i M_i^T g_u_{i+1} + Mo^T g_out
// (M_i = W_{i+1}[hidden] Wc_i), so the running gradient needs no un-masked g_h operand.  Only the first 32 feature
// channels carry gradient (the fine decoder's middle half is no_grad, decoder.py:184-186), so the blob has one size.
//   WhT_i [32 in][32 out], i = 1..4:   WhT_i[k][n] = Wh_i[n][k]
//   MT_i  [32 ch][32 out], i = 0..3:   MT_i[c][n]  = M_i[n][c]
//   W0T   [96][32], W3eT [96][32]:      W0T[k][n]  = pts_linears.0.weight[n][k]  (rows k >= 93 zero)
//   then: B [3][96], Wo [4][32] (rows >= NO zero), MoF [4][32] = Mo[:, :32]
// ---------------------------------------------------------------------------------------------
struct MlpPackTCB {
  // the matrices one block's data-gradient GEMM reads are CONTIGUOUS, so  [g_h | g_c | g_e] (adjacent TMEM columns) is ONE
  // accumulating GEMM per block:   L4 [WhT_4; MT_3]   L3 [WhT_3; MT_2; W3eT]   L2 [WhT_2; MT_1]   L1 [WhT_1; MT_0]   L0 [W0T]
  __host__ __device__ static constexpr int off_G(int i) { return i == 4 ? 0 : (i == 3 ? 2048 : (i == 2 ? 7168 : (i == 1 ? 9216 : 11264))); }
  __host__ __device__ static constexpr int off_WhT(int i) { return off_G(i); }                   // i = 1..4
  __host__ __device__ static constexpr int off_MT(int i) { return off_G(i + 1) + 1024; }         // i = 0..3
  __host__ __device__ static constexpr int off_W3eT() { return off_G(3) + 2048; }
  __host__ __device__ static constexpr int off_W0T() { return off_G(0); }
  __host__ __device__ static constexpr int TOT() { return 8192 + 2 * 32 * EMBP; }
  __host__ __device__ static constexpr int off_B() { return 2 * TOT(); }
  __host__ __device__ static constexpr int off_Wo() { return off_B() + 3 * EMBP; }
  __host__ __device__ static constexpr int off_MoF() { return off_Wo() + 128; }
  __host__ __device__ static constexpr int total() { return off_MoF() + 128; }
};

// A packed blob holds the fma layout, the mma forward layout and the mma backward layout (coarse: fma only).
__host__ __device__ inline int packed_floats(int level) {
  switch (level) {
    case ENS_LEVEL_COARSE: return CoarsePack::total();
    case ENS_LEVEL_FINE: return MlpPack<64>::total() + MlpPackV2<64>::total() + MlpPackV2B::total() + MlpPackTC<64>::total() + MlpPackTCB::total();
    default: return MlpPack<32>::total() + MlpPackV2<32>::total() + MlpPackV2B::total() + MlpPackTC<32>::total() + MlpPackTCB::total();
  }
}
template <int CD> __host__ __device__ constexpr int off_v2() { return MlpPack<CD>::total(); }
template <int CD> __host__ __device__ constexpr int off_v2b() { return MlpPack<CD>::total() + MlpPackV2<CD>::total(); }
template <int CD> __host__ __device__ constexpr int off_tc() { return off_v2b<CD>() + MlpPackV2B::total(); }
template <int CD> __host__ __device__ constexpr int off_tcb() { return off_tc<CD>() + MlpPackTC<CD>::total(); }
__host__ __device__ inline int grad_floats(int level) {
  switch (level) {
    case ENS_LEVEL_COARSE: return CoarseGrad::total();
    case ENS_LEVEL_MIDDLE: return MlpGrad<32, 1>::total();
    case ENS_LEVEL_FINE: return MlpGrad<64, 1>::total();
    default: return MlpGrad<32, 4>::total();
  }
}

// Scene as the kernels see it (passed by value as a kernel parameter).
struct DevScene {
  const float *grid[4];
  int dims[4][3];        // Z,Y,X
  double lo[3], hi[3];   // slam.bound
  double clo[3], chi[3]; // coarse bound
  const float *w[4];
};

inline DevScene make_dev_scene(const EnsScene *s) {
  DevScene d;
  for (int l = 0; l < 4; ++l) {
    d.grid[l] = s->grid[l];
    d.w[l] = s->weights[l];
    for (int k = 0; k < 3; ++k) d.dims[l][k] = s->dims[l][k];
  }
  for (int k = 0; k < 3; ++k) {
    d.lo[k] = s->bound[k][0];
    d.hi[k] = s->bound[k][1];
    d.clo[k] = s->coarse_bound[k][0];
    d.chi[k] = s->coarse_bound[k][1];
  }
  return d;
}

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double tmax_nan(double a, double b) {  // torch.max: NaN-propagating
  return (a != a) ? a : ((b != b) ? b : (a > b ? a : b));
}
__device__ __forceinline__ double tmin_nan(double a, double b) {
  return (a != a) ? a : ((b != b) ? b : (a < b ? a : b));
}

__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d) {
  // vectorised no-return global reduction (sm_90+): one 16-byte L2 atomic instead of four
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

}  // namespace ens

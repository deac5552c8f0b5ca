// Tensor-core (mma.sync m16n8k8 TF32, 3xTF32 split) building blocks shared by the forward and backward
// render kernels.  See ens_render_mma.cu for the mapping of points / features to fragments.
#pragma once
#include "ens_device.cuh"

namespace ens {

// ---------------------------------------------------------------------------------------------
// tensor-core primitives
// ---------------------------------------------------------------------------------------------
// hi = x with the low 13 mantissa bits cleared (what the tensor core reads anyway), lo = x - hi (exact).
// Bit masking keeps the split on the ALU pipe: cvt.rna.tf32 issues on the quarter-rate XU pipe, which the
// first capture of this kernel showed saturated (profiles/r01_mma_fwd_*).  lo is passed as is; the tensor
// core truncates it to 11 bits, a 2^-21 relative effect on the product.
__device__ __forceinline__ void split_tf32(float x, uint32_t &hi, uint32_t &lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}

// sin / cos of a Fourier argument (|q| up to a few thousand rad: p.B with B ~ N(0, 25^2)).
// Cody-Waite reduction by 2*pi in three float32 pieces (k * C1 is exact for |k| < 2^16: C1 has 8 mantissa
// bits), then the SFU on r in [-pi, pi], where MUFU.SIN/COS err by < 2^-21 absolute -- an order of magnitude
// below the float32 rounding of q itself (ulp(300) = 3e-5).  libdevice sinf() costs ~40 instructions per call
// and was half of the kernel's issue slots.
__device__ __forceinline__ float reduce_2pi(float q) {
  // round-to-nearest by the 1.5 * 2^23 trick (|q / 2 pi| < 2^22): two FMA-pipe adds instead of FRND on the quarter-rate XU pipe,
  // which the sines already load
  const float k = __fsub_rn(__fadd_rn(q * 0.15915494309189535f, 12582912.f), 12582912.f);
  float r = fmaf(k, -6.28125f, q);
  r = fmaf(k, -1.9353071693331003e-3f, r);   // float32(2*pi - 6.28125)
  r = fmaf(k, -1.0253131677e-11f, r);        // what the float32 rounding of the previous constant left
  return r;
}
__device__ __forceinline__ float fast_sin(float q) { return __sinf(reduce_2pi(q)); }
__device__ __forceinline__ void fast_sincos(float q, float &s, float &c) {
  const float r = reduce_2pi(q);
  s = __sinf(r);
  c = __cosf(r);
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  // not volatile: the instruction has no side effect beyond d, so ptxas may interleave independent tiles
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// acc[m][nt] += A[m] (16x8, given as hi/lo) * W[8nt..8nt+7][k-tile kt]^T for all four n-tiles; 3xTF32.
// W: swizzled [32][WROW] in shared memory.  The three passes (lo.hi, hi.lo, hi.hi) are the OUTER loops, so
// the eight accumulator tiles give eight independent HMMAs between two dependent ones.
template <int WROW>
__device__ __forceinline__ void mma_ktile(float (&acc)[2][4][4], const uint32_t (&ah)[2][4], const uint32_t (&al)[2][4],
                                          const float *__restrict__ W, int kt, int g, int t) {
  uint32_t bh[4][2], bl[4][2];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const int n = 8 * nt + g;
    const float2 w = *reinterpret_cast<const float2 *>(W + n * WROW + ((8 * kt + 2 * t) ^ ((g & 3) << 3)));
    split_tf32(w.x, bh[nt][0], bl[nt][0]);
    split_tf32(w.y, bh[nt][1], bl[nt][1]);
  }
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int m = 0; m < 2; ++m) mma_tf32(acc[m][nt], al[m], bh[nt][0], bh[nt][1]);
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int m = 0; m < 2; ++m) mma_tf32(acc[m][nt], ah[m], bl[nt][0], bl[nt][1]);
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int m = 0; m < 2; ++m) mma_tf32(acc[m][nt], ah[m], bh[nt][0], bh[nt][1]);
}

// A fragments of k-tile kt from an activation held as accumulator fragments x[m][kt][0..3]
__device__ __forceinline__ void frag_from_regs(const float (&x)[4], uint32_t (&ah)[4], uint32_t (&al)[4]) {
  split_tf32(x[0], ah[0], al[0]);   // (row g,   feature 2t)
  split_tf32(x[2], ah[1], al[1]);   // (row g+8, feature 2t)
  split_tf32(x[1], ah[2], al[2]);   // (row g,   feature 2t+1)
  split_tf32(x[3], ah[3], al[3]);   // (row g+8, feature 2t+1)
}

// acc += x * Wh^T  (hidden layer, K = 32)
__device__ __forceinline__ void gemm_hidden(float (&acc)[2][4][4], const float (&x)[2][4][4],
                                            const float *__restrict__ Wh, int g, int t) {
#pragma unroll
  for (int kt = 0; kt < 4; ++kt) {
    uint32_t ah[2][4], al[2][4];
    frag_from_regs(x[0][kt], ah[0], al[0]);
    frag_from_regs(x[1][kt], ah[1], al[1]);
    mma_ktile<32>(acc, ah, al, Wh, kt, g, t);
  }
}

// acc += c * Wc^T with the features read from the warp's shared tile: rows of RS floats, columns C0..C0+CD-1
template <int CD, int RS>
__device__ __forceinline__ void gemm_features(float (&acc)[2][4][4], const float *__restrict__ crow, int c0,
                                              const float *__restrict__ Wc, int g, int t) {
#pragma unroll
  for (int kt = 0; kt < CD / 8; ++kt) {
    uint32_t ah[2][4], al[2][4];
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      const int r0 = 16 * m + g, r1 = r0 + 8;                     // (r & 3) == (g & 3) for both
      const int col = (c0 + 8 * kt + 2 * t) ^ ((g & 3) << 3);
      const float2 v0 = *reinterpret_cast<const float2 *>(crow + r0 * RS + col);
      const float2 v1 = *reinterpret_cast<const float2 *>(crow + r1 * RS + col);
      split_tf32(v0.x, ah[m][0], al[m][0]);
      split_tf32(v1.x, ah[m][1], al[m][1]);
      split_tf32(v0.y, ah[m][2], al[m][2]);
      split_tf32(v1.y, ah[m][3], al[m][3]);
    }
    mma_ktile<CD>(acc, ah, al, Wc, kt, g, t);
  }
}

__device__ __forceinline__ void set_bias(float (&acc)[2][4][4], const float *__restrict__ b, int t) {
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const float2 v = *reinterpret_cast<const float2 *>(b + 8 * nt + 2 * t);
#pragma unroll
    for (int m = 0; m < 2; ++m) { acc[m][nt][0] = v.x; acc[m][nt][1] = v.y; acc[m][nt][2] = v.x; acc[m][nt][3] = v.y; }
  }
}
// acc = relu(acc) + bc
__device__ __forceinline__ void relu_add_bias(float (&acc)[2][4][4], const float *__restrict__ bc, int t) {
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const float2 v = *reinterpret_cast<const float2 *>(bc + 8 * nt + 2 * t);
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      acc[m][nt][0] = fmaxf(acc[m][nt][0], 0.f) + v.x; acc[m][nt][1] = fmaxf(acc[m][nt][1], 0.f) + v.y;
      acc[m][nt][2] = fmaxf(acc[m][nt][2], 0.f) + v.x; acc[m][nt][3] = fmaxf(acc[m][nt][3], 0.f) + v.y;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// small helpers on accumulator-layout tiles
// ---------------------------------------------------------------------------------------------
#define ENS_FOR_TILE(m, nt, e)            \
  _Pragma("unroll") for (int m = 0; m < 2; ++m) \
  _Pragma("unroll") for (int nt = 0; nt < 4; ++nt) \
  _Pragma("unroll") for (int e = 0; e < 4; ++e)

__device__ __forceinline__ void zero_tile(float (&a)[2][4][4]) {
  ENS_FOR_TILE(m, nt, e) a[m][nt][e] = 0.f;
}

// acc = relu(acc) + bc ; returns the mask of acc > 0 (bit (m*4+nt)*4+e)
__device__ __forceinline__ uint32_t relu_add_bias_mask(float (&acc)[2][4][4], const float *__restrict__ bc, int t) {
  uint32_t mk = 0;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const float2 v = *reinterpret_cast<const float2 *>(bc + 8 * nt + 2 * t);
#pragma unroll
    for (int m = 0; m < 2; ++m) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float a = acc[m][nt][e];
        if (a > 0.f) mk |= 1u << ((m * 4 + nt) * 4 + e);
        acc[m][nt][e] = fmaxf(a, 0.f) + ((e & 1) ? v.y : v.x);
      }
    }
  }
  return mk;
}

// store an accumulator-layout tile as a swizzled [32][LD] row-major tile (shared or global):
// element (row r, col c) at base + r*LD + ((c0 + c) ^ ((r & 3) << 3))
template <int LD>
__device__ __forceinline__ void store_tile(float *__restrict__ base, int c0, const float (&x)[2][4][4], int g, int t) {
  const int sw = (g & 3) << 3;
#pragma unroll
  for (int m = 0; m < 2; ++m) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int col = (c0 + 8 * nt + 2 * t) ^ sw;
      *reinterpret_cast<float2 *>(base + (16 * m + g) * LD + col) = make_float2(x[m][nt][0], x[m][nt][1]);
      *reinterpret_cast<float2 *>(base + (16 * m + g + 8) * LD + col) = make_float2(x[m][nt][2], x[m][nt][3]);
    }
  }
}

template <int LD>
__device__ __forceinline__ void load_tile(const float *__restrict__ base, int c0, float (&x)[2][4][4], int g, int t) {
  const int sw = (g & 3) << 3;
#pragma unroll
  for (int m = 0; m < 2; ++m) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int col = (c0 + 8 * nt + 2 * t) ^ sw;
      const float2 a = *reinterpret_cast<const float2 *>(base + (16 * m + g) * LD + col);
      const float2 b = *reinterpret_cast<const float2 *>(base + (16 * m + g + 8) * LD + col);
      x[m][nt][0] = a.x; x[m][nt][1] = a.y; x[m][nt][2] = b.x; x[m][nt][3] = b.y;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// One Fourier-feature decoder for the warp's 32 points (decoder.py:177-203).
// sw: MlpPackV2<CD> blob in shared memory.  crow: the warp's feature tile.  (px,py,pz): the OWNER lane's
// point (p.float()).  Returns the decoder outputs of the owner lane's point in out[NO].
// ---------------------------------------------------------------------------------------------
// SAVE: 0 = nothing; 1 = the five relu masks (one word per lane and block) to msave[i*32]; 2 = masks and the
// activations h_0..h_4 as swizzled 32x32 tiles to hsave[i*1024..] -- what the backward kernel needs to run without
// recomputing this forward (ens_render_mma_bwd.cu).
template <int CD, int RS, int NO, int SAVE = 0>
__device__ __forceinline__ void mlp_mma(const float *__restrict__ sw, const float *__restrict__ crow, int c0,
                                        float px, float py, float pz, float (&out)[NO],
                                        uint32_t *__restrict__ msave = nullptr, float *__restrict__ hsave = nullptr) {
  using P = MlpPackV2<CD>;
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  float rx[4], ry[4], rz[4];                                     // my four rows: points g, g+8, g+16, g+24
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    rx[j] = __shfl_sync(0xffffffffu, px, g + 8 * j);
    ry[j] = __shfl_sync(0xffffffffu, py, g + 8 * j);
    rz[j] = __shfl_sync(0xffffffffu, pz, g + 8 * j);
  }
  float acc[2][4][4], acc3[2][4][4];
  set_bias(acc, sw + P::off_L(0) + P::in_b(), t);
  set_bias(acc3, sw + P::off_L(3) + P::in_b(), t);
  // ---- Fourier embedding, consumed k-tile by k-tile by block 0 and by the skip block's embedding half ----
#pragma unroll 1
  for (int kt = 0; kt < EMBP / 8; ++kt) {
    const float *B = sw + P::off_B() + 8 * kt + 2 * t;
    const float2 b0 = *reinterpret_cast<const float2 *>(B);
    const float2 b1 = *reinterpret_cast<const float2 *>(B + EMBP);
    const float2 b2 = *reinterpret_cast<const float2 *>(B + 2 * EMBP);
    uint32_t ah[2][4], al[2][4];
#pragma unroll
    for (int m = 0; m < 2; ++m) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {                              // h = 0: row g, h = 1: row g+8
        const int j = 2 * m + h;
        const float q0 = fmaf(rz[j], b2.x, fmaf(ry[j], b1.x, rx[j] * b0.x));
        const float q1 = fmaf(rz[j], b2.y, fmaf(ry[j], b1.y, rx[j] * b0.y));
        split_tf32(fast_sin(q0), ah[m][h], al[m][h]);
        split_tf32(fast_sin(q1), ah[m][2 + h], al[m][2 + h]);
      }
    }
    mma_ktile<EMBP>(acc, ah, al, sw + P::off_W0(), kt, g, t);
    mma_ktile<EMBP>(acc3, ah, al, sw + P::off_W3e(), kt, g, t);
  }
  // ---- blocks 0..4: h_i = relu(u_i) + fc_c[i](c); one copy of the GEMM code, looped (I-cache) ----
  float x[2][4][4];
#pragma unroll 1
  for (int i = 0; i < 5; ++i) {
    const float *L = sw + P::off_L(0) + i * P::block_floats();
    if (i > 0) {
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            x[m][nt][e] = acc[m][nt][e];
            acc[m][nt][e] = acc3[m][nt][e];          // only meaningful for i == 3
          }
      if (i != 3) set_bias(acc, L + P::in_b(), t);
      gemm_hidden(acc, x, L + P::in_Wh(), g, t);
    }
    if (SAVE) msave[i * 32] = relu_add_bias_mask(acc, L + P::in_bc(), t);
    else relu_add_bias(acc, L + P::in_bc(), t);
    gemm_features<CD, RS>(acc, crow, c0, L + P::in_Wc(), g, t);
    if (SAVE == 2) store_tile<32>(hsave + i * 1024, 0, acc, g, t);
  }
  // ---- output layer on the FMA pipe: per-lane partial dot over its 8 features, quad reduce ----
  float part[4][NO];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int o = 0; o < NO; ++o) part[j][o] = 0.f;
#pragma unroll
  for (int o = 0; o < NO; ++o) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const float2 w = *reinterpret_cast<const float2 *>(sw + P::off_Wo() + o * 32 + 8 * nt + 2 * t);
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        part[2 * m][o] = fmaf(acc[m][nt][1], w.y, fmaf(acc[m][nt][0], w.x, part[2 * m][o]));
        part[2 * m + 1][o] = fmaf(acc[m][nt][3], w.y, fmaf(acc[m][nt][2], w.x, part[2 * m + 1][o]));
      }
    }
  }
#pragma unroll
  for (int o = 0; o < NO; ++o) {
    float mine = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = part[j][o];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      // row j of quad q is point q + 8j: deliver to its owner lane
      const float got = __shfl_sync(0xffffffffu, v, (lane & 7) * 4);
      if ((lane >> 3) == j) mine = got;
    }
    out[o] = mine + sw[P::off_bo() + o];
  }
}

// ---------------------------------------------------------------------------------------------
// warp-cooperative trilinear gather: lane L owns point L (vox); 8 lanes fetch one point's 128-byte voxel
// lines.  Writes the 32 channels of the warp's 32 points to crow[pt][c0 .. c0+31] (swizzled rows of RS).
// ---------------------------------------------------------------------------------------------
template <int RS>
__device__ __forceinline__ void gather_warp(const float *__restrict__ grid, const int dims[3], const Vox &v,
                                            float *__restrict__ crow, int c0) {
  const int lane = threadIdx.x & 31;
  const int X = dims[2], Y = dims[1], Z = dims[0];
  // owner-side packing: base element offset of corner 0, per-axis steps (0 if the +1 neighbour is out of range)
  const int base = ((v.z0 * Y + v.y0) * X + v.x0) * C;
  const bool okx = v.x0 + 1 < X, oky = v.y0 + 1 < Y, okz = v.z0 + 1 < Z;
  const float wx1 = okx ? v.fx : 0.f, wy1 = oky ? v.fy : 0.f, wz1 = okz ? v.fz : 0.f;
  const int cq = lane & 7, pp = lane >> 3;
  const int sx = C, sy = X * C, sz = X * Y * C;
#pragma unroll 2
  for (int grp = 0; grp < 8; ++grp) {
    const int src = 4 * grp + pp;
    const int b = __shfl_sync(0xffffffffu, base, src);
    const float fx1 = __shfl_sync(0xffffffffu, wx1, src), fx0 = __shfl_sync(0xffffffffu, v.gx, src);
    const float fy1 = __shfl_sync(0xffffffffu, wy1, src), fy0 = __shfl_sync(0xffffffffu, v.gy, src);
    const float fz1 = __shfl_sync(0xffffffffu, wz1, src), fz0 = __shfl_sync(0xffffffffu, v.gz, src);
    const unsigned okb = __shfl_sync(0xffffffffu, (unsigned)okx | ((unsigned)oky << 1) | ((unsigned)okz << 2), src);
    const int ox = (okb & 1u) ? sx : 0, oy = (okb & 2u) ? sy : 0, oz = (okb & 4u) ? sz : 0;
    const float *p = grid + b + 4 * cq;
    float4 a[8];
#pragma unroll
    for (int c = 0; c < 8; ++c)
      a[c] = __ldg(reinterpret_cast<const float4 *>(p + ((c & 1) ? ox : 0) + ((c & 2) ? oy : 0) + ((c & 4) ? oz : 0)));
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      // ATen order of the product: (wx * wy) * wz
      const float w = __fmul_rn(__fmul_rn((c & 1) ? fx1 : fx0, (c & 2) ? fy1 : fy0), (c & 4) ? fz1 : fz0);
      r.x = fmaf(a[c].x, w, r.x); r.y = fmaf(a[c].y, w, r.y); r.z = fmaf(a[c].z, w, r.z); r.w = fmaf(a[c].w, w, r.w);
    }
    *reinterpret_cast<float4 *>(crow + src * RS + ((c0 + 4 * cq) ^ ((src & 3) << 3))) = r;
  }
  __syncwarp();
}

// one group of four points (grp = 0..7) of gather_warp: callers that have latency to hide elsewhere (the tcgen05 decode walks
// its MMA round trips) run the eight groups of the NEXT tile one at a time in between
template <int RS>
__device__ __forceinline__ void gather_warp_step(const float *__restrict__ grid, const int dims[3], const Vox &v,
                                                 float *__restrict__ crow, int c0, int grp) {
  const int lane = threadIdx.x & 31;
  const int X = dims[2], Y = dims[1], Z = dims[0];
  // owner-side packing: base element offset of corner 0, per-axis steps (0 if the +1 neighbour is out of range)
  const int base = ((v.z0 * Y + v.y0) * X + v.x0) * C;
  const bool okx = v.x0 + 1 < X, oky = v.y0 + 1 < Y, okz = v.z0 + 1 < Z;
  const float wx1 = okx ? v.fx : 0.f, wy1 = oky ? v.fy : 0.f, wz1 = okz ? v.fz : 0.f;
  const int cq = lane & 7, pp = lane >> 3;
  const int sx = C, sy = X * C, sz = X * Y * C;
  const int src = 4 * grp + pp;
  const int b = __shfl_sync(0xffffffffu, base, src);
  const float fx1 = __shfl_sync(0xffffffffu, wx1, src), fx0 = __shfl_sync(0xffffffffu, v.gx, src);
  const float fy1 = __shfl_sync(0xffffffffu, wy1, src), fy0 = __shfl_sync(0xffffffffu, v.gy, src);
  const float fz1 = __shfl_sync(0xffffffffu, wz1, src), fz0 = __shfl_sync(0xffffffffu, v.gz, src);
  const unsigned okb = __shfl_sync(0xffffffffu, (unsigned)okx | ((unsigned)oky << 1) | ((unsigned)okz << 2), src);
  const int ox = (okb & 1u) ? sx : 0, oy = (okb & 2u) ? sy : 0, oz = (okb & 4u) ? sz : 0;
  const float *p = grid + b + 4 * cq;
  float4 a[8];
#pragma unroll
  for (int c = 0; c < 8; ++c)
    a[c] = __ldg(reinterpret_cast<const float4 *>(p + ((c & 1) ? ox : 0) + ((c & 2) ? oy : 0) + ((c & 4) ? oz : 0)));
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    // ATen order of the product: (wx * wy) * wz
    const float w = __fmul_rn(__fmul_rn((c & 1) ? fx1 : fx0, (c & 2) ? fy1 : fy0), (c & 4) ? fz1 : fz0);
    r.x = fmaf(a[c].x, w, r.x); r.y = fmaf(a[c].y, w, r.y); r.z = fmaf(a[c].z, w, r.z); r.w = fmaf(a[c].w, w, r.w);
  }
  *reinterpret_cast<float4 *>(crow + src * RS + ((c0 + 4 * cq) ^ ((src & 3) << 3))) = r;
}

// ---------------------------------------------------------------------------------------------
// trilinear backward for the warp's 32 points.  crow: the warp's feature-tile rows; columns 0..31 hold the
// feature gradient g_c (swizzled) -- written by store_tile just before.  8 lanes per point.
//   ggrid != null : scatter  w_corner * g_c  into the native-layout gradient grid (valid points only)
//   want_coord    : gpn = d L / d (normalised coords) of the owner lane's point
// ---------------------------------------------------------------------------------------------
template <int RS>
__device__ __forceinline__ void gather_bwd_warp(const float *__restrict__ grid, float *__restrict__ ggrid,
                                                const int dims[3], const Vox &v, bool valid,
                                                const float *__restrict__ crow, bool want_coord, float (&gpn)[3]) {
  const int lane = threadIdx.x & 31;
  const int X = dims[2], Y = dims[1], Z = dims[0];
  const int base = ((v.z0 * Y + v.y0) * X + v.x0) * C;
  const bool okx = v.x0 + 1 < X, oky = v.y0 + 1 < Y, okz = v.z0 + 1 < Z;
  const unsigned okbits = (unsigned)okx | ((unsigned)oky << 1) | ((unsigned)okz << 2) | ((unsigned)valid << 3);
  const int cq = lane & 7, pp = lane >> 3;
  const int sx = C, sy = X * C, sz = X * Y * C;
  float mx = 0.f, my = 0.f, mz = 0.f;
  // the eight corner lines of a point are re-read for the coordinate gradient: with few warps per SM (the tcgen05 kernels)
  // that L2 latency is exposed once per group, so the lines of the group after next are prefetched into L1 meanwhile
  auto prefetch_group = [&](int g2) {
    const int src2 = 4 * g2 + pp;
    const int b2 = __shfl_sync(0xffffffffu, base, src2);
    const unsigned ok2 = __shfl_sync(0xffffffffu, okbits, src2);
    if (cq < 8) {
      const int c = cq;                                              // lane cq of the point's eight fetches corner cq
      const int off = b2 + ((c & 1) && (ok2 & 1u) ? sx : 0) + ((c & 2) && (ok2 & 2u) ? sy : 0) + ((c & 4) && (ok2 & 4u) ? sz : 0);
      asm volatile("prefetch.global.L1 [%0];" ::"l"(grid + off));
    }
  };
  if (want_coord) { prefetch_group(0); prefetch_group(1); prefetch_group(2); prefetch_group(3); }
#pragma unroll 2
  for (int grp = 0; grp < 8; ++grp) {
    if (want_coord && grp + 4 < 8) prefetch_group(grp + 4);
    const int src = 4 * grp + pp;
    const int b = __shfl_sync(0xffffffffu, base, src);
    const float fx1 = __shfl_sync(0xffffffffu, v.fx, src), fx0 = __shfl_sync(0xffffffffu, v.gx, src);
    const float fy1 = __shfl_sync(0xffffffffu, v.fy, src), fy0 = __shfl_sync(0xffffffffu, v.gy, src);
    const float fz1 = __shfl_sync(0xffffffffu, v.fz, src), fz0 = __shfl_sync(0xffffffffu, v.gz, src);
    const unsigned okb = __shfl_sync(0xffffffffu, okbits, src);
    const int ox = (okb & 1u) ? sx : 0, oy = (okb & 2u) ? sy : 0, oz = (okb & 4u) ? sz : 0;
    const float4 gc = *reinterpret_cast<const float4 *>(crow + src * RS + ((4 * cq) ^ ((src & 3) << 3)));
    const int off0 = b + 4 * cq;
    float gix = 0.f, giy = 0.f, giz = 0.f;
    float4 a[8];
    if (want_coord) {       // all eight corner lines in flight before the first use (addresses are always valid)
#pragma unroll
      for (int c = 0; c < 8; ++c)
        a[c] = __ldg(reinterpret_cast<const float4 *>(grid + off0 + ((c & 1) ? ox : 0) + ((c & 2) ? oy : 0) + ((c & 4) ? oz : 0)));
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const bool inr = (!(c & 1) || (okb & 1u)) && (!(c & 2) || (okb & 2u)) && (!(c & 4) || (okb & 4u));
      const int off = off0 + ((c & 1) ? ox : 0) + ((c & 2) ? oy : 0) + ((c & 4) ? oz : 0);
      const float wx = (c & 1) ? fx1 : fx0, wy = (c & 2) ? fy1 : fy0, wz = (c & 4) ? fz1 : fz0;
      if (inr) {
        if (ggrid != nullptr && (okb & 8u)) {
          const float w = __fmul_rn(__fmul_rn(wx, wy), wz);
          if (w != 0.f) red_add_v4(ggrid + off, w * gc.x, w * gc.y, w * gc.z, w * gc.w);
        }
        if (want_coord) {
          const float dot = fmaf(a[c].w, gc.w, fmaf(a[c].z, gc.z, fmaf(a[c].y, gc.y, a[c].x * gc.x)));
          gix += ((c & 1) ? 1.f : -1.f) * wy * wz * dot;
          giy += ((c & 2) ? 1.f : -1.f) * wx * wz * dot;
          giz += ((c & 4) ? 1.f : -1.f) * wx * wy * dot;
        }
      }
    }
    if (want_coord) {
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        gix += __shfl_xor_sync(0xffffffffu, gix, o);
        giy += __shfl_xor_sync(0xffffffffu, giy, o);
        giz += __shfl_xor_sync(0xffffffffu, giz, o);
      }
      // point 4*grp + pp was handled by lanes 8pp..8pp+7: deliver to its owner lane
      const float ax = __shfl_sync(0xffffffffu, gix, 8 * (lane & 3));
      const float ay = __shfl_sync(0xffffffffu, giy, 8 * (lane & 3));
      const float az = __shfl_sync(0xffffffffu, giz, 8 * (lane & 3));
      if ((lane >> 2) == grp) { mx = ax; my = ay; mz = az; }
    }
  }
  gpn[0] = mx * v.sx; gpn[1] = my * v.sy; gpn[2] = mz * v.sz;
}

// Copy a packed weight blob global -> shared with cp.async (LDGSTS, 16 bytes per lane, no register round trip):
// all copies of the CTA are in flight at once; the caller's __syncthreads() (after stage_blob_wait) publishes them.
__device__ __forceinline__ void stage_blob(float *__restrict__ sw, const float *__restrict__ gw, int nfloats) {
  const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(sw);
  for (int i = threadIdx.x; i < nfloats / 4; i += blockDim.x)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s0 + i * 16), "l"(gw + i * 4) : "memory");
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void stage_blob_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int STAGE> struct MmaStage;
template <> struct MmaStage<ENS_STAGE_MIDDLE> { static constexpr int RS = 32; static constexpr int WMAX = MlpPackV2<32>::total(); };
template <> struct MmaStage<ENS_STAGE_FINE> { static constexpr int RS = 64; static constexpr int WMAX = MlpPackV2<64>::total(); };
template <> struct MmaStage<ENS_STAGE_COLOR> { static constexpr int RS = 64; static constexpr int WMAX = MlpPackV2<64>::total(); };

}  // namespace ens

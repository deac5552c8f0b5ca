// Fused render kernels, variant "mma": the 32-wide decoder layers run on the tensor pipe.
//
// ncu on the fma variant (profiles/r01_fma_*) shows the kernels contraction-bound (FMA pipe the busiest
// unit at 19-29 %, DRAM < 0.2 %, one 6-warp CTA per SM), so -- as BASELINE.json's north_star prescribes --
// the layer GEMMs move to mma.sync.m16n8k8 TF32 with the 3xTF32 error-compensated split
// (a = a_hi + a_lo, b = b_hi + b_lo;  a.b ~= a_lo.b_hi + a_hi.b_lo + a_hi.b_hi), which keeps float32-level
// accuracy (outputs 1e-4, gradients 1e-3 are the contract; plain TF32 would not meet it).
//
// Mapping.  A warp owns 32 consecutive sample points = two m16 row tiles.  Lane L is the "owner" of point
// L (sample placement, mask, trilinear coordinates, final raw value); for the GEMMs the same lane, as
// (g = L/4, t = L%4), holds rows {g, g+8, g+16, g+24} of the accumulator fragments.  Activations never
// leave registers between layers: an accumulator fragment (row g, cols 8n+2t, 8n+2t+1) is reused directly
// as the A fragment of k-tile n of the next layer, with the k order inside a tile permuted to
// (2t, 2t+1); the B fragments are loaded with the same permutation, which in the [out][in] weight layout
// is one 8-byte shared load per lane (XOR-swizzled, conflict-free; see MlpPackV2).
// Grid features are gathered by 8 lanes per point (one 128-byte voxel line per 8 lanes, LDG.128) into a
// warp-private swizzled [32][CD] shared tile and read back as A fragments for the fc_c GEMMs.
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
  const float k = rintf(q * 0.15915494309189535f);
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
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// acc[m][nt] += A[m] (16x8, given as hi/lo) * W[8nt..8nt+7][k-tile kt]^T for all four n-tiles; 3xTF32.
// W: swizzled [32][WROW] in shared memory.
template <int WROW>
__device__ __forceinline__ void mma_ktile(float (&acc)[2][4][4], const uint32_t (&ah)[2][4], const uint32_t (&al)[2][4],
                                          const float *__restrict__ W, int kt, int g, int t) {
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const int n = 8 * nt + g;
    const float2 w = *reinterpret_cast<const float2 *>(W + n * WROW + ((8 * kt + 2 * t) ^ ((g & 3) << 3)));
    uint32_t bh0, bl0, bh1, bl1;
    split_tf32(w.x, bh0, bl0);
    split_tf32(w.y, bh1, bl1);
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      mma_tf32(acc[m][nt], al[m], bh0, bh1);
      mma_tf32(acc[m][nt], ah[m], bl0, bl1);
      mma_tf32(acc[m][nt], ah[m], bh0, bh1);
    }
  }
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
// One Fourier-feature decoder for the warp's 32 points (decoder.py:177-203).
// sw: MlpPackV2<CD> blob in shared memory.  crow: the warp's feature tile.  (px,py,pz): the OWNER lane's
// point (p.float()).  Returns the decoder outputs of the owner lane's point in out[NO].
// ---------------------------------------------------------------------------------------------
template <int CD, int RS, int NO>
__device__ __forceinline__ void mlp_mma(const float *__restrict__ sw, const float *__restrict__ crow, int c0,
                                        float px, float py, float pz, float (&out)[NO]) {
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
    relu_add_bias(acc, L + P::in_bc(), t);
    gemm_features<CD, RS>(acc, crow, c0, L + P::in_Wc(), g, t);
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

__device__ __forceinline__ void stage_blob(float *__restrict__ sw, const float *__restrict__ gw, int nfloats) {
  const float4 *src = reinterpret_cast<const float4 *>(gw);
  float4 *dst = reinterpret_cast<float4 *>(sw);
  for (int i = threadIdx.x; i < nfloats / 4; i += blockDim.x) dst[i] = __ldg(src + i);
}

template <int STAGE> struct MmaStage;
template <> struct MmaStage<ENS_STAGE_MIDDLE> { static constexpr int RS = 32; static constexpr int WMAX = MlpPackV2<32>::total(); };
template <> struct MmaStage<ENS_STAGE_FINE> { static constexpr int RS = 64; static constexpr int WMAX = MlpPackV2<64>::total(); };
template <> struct MmaStage<ENS_STAGE_COLOR> { static constexpr int RS = 64; static constexpr int WMAX = MlpPackV2<64>::total(); };

// NICE.forward for the owner lane's point (decoder.py:312-342).  CTA-collective.
// sw: weight region; sfeat: the CTA's [NT][RS] feature tile.
template <int STAGE>
__device__ __forceinline__ float4 decode_stage_mma(const DevScene &sc, float *__restrict__ sw, float *__restrict__ sfeat,
                                                   const float pn[3], const float p32[3]) {
  constexpr int RS = MmaStage<STAGE>::RS;
  float *crow = sfeat + (threadIdx.x >> 5) * 32 * RS;             // this warp's 32 rows
  float4 raw = make_float4(0.f, 0.f, 0.f, 0.f);
  {   // middle decoder; in the fine/colour stages its features also form the fine decoder's concat half
    constexpr int C0 = (RS == 64) ? 32 : 0;
    stage_blob(sw, sc.w[ENS_LEVEL_MIDDLE] + MlpPack<32>::total(), MlpPackV2<32>::total());
    const Vox v = make_vox(pn, sc.dims[ENS_LEVEL_MIDDLE]);
    gather_warp<RS>(sc.grid[ENS_LEVEL_MIDDLE], sc.dims[ENS_LEVEL_MIDDLE], v, crow, C0);
    __syncthreads();
    float o[1];
    mlp_mma<32, RS, 1>(sw, crow, C0, p32[0], p32[1], p32[2], o);
    raw.w = o[0];
  }
  if (STAGE == ENS_STAGE_FINE || STAGE == ENS_STAGE_COLOR) {
    __syncthreads();
    stage_blob(sw, sc.w[ENS_LEVEL_FINE] + MlpPack<64>::total(), MlpPackV2<64>::total());
    const Vox v = make_vox(pn, sc.dims[ENS_LEVEL_FINE]);
    gather_warp<RS>(sc.grid[ENS_LEVEL_FINE], sc.dims[ENS_LEVEL_FINE], v, crow, 0);
    __syncthreads();
    float o[1];
    mlp_mma<64, RS, 1>(sw, crow, 0, p32[0], p32[1], p32[2], o);
    raw.w = __fadd_rn(o[0], raw.w);                               // fine_occ + middle_occ
  }
  if (STAGE == ENS_STAGE_COLOR) {
    __syncthreads();
    stage_blob(sw, sc.w[ENS_LEVEL_COLOR] + MlpPack<32>::total(), MlpPackV2<32>::total());
    const Vox v = make_vox(pn, sc.dims[ENS_LEVEL_COLOR]);
    gather_warp<RS>(sc.grid[ENS_LEVEL_COLOR], sc.dims[ENS_LEVEL_COLOR], v, crow, 0);
    __syncthreads();
    float o[4];
    mlp_mma<32, RS, 4>(sw, crow, 0, p32[0], p32[1], p32[2], o);
    raw.x = o[0]; raw.y = o[1]; raw.z = o[2];
  }
  return raw;
}

// ---------------------------------------------------------------------------------------------
// eval_points (mma)
// ---------------------------------------------------------------------------------------------
template <int STAGE, bool F64>
__global__ void __launch_bounds__(NT_MMA) eval_points_mma_kernel(DevScene sc, const void *__restrict__ pts, int64_t n,
                                                                 int apply_mask, float *__restrict__ out4) {
  extern __shared__ __align__(16) float smem[];
  float *sw = smem;
  float *sfeat = smem + MmaStage<STAGE>::WMAX;
  const int64_t t = (int64_t)blockIdx.x * NT_MMA + threadIdx.x;
  const bool valid = t < n;
  float pn[3], p32[3];
  bool inside = true;
  if (F64) {
    double p[3] = {0.0, 0.0, 0.0};
    if (valid) { const double *pp = (const double *)pts + t * 3; p[0] = pp[0]; p[1] = pp[1]; p[2] = pp[2]; }
    normalize64(p, sc.lo, sc.hi, pn);
#pragma unroll
    for (int k = 0; k < 3; ++k) { p32[k] = __double2float_rn(p[k]); inside &= (p[k] < sc.hi[k]) && (p[k] > sc.lo[k]); }
  } else {
    if (valid) { const float *pp = (const float *)pts + t * 3; p32[0] = pp[0]; p32[1] = pp[1]; p32[2] = pp[2]; }
    else { p32[0] = p32[1] = p32[2] = 0.f; }
    normalize32(p32, sc.lo, sc.hi, pn);
#pragma unroll
    for (int k = 0; k < 3; ++k)
      inside &= (p32[k] < __double2float_rn(sc.hi[k])) && (p32[k] > __double2float_rn(sc.lo[k]));
  }
  float4 raw = decode_stage_mma<STAGE>(sc, sw, sfeat, pn, p32);
  if (apply_mask && !inside) raw.w = 100.f;
  if (valid) reinterpret_cast<float4 *>(out4)[t] = raw;
}

// ---------------------------------------------------------------------------------------------
// render forward (mma): same structure as the fma variant, different decode
// ---------------------------------------------------------------------------------------------
template <int STAGE>
__global__ void __launch_bounds__(NT_MMA) render_fwd_mma_kernel(FwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  constexpr int NT = NT_MMA;
  constexpr int RS = MmaStage<STAGE>::RS;
  float *sw = smem;
  float *sfeat = smem + MmaStage<STAGE>::WMAX;
  double *zc = reinterpret_cast<double *>(sfeat + NT * RS);
  double *zs = zc + NT;
  float4 *sraw = reinterpret_cast<float4 *>(zs + NT);
  float *salpha = reinterpret_cast<float *>(sraw + NT);
  float *sT = salpha + NT;

  const RayArgs &ra = a.ra;
  const int S = ra.S;
  const int rl = threadIdx.x / S, s = threadIdx.x % S;
  const int64_t ray = (int64_t)blockIdx.x * ra.rpc + rl;
  const bool valid = (rl < ra.rpc) && (ray < ra.R);
  float o[3] = {0.f, 0.f, 0.f}, d[3] = {0.f, 0.f, 0.f};
  if (valid) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { o[k] = ra.rays_o[ray * 3 + k]; d[k] = ra.rays_d[ray * 3 + k]; }
  }
  const double z = place_sample(ra, a.sc, valid, ray, rl, s, o, d, zc, zs);
  double p[3];
  float pn[3], p32[3];
  bool inside = true;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    p[k] = __dadd_rn((double)o[k], __dmul_rn((double)d[k], z));
    p32[k] = __double2float_rn(p[k]);
    inside &= (p[k] < a.sc.hi[k]) && (p[k] > a.sc.lo[k]);
  }
  normalize64(p, a.sc.lo, a.sc.hi, pn);
  float4 raw = decode_stage_mma<STAGE>(a.sc, sw, sfeat, pn, p32);
  if (!inside) raw.w = 100.f;
  // ---- compositing (common.py:285-296) ----
  const float alpha = 1.f / (1.f + expf(-(10.f * raw.w)));
  __syncthreads();
  zs[threadIdx.x] = z;
  sraw[threadIdx.x] = raw;
  salpha[threadIdx.x] = alpha;
  __syncthreads();
  if (valid && s == 0) {
    float T = 1.f;
    for (int k = 0; k < S; ++k) {
      sT[threadIdx.x + k] = T;
      T = __fmul_rn(T, __fadd_rn(__fsub_rn(1.f, salpha[threadIdx.x + k]), 1e-10f));
    }
  }
  __syncthreads();
  const float w = __fmul_rn(alpha, sT[threadIdx.x]);
  __syncthreads();
  salpha[threadIdx.x] = w;
  __syncthreads();
  if (valid && s == 0) {
    double dep = 0.0;
    float cr = 0.f, cg = 0.f, cb = 0.f;
    for (int k = 0; k < S; ++k) {
      const float wk = salpha[threadIdx.x + k];
      const float4 rk = sraw[threadIdx.x + k];
      dep = __dadd_rn(dep, __dmul_rn((double)wk, zs[threadIdx.x + k]));
      cr = __fadd_rn(cr, __fmul_rn(wk, rk.x)); cg = __fadd_rn(cg, __fmul_rn(wk, rk.y)); cb = __fadd_rn(cb, __fmul_rn(wk, rk.z));
    }
    double var = 0.0;
    for (int k = 0; k < S; ++k) {
      const double tmp = __dsub_rn(zs[threadIdx.x + k], dep);
      var = __dadd_rn(var, __dmul_rn(__dmul_rn((double)salpha[threadIdx.x + k], tmp), tmp));
    }
    a.depth[ray] = dep;
    a.var[ray] = var;
    a.color[ray * 3 + 0] = cr; a.color[ray * 3 + 1] = cg; a.color[ray * 3 + 2] = cb;
  }
  if (valid) {
    const int64_t pi = ray * S + s;
    if (a.z_out) a.z_out[pi] = z;
    if (a.w_out) a.w_out[pi] = w;
    if (a.raw_out) reinterpret_cast<float4 *>(a.raw_out)[pi] = raw;
  }
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
template <int STAGE>
static int launch_eval_mma(const DevScene &sc, const void *pts, int f64, int64_t n, int am, float *out4, cudaStream_t s) {
  const size_t smem = (size_t)(MmaStage<STAGE>::WMAX + NT_MMA * MmaStage<STAGE>::RS) * 4;
  const unsigned g = (unsigned)((n + NT_MMA - 1) / NT_MMA);
  if (f64) {
    if (cudaFuncSetAttribute(eval_points_mma_kernel<STAGE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return ENS_ECUDA;
    eval_points_mma_kernel<STAGE, true><<<g, NT_MMA, smem, s>>>(sc, pts, n, am, out4);
  } else {
    if (cudaFuncSetAttribute(eval_points_mma_kernel<STAGE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return ENS_ECUDA;
    eval_points_mma_kernel<STAGE, false><<<g, NT_MMA, smem, s>>>(sc, pts, n, am, out4);
  }
  ENS_CHECK_CUDA();
  return ENS_OK;
}

int mma_eval_points(const DevScene &sc, int stage, const void *pts, int pts_is_f64, int64_t n, int apply_mask,
                    float *out4, cudaStream_t s) {
  switch (stage) {
    case ENS_STAGE_MIDDLE: return launch_eval_mma<ENS_STAGE_MIDDLE>(sc, pts, pts_is_f64, n, apply_mask, out4, s);
    case ENS_STAGE_FINE: return launch_eval_mma<ENS_STAGE_FINE>(sc, pts, pts_is_f64, n, apply_mask, out4, s);
    case ENS_STAGE_COLOR: return launch_eval_mma<ENS_STAGE_COLOR>(sc, pts, pts_is_f64, n, apply_mask, out4, s);
    default: return ENS_EUNSUPPORTED;
  }
}

template <int STAGE>
static int launch_fwd_mma(FwdArgs &a, cudaStream_t s) {
  constexpr int NT = NT_MMA;
  a.ra.rpc = NT / a.ra.S;
  const size_t smem = (size_t)(MmaStage<STAGE>::WMAX + NT * MmaStage<STAGE>::RS) * 4 + (size_t)NT * (8 + 8 + 16 + 4 + 4);
  if (cudaFuncSetAttribute(render_fwd_mma_kernel<STAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return ENS_ECUDA;
  const unsigned g = (unsigned)((a.ra.R + a.ra.rpc - 1) / a.ra.rpc);
  render_fwd_mma_kernel<STAGE><<<g, NT, smem, s>>>(a);
  ENS_CHECK_CUDA();
  return ENS_OK;
}

int mma_render_fwd(FwdArgs &a, int stage, cudaStream_t s) {
  switch (stage) {
    case ENS_STAGE_MIDDLE: return launch_fwd_mma<ENS_STAGE_MIDDLE>(a, s);
    case ENS_STAGE_FINE: return launch_fwd_mma<ENS_STAGE_FINE>(a, s);
    case ENS_STAGE_COLOR: return launch_fwd_mma<ENS_STAGE_COLOR>(a, s);
    default: return ENS_EUNSUPPORTED;
  }
}

}  // namespace ens

// tcgen05 probe 2 (sm_100a): the operand forms the tcgen05 backward needs, each checked against a CPU reference.
//   mode 0: TS, A [128 x 32] in TMEM, B = the TRANSPOSE view (MN-major descriptor) of a canonical K-major [32 x 32] matrix
//           D[m][c] = sum_r A[m][r] W[r][c]                        (data gradient  g_x = g_u W  from the forward's blob)
//   mode 1: same with a canonical [32 x 96] matrix, N = 96        (g_e = g_u W0)
//   mode 2: SS, A K-major from the point-staging layout [fg][pt][4] (fg = k/4), B as in mode 0
//   mode 3: SS "weight gradient": A = X staged [fg][pt][4] read MN-major (M = column of X, K = point), B = G staged the same
//           way (N = 32): D[j][n] = sum_pt X[pt][j] G[pt][n], K = 128 points (16 k-steps), M = 128
//   mode 4: as mode 3 with N = 16
//   mode 5: as mode 3 with M = 64: dumps all 128 lanes so the host can find which lane holds which row
//   mode 6: mode 3 issued as 3 x 16 MMAs (the 3xTF32 pattern) for timing
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe2 tc_probe2.cu ; run on a B200.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__host__ __device__ inline int canon_off_floats(int r, int k, int K) {
  return (r / 8) * (K / 4) * 32 + (k / 4) * 32 + (r % 8) * 4 + (k % 4);
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// SWIZZLE_128B_BASE32B (layout type 1): the only MN-major layout of 32-bit operands
__device__ __forceinline__ uint64_t make_desc_sw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return make_desc(saddr, lbo_bytes, sbo_bytes) | ((uint64_t)1 << 61);
}
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
      :: "r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
      :: "r"(d_tmem), "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
__device__ __forceinline__ void commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void wait_bar(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
#define TMEM_LD32(r, taddr)                                                                                   \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                      \
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "  \
               "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"                        \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), \
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), \
                 "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), \
                 "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) \
               : "r"(taddr) : "memory")
#define TMEM_ST32(taddr, r)                                                                                   \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                                \
               "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, " \
               "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n"                               \
               :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), \
                 "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), \
                 "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), \
                 "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) \
               : "memory")

__host__ __device__ constexpr uint32_t idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// inputs (row-major, global):  A [128][32], W32 [32][32], W96 [32][96], X [128 pts][128 cols], G [128 pts][32]
// output D: [128 lanes][96 cols]
__global__ void __launch_bounds__(128) probe_kernel(const float *__restrict__ A, const float *__restrict__ W32,
                                                    const float *__restrict__ W96, const float *__restrict__ X,
                                                    const float *__restrict__ G, float *__restrict__ D, int mode,
                                                    long long *cycles, int ks0, int ks1, float *dump) {
  extern __shared__ __align__(128) float smem_raw[];
  float *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u) / 4;   // 1024-aligned in the SHARED address space
  float *sW32 = smem;                    // canonical K-major [32][32]
  float *sW96 = sW32 + 32 * 32;          // canonical K-major [32][96]
  float *sA = sW96 + 32 * 96;            // [fg = 8][pt = 128][4]
  float *sX = sA + 128 * 32;             // [fg = 32][pt = 128][4]
  float *sG = sX + 128 * 128;            // [fg = 8][pt = 128][4]
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;

  for (int i = tid; i < 32 * 32; i += 128) sW32[canon_off_floats(i / 32, i % 32, 32)] = W32[i];
  for (int i = tid; i < 32 * 96; i += 128) sW96[canon_off_floats(i / 96, i % 96, 96)] = W96[i];
  for (int k = 0; k < 32; ++k) sA[(k / 4) * 512 + tid * 4 + (k % 4)] = A[tid * 32 + k];
  for (int k = 0; k < 128; ++k) sX[(k / 32) * 4096 + tid * 32 + ((((k % 32) / 8) ^ (tid & 3)) * 8) + (k % 8)] = X[tid * 128 + k];
  for (int k = 0; k < 32; ++k) sG[tid * 32 + (((k / 8) ^ (tid & 3)) * 8) + (k % 8)] = G[tid * 32 + k];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_base_s;
  const uint32_t lane_off = (uint32_t)(32 * warp) << 16;
  const uint32_t tD = tbase, tA = tbase + 128;        // D: up to 96 columns; A operand: 32 columns

  {   // zero the accumulator columns so untouched lanes read as 0 (mode 5), and put A into TMEM (modes 0, 1)
    uint32_t z[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) z[k] = 0u;
    TMEM_ST32(tD + lane_off, z);
    TMEM_ST32(tD + 32 + lane_off, z);
    TMEM_ST32(tD + 64 + lane_off, z);
    uint32_t r[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) r[k] = __float_as_uint(A[tid * 32 + k]);
    TMEM_ST32(tA + lane_off, r);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }

  long long t0 = clock64();
  if (tid == 0) {
    if (mode == 0) {
      // B: transpose view of sW32: n = column c, k = row r.  MN-major: SBO = stride between 4-column groups = 128 B,
      // LBO = stride between 8-row groups = (32/4)*128 B; one instruction covers one 8-row group
      for (int ks = 0; ks < 4; ++ks)
        mma_ts(tD, tA + 8 * ks, make_desc(smem_u32(sW32) + ks * 1024, 1024, 128), idesc(128, 32, 0, 1), ks > 0);
    } else if (mode == 1) {
      for (int ks = 0; ks < 4; ++ks)
        mma_ts(tD, tA + 8 * ks, make_desc(smem_u32(sW96) + ks * 3072, 3072, 128), idesc(128, 96, 0, 1), ks > 0);
    } else if (mode == 2) {
      // A K-major from [fg][pt][4]: SBO (8-point groups) = 128 B, LBO (next 4 k) = 2048 B; k-step = 2 fg = 4096 B
      for (int ks = 0; ks < 4; ++ks)
        mma_ss(tD, make_desc(smem_u32(sA) + ks * 4096, 2048, 128), make_desc(smem_u32(sW32) + ks * 1024, 1024, 128),
               idesc(128, 32, 0, 1), ks > 0);
    } else if (mode == 3 || mode == 4 || mode == 5 || mode == 11) {
      // A = X MN-major SW128_32B: rows of 128 B (32 columns) per point, atoms of 4 points; SBO (next 4 points) = 512 B,
      // LBO (next 32 columns) = 16 KB; a k-step (8 points) = 1024 B.  B = G the same (one 32-column block).
      const int M = (mode == 5 || mode == 11) ? 64 : 128, N = (mode == 4) ? 16 : 32;
      const uint32_t td = (mode == 11) ? tD + (16u << 16) : tD;
      for (int ks = ks0; ks < ks1; ++ks)
        mma_ss(td, make_desc_sw(smem_u32(sX) + ks * 1024, 16384, 512), make_desc_sw(smem_u32(sG) + ks * 1024, 16384, 512),
               idesc(M, N, 1, 1), ks > ks0);
    } else if (mode == 6 || mode == 12) {
      const int M = (mode == 12) ? 64 : 128;
      for (int pass = 0; pass < 3; ++pass)
        for (int ks = 0; ks < 16; ++ks)
          mma_ss(tD, make_desc_sw(smem_u32(sX) + ks * 1024, 16384, 512), make_desc_sw(smem_u32(sG) + ks * 1024, 16384, 512),
                 idesc(M, 32, 1, 1), (pass | ks) > 0);
    } else if (mode == 7) {
      for (int ks = 0; ks < 4; ++ks)
        mma_ts(tD, tA + 8 * ks, make_desc(smem_u32(sW32) + ks * 256, 128, 1024), idesc(128, 32, 0, 0), ks > 0);
    } else if (mode == 8) {
      for (int ks = 0; ks < 4; ++ks)
        mma_ss(tD, make_desc(smem_u32(sA) + ks * 4096, 2048, 128), make_desc(smem_u32(sW32) + ks * 256, 128, 1024),
               idesc(128, 32, 0, 0), ks > 0);
    } else if (mode == 9) {
      for (int ks = 0; ks < 4; ++ks)
        mma_ts(tD, tA + 8 * ks, make_desc(smem_u32(sW32) + ks * 1024, 128, 1024), idesc(128, 32, 0, 1), ks > 0);
    }
    commit(&bar);
  }
  wait_bar(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  long long t1 = clock64();
  for (int c0 = 0; c0 < 96; c0 += 32) {
    uint32_t r[32];
    TMEM_LD32(r, tD + c0 + lane_off);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int n = 0; n < 32; ++n) D[tid * 96 + c0 + n] = __uint_as_float(r[n]);
  }
  if (tid == 0 && cycles) *cycles = t1 - t0;
  if (mode == 3 && dump) { for (int i = tid; i < 128 * 128; i += 128) dump[i] = sX[i]; for (int i = tid; i < 128 * 32; i += 128) dump[128 * 128 + i] = sG[i]; if (tid == 0) { dump[128*128+128*32] = __uint_as_float(smem_u32(sX)); dump[128*128+128*32+1] = __uint_as_float(smem_u32(sG)); } }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"(256u) : "memory");
}

static float tr(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; float y; memcpy(&y, &u, 4); return y; }

int main() {
  std::vector<float> A(128 * 32), W32(32 * 32), W96(32 * 96), X(128 * 128), G(128 * 32), D(128 * 96);
  srand(1);
  auto fill = [](std::vector<float> &v) { for (auto &x : v) x = (float)rand() / RAND_MAX * 2.f - 1.f; };
  fill(A); fill(W32); fill(W96); fill(X); fill(G);
  float *dA, *dW32, *dW96, *dX, *dG, *dD; long long *dc;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dW32, W32.size() * 4)); CK(cudaMalloc(&dW96, W96.size() * 4));
  CK(cudaMalloc(&dX, X.size() * 4)); CK(cudaMalloc(&dG, G.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4)); CK(cudaMalloc(&dc, 8));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dW32, W32.data(), W32.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dW96, W96.data(), W96.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dG, G.data(), G.size() * 4, cudaMemcpyHostToDevice));
  float *ddump; CK(cudaMalloc(&ddump, (128 * 128 + 128 * 32 + 2) * 4));
  const size_t smem = (size_t)(32 * 32 + 32 * 96 + 128 * 32 + 128 * 128 + 128 * 32) * 4 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int mode = 0; mode < 13; ++mode) {
    if (mode == 10) continue;
    CK(cudaMemset(dD, 0, D.size() * 4));
    probe_kernel<<<1, 128, smem>>>(dA, dW32, dW96, dX, dG, dD, mode, dc, 0, 16, ddump);
    CK(cudaGetLastError());
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 1; }
    long long cyc; CK(cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    if (mode == 3) {
      FILE *f = fopen("gpurun_out/probe2_dump.bin", "wb");
      std::vector<float> dd(128 * 128 + 128 * 32 + 2); CK(cudaMemcpy(dd.data(), ddump, dd.size() * 4, cudaMemcpyDeviceToHost));
      if (f) { fwrite(D.data(), 4, D.size(), f); fwrite(X.data(), 4, X.size(), f); fwrite(G.data(), 4, G.size(), f); fwrite(dd.data(), 4, dd.size(), f); fclose(f); }
    }
    double err = 0, mx = 0;
    if (mode == 7 || mode == 8) {
      for (int m = 0; m < 128; ++m)
        for (int r = 0; r < 32; ++r) {
          double s = 0;
          for (int c = 0; c < 32; ++c) s += (double)tr(A[m * 32 + c]) * (double)tr(W32[r * 32 + c]);
          err = fmax(err, fabs(D[m * 96 + r] - s)); mx = fmax(mx, fabs(s));
        }
      printf("mode %d (control, K-major B): max|D - tf32_ref| = %.3e (max|ref| %.2f)  cycles %lld\n", mode, err, mx, cyc);
    } else if (mode <= 2 || mode == 9) {
      const int N = (mode == 1) ? 96 : 32;
      const std::vector<float> &W = (mode == 1) ? W96 : W32;
      for (int m = 0; m < 128; ++m)
        for (int c = 0; c < N; ++c) {
          double s = 0;
          for (int r = 0; r < 32; ++r) s += (double)tr(A[m * 32 + r]) * (double)tr(W[r * N + c]);
          err = fmax(err, fabs(D[m * 96 + c] - s)); mx = fmax(mx, fabs(s));
        }
      printf("mode %d: max|D - tf32_ref| = %.3e (max|ref| %.2f)  cycles %lld  D[0][0..3] = %g %g %g %g\n", mode, err, mx, cyc, D[0], D[1], D[2], D[3]);
    } else {
      const int N = (mode == 4) ? 16 : 32;
      const double scale = (mode == 6 || mode == 12) ? 3.0 : 1.0;
      std::vector<double> ref(128 * N);
      for (int j = 0; j < 128; ++j)
        for (int n = 0; n < N; ++n) {
          double s = 0;
          for (int p = 0; p < 128; ++p) s += (double)tr(X[p * 128 + j]) * (double)tr(G[p * 32 + n]);
          ref[j * N + n] = s * scale;
        }
      if (mode != 5 && mode != 11 && mode != 12) {
        for (int j = 0; j < 128; ++j)
          for (int n = 0; n < N; ++n) { err = fmax(err, fabs(D[j * 96 + n] - ref[j * N + n])); mx = fmax(mx, fabs(ref[j * N + n])); }
        printf("mode %d: max|D - tf32_ref| = %.3e (max|ref| %.2f)  cycles %lld\n", mode, err, mx, cyc);
        for (int j = 0; j < 3; ++j) printf("   row %d: D %g %g %g %g | ref %g %g %g %g\n", j, D[j * 96], D[j * 96 + 1], D[j * 96 + 2], D[j * 96 + 3], ref[j * N], ref[j * N + 1], ref[j * N + 2], ref[j * N + 3]);
      } else {
        printf("mode %d (M = 64): row -> lane map:", mode);
        for (int j = 0; j < 64; ++j) {
          int found = -1;
          for (int l = 0; l < 128 && found < 0; ++l) {
            double e2 = 0;
            for (int n = 0; n < N; ++n) e2 = fmax(e2, fabs(D[l * 96 + n] - ref[j * N + n]));
            if (e2 < 1e-3) found = l;
          }
          printf(" %d:%d", j, found);
        }
        int nz = 0;
        for (int l = 0; l < 128; ++l) { bool any = false; for (int n = 0; n < N; ++n) any |= D[l * 96 + n] != 0.f; nz += any; }
        printf("\n   lanes written: %d  cycles %lld\n", nz, cyc);
      }
    }
  }
  const int exps[6][2] = {{0, 1}, {1, 2}, {0, 2}, {2, 3}, {4, 5}, {0, 4}};
  for (int e = 0; e < 6; ++e) {
    CK(cudaMemset(dD, 0, D.size() * 4));
    probe_kernel<<<1, 128, smem>>>(dA, dW32, dW96, dX, dG, dD, 3, dc, exps[e][0], exps[e][1], nullptr);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double err = 0, mx = 0;
    for (int j = 0; j < 128; ++j)
      for (int n = 0; n < 32; ++n) {
        double s2 = 0;
        for (int p = 8 * exps[e][0]; p < 8 * exps[e][1]; ++p) s2 += (double)tr(X[p * 128 + j]) * (double)tr(G[p * 32 + n]);
        err = fmax(err, fabs(D[j * 96 + n] - s2)); mx = fmax(mx, fabs(s2));
      }
    printf("k-steps [%d,%d): max err %.3e (max|ref| %.2f)\n", exps[e][0], exps[e][1], err, mx);
  }
  return 0;
}

// Touched-voxel compaction of the dense grid gradient, for the sharded mapping step.
//
// ens_render_bwd accumulates a DENSE native-layout gradient [Z][Y][X][32] per level (48 MB for room0); a 1000-ray batch
// touches ~5 % of its voxels.  All-reducing the dense arena is what limits the 1 -> 8 GPU curve of the mapping step, so the
// ranks exchange only the rows any of them touched:
//   ens_grid_touched   flags[v] = 1 iff any of the 32 channels of voxel v is non-zero            (8 lanes per 128-byte line)
//   (host: MAX all-reduce of the flags over the ranks -> the same union on every rank; inclusive scan -> positions)
//   ens_grid_compact   direction 1: compact[pos[v] - 1] = grad[v] for flagged v;  direction 0: the reverse (after the SUM
//                      all-reduce of `compact`); rows past `capacity` are dropped and counted in *overflow.
// Replaces nothing in the reference (it has no multi-GPU path); the data layout is that of ens_render_bwd's gradient sinks.
#include "ens_common.cuh"

namespace ens {

__global__ void __launch_bounds__(256) grid_touched_kernel(const float4 *__restrict__ g, int64_t n_vox, int32_t *__restrict__ flags) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t v = t >> 3;
  float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
  if (v < n_vox) x = g[t];
  unsigned nz = (x.x != 0.f) | (x.y != 0.f) | (x.z != 0.f) | (x.w != 0.f);
  nz |= __shfl_xor_sync(0xffffffffu, nz, 1);
  nz |= __shfl_xor_sync(0xffffffffu, nz, 2);
  nz |= __shfl_xor_sync(0xffffffffu, nz, 4);
  if (v < n_vox && (threadIdx.x & 7) == 0) flags[v] = (int32_t)nz;
}

template <bool TO_COMPACT>
__global__ void __launch_bounds__(256) grid_compact_kernel(float4 *__restrict__ g, const int32_t *__restrict__ flags,
                                                           const int32_t *__restrict__ pos, int64_t n_vox, float4 *__restrict__ compact,
                                                           int64_t capacity, int32_t *__restrict__ overflow) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t v = t >> 3;
  if (v >= n_vox || flags[v] == 0) return;
  const int64_t row = (int64_t)pos[v] - 1;
  if (row >= capacity) {
    if ((threadIdx.x & 7) == 0) atomicAdd(overflow, 1);
    return;
  }
  const int q = threadIdx.x & 7;
  if (TO_COMPACT) compact[row * 8 + q] = g[t];
  else g[t] = compact[row * 8 + q];
}

}  // namespace ens

using namespace ens;

extern "C" int ens_grid_touched(const float *grad, int64_t n_vox, int32_t *flags, ens_stream_t stream) {
  if (!grad || !flags || n_vox < 0) return ENS_EINVAL;
  if (n_vox == 0) return ENS_OK;
  const int64_t threads = n_vox * 8;
  grid_touched_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4 *>(grad), n_vox, flags);
  ENS_CHECK_CUDA();
  return ENS_OK;
}

extern "C" int ens_grid_compact(float *grad, const int32_t *flags, const int32_t *pos, int64_t n_vox, float *compact,
                                int64_t capacity, int to_compact, int32_t *overflow, ens_stream_t stream) {
  if (!grad || !flags || !pos || !compact || !overflow || n_vox < 0 || capacity < 0) return ENS_EINVAL;
  if (n_vox == 0) return ENS_OK;
  const int64_t threads = n_vox * 8;
  const unsigned nb = (unsigned)((threads + 255) / 256);
  if (to_compact)
    grid_compact_kernel<true><<<nb, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<float4 *>(grad), flags, pos, n_vox,
                                                                    reinterpret_cast<float4 *>(compact), capacity, overflow);
  else
    grid_compact_kernel<false><<<nb, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<float4 *>(grad), flags, pos, n_vox,
                                                                     reinterpret_cast<float4 *>(compact), capacity, overflow);
  ENS_CHECK_CUDA();
  return ENS_OK;
}

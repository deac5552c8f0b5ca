// Fused frustum-masked Adam step on the feature grids, in the render kernels' layout.
//
// What it replaces, per mapping iteration and grid level (paths under /root/reference/src):
//   Mapper.py:343-361  val_grad = val[mask].clone() -> Variable             (boolean-mask gather, implicit nonzero + sync)
//   Mapper.py:451-458  val[mask] = val_grad                                 (index_put before every render)
//   Mapper.py:625      optimizer.step()  (torch.optim.Adam on the group)    (4 elementwise passes per tensor)
//   Mapper.py:633-641  val[mask] = val_grad.clone().detach()                (write-back)
// The dense gradient the render backward produced is consumed in place: the voxels named by the (ascending) index list
// -- the nonzero positions of the frustum mask, built once per optimize_map call, where the reference re-derives them
// inside every boolean-mask index_put -- are updated (all 32 channels: the reference repeats the voxel mask over the
// channel axis, Mapper.py:345-346); every other voxel is left untouched, exactly as if only the gathered copy had been
// optimised and written back.
//
// HBM-bound by construction: a selected voxel moves 4 lines in (value, gradient, both moments) and 3 out
// (+1 when the gradient is cleared); an unselected voxel costs nothing.
// One 128-byte line per voxel -> 8 threads x float4, a warp covers 4 selected voxels (512 contiguous bytes where the
// selection is contiguous, as it is along X inside a frustum).
#include "ens_common.cuh"

namespace ens {

struct AdamLevel {
  float *grid, *grad, *m, *v;
  const int32_t *index;          // selected voxels, ascending; nullptr = voxels 0..n_sel-1
  long long n_sel;
  float lr;
};

struct AdamArgs {
  AdamLevel lv[4];
  int n_levels;
  long long total_sel;
  float beta1, beta2, one_m_beta1, one_m_beta2, eps;
  float inv_bc2_sqrt;            // host-computed from `step` (double), used when dyn == nullptr; lv[].lr is then lr/bc1
  double beta1_d, beta2_d;
  const double *dyn;             // device [1 + n_levels]: step, lr per level -- read at run time (CUDA-graph replays)
  int clear_grad;
};

constexpr int ADAM_THREADS = 256;
constexpr int ADAM_UNROLL = 4;

__device__ __forceinline__ float4 ldg_stream(const float *p) { return __ldcs(reinterpret_cast<const float4 *>(p)); }

__global__ void __launch_bounds__(ADAM_THREADS) grid_adam_kernel(const AdamArgs a) {
  __shared__ float s_lr[4];
  __shared__ float s_ibc2s;
  __shared__ AdamLevel s_lv[4];
  if (threadIdx.x < 4) s_lv[threadIdx.x] = a.lv[threadIdx.x];
  if (threadIdx.x == 0) {
    if (a.dyn) {
      const double step = a.dyn[0];
      const double bc1 = 1.0 - pow(a.beta1_d, step);
      s_ibc2s = (float)(1.0 / sqrt(1.0 - pow(a.beta2_d, step)));
      for (int l = 0; l < a.n_levels; ++l) s_lr[l] = (float)(a.dyn[1 + l] / bc1);   // step_size = lr / bias_correction1
    } else {
      s_ibc2s = a.inv_bc2_sqrt;
      for (int l = 0; l < a.n_levels; ++l) s_lr[l] = a.lv[l].lr;                     // already lr / bias_correction1
    }
  }
  __syncthreads();
  const float ibc2s = s_ibc2s;
  const int sub = threadIdx.x & 7;                              // float4 slot inside the voxel line
  const long long vox_per_iter = (long long)gridDim.x * (ADAM_THREADS / 8);
  long long vox0 = (long long)blockIdx.x * (ADAM_THREADS / 8) + (threadIdx.x >> 3);

  for (; vox0 < a.total_sel; vox0 += vox_per_iter * ADAM_UNROLL) {
    float4 p[ADAM_UNROLL], g[ADAM_UNROLL], m[ADAM_UNROLL], v[ADAM_UNROLL];
    float *pp[ADAM_UNROLL], *pg[ADAM_UNROLL], *pm[ADAM_UNROLL], *pv[ADAM_UNROLL];
    float lr[ADAM_UNROLL];
    bool live[ADAM_UNROLL];
    long long item[ADAM_UNROLL];
    int lvl[ADAM_UNROLL];
    int vox[ADAM_UNROLL];
#pragma unroll
    for (int u = 0; u < ADAM_UNROLL; ++u) {                     // the position in the concatenated selections
      long long it = vox0 + (long long)u * vox_per_iter;
      live[u] = it < a.total_sel;
      int l = 0;
      if (live[u])
        while (l + 1 < a.n_levels && it >= s_lv[l].n_sel) { it -= s_lv[l].n_sel; ++l; }
      item[u] = it; lvl[u] = l;
    }
#pragma unroll
    for (int u = 0; u < ADAM_UNROLL; ++u)                       // index loads of all unrolled items in flight together
      vox[u] = !live[u] ? 0 : (s_lv[lvl[u]].index ? __ldg(s_lv[lvl[u]].index + item[u]) : (int)item[u]);
#pragma unroll
    for (int u = 0; u < ADAM_UNROLL; ++u) {
      const AdamLevel &L = s_lv[lvl[u]];
      const long long off = (long long)vox[u] * 32 + sub * 4;
      pp[u] = L.grid + off; pg[u] = L.grad + off; pm[u] = L.m + off; pv[u] = L.v + off;
      lr[u] = s_lr[lvl[u]];
    }
#pragma unroll
    for (int u = 0; u < ADAM_UNROLL; ++u)
      if (live[u]) {
        g[u] = ldg_stream(pg[u]);
        p[u] = ldg_stream(pp[u]);
        m[u] = ldg_stream(pm[u]);
        v[u] = ldg_stream(pv[u]);
      }
#pragma unroll
    for (int u = 0; u < ADAM_UNROLL; ++u) {
      if (live[u]) {
        const float step_size = lr[u];
        float *pe = &p[u].x, *ge = &g[u].x, *me = &m[u].x, *ve = &v[u].x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // torch.optim.Adam (reference pin 1.11, optim/_functional.py adam): exp_avg.mul_(b1).add_(g, alpha=1-b1);
          // exp_avg_sq.mul_(b2).addcmul_(g, g, value=1-b2); denom = sqrt(exp_avg_sq)/sqrt(bc2) + eps;
          // param.addcdiv_(exp_avg, denom, value=-lr/bc1)
          const float gk = ge[k];
          const float mk = fmaf(gk, a.one_m_beta1, me[k] * a.beta1);
          const float vk = fmaf(gk * gk, a.one_m_beta2, ve[k] * a.beta2);
          const float denom = sqrtf(vk) * ibc2s + a.eps;
          pe[k] = pe[k] - step_size * (mk / denom);
          me[k] = mk;
          ve[k] = vk;
        }
        __stcs(reinterpret_cast<float4 *>(pp[u]), p[u]);
        __stcs(reinterpret_cast<float4 *>(pm[u]), m[u]);
        __stcs(reinterpret_cast<float4 *>(pv[u]), v[u]);
      }
      if (live[u] && a.clear_grad) __stcs(reinterpret_cast<float4 *>(pg[u]), make_float4(0.f, 0.f, 0.f, 0.f));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Adam over many small tensors in one launch: the decoder weights (69 tensors, 58 956 floats) and the camera tensors of
// the reference's optimizer (Mapper.py:396-423: groups with their own learning rate).  torch.optim.Adam walks them with
// a dozen multi-tensor launches per step (more in its capturable form); here blockIdx.y is the tensor.
// ---------------------------------------------------------------------------------------------
constexpr int MT_MAX = 96;

struct MultiAdamArgs {
  float *param[MT_MAX];
  const float *grad[MT_MAX];
  int size[MT_MAX];
  int state_off[MT_MAX];
  unsigned char group[MT_MAX];
  float step_size[8];            // lr / bias_correction1 per group (dyn == nullptr)
  float *m, *v;                  // flat moments, tensor i at state_off[i]
  float beta1, beta2, one_m_beta1, one_m_beta2, eps, inv_bc2_sqrt;
  double beta1_d, beta2_d;
  const double *dyn;             // device [1 + n_groups]: step, lr per group
  int n_groups;
};

__global__ void __launch_bounds__(256) multi_adam_kernel(const __grid_constant__ MultiAdamArgs a) {
  const int ti = blockIdx.y;
  const int n = a.size[ti];
  const int i0 = blockIdx.x * 256 + threadIdx.x;
  if (blockIdx.x * 256 >= n) return;
  float step_size, ibc2s;
  if (a.dyn) {
    const double step = a.dyn[0];
    step_size = (float)(a.dyn[1 + a.group[ti]] / (1.0 - pow(a.beta1_d, step)));
    ibc2s = (float)(1.0 / sqrt(1.0 - pow(a.beta2_d, step)));
  } else {
    step_size = a.step_size[a.group[ti]];
    ibc2s = a.inv_bc2_sqrt;
  }
  float *p = a.param[ti];
  const float *g = a.grad[ti];
  float *m = a.m + a.state_off[ti], *v = a.v + a.state_off[ti];
  for (int i = i0; i < n; i += gridDim.x * 256) {
    const float gk = g[i];
    const float mk = fmaf(gk, a.one_m_beta1, m[i] * a.beta1);
    const float vk = fmaf(gk * gk, a.one_m_beta2, v[i] * a.beta2);
    const float denom = sqrtf(vk) * ibc2s + a.eps;
    p[i] = p[i] - step_size * (mk / denom);
    m[i] = mk;
    v[i] = vk;
  }
}

}  // namespace ens

using namespace ens;

extern "C" int ens_tensors_adam_step(float *const *params_host, const float *const *grads_host, const int64_t *sizes_host,
                                     const int *groups_host, int n_tensors, const double *lrs_host, int n_groups,
                                     float *exp_avg, float *exp_avg_sq, double beta1, double beta2, double eps,
                                     int64_t step, const double *dyn, ens_stream_t stream) {
  if (n_tensors < 0 || n_groups < 1 || n_groups > 8 || (step < 1 && !dyn)) return ENS_EINVAL;
  if (n_tensors == 0) return ENS_OK;
  if (!params_host || !grads_host || !sizes_host || !groups_host || !lrs_host || !exp_avg || !exp_avg_sq) return ENS_EINVAL;
  if (!(beta1 >= 0.0 && beta1 < 1.0) || !(beta2 >= 0.0 && beta2 < 1.0) || !(eps >= 0.0)) return ENS_EINVAL;
  const double bc1 = dyn ? 1.0 : 1.0 - pow(beta1, (double)step);
  int64_t off_total = 0;
  for (int base = 0; base < n_tensors; base += MT_MAX) {        // more tensors than one launch holds: several launches
    MultiAdamArgs a{};
    const int cnt = n_tensors - base < MT_MAX ? n_tensors - base : MT_MAX;
    int max_n = 0;
    for (int i = 0; i < cnt; ++i) {
      const int j = base + i;
      if (!params_host[j] || !grads_host[j] || sizes_host[j] < 0 || sizes_host[j] > 0x7fffffffLL) return ENS_EINVAL;
      if (groups_host[j] < 0 || groups_host[j] >= n_groups) return ENS_EINVAL;
      a.param[i] = params_host[j]; a.grad[i] = grads_host[j]; a.size[i] = (int)sizes_host[j];
      a.state_off[i] = (int)off_total; a.group[i] = (unsigned char)groups_host[j];
      off_total += sizes_host[j];
      if (off_total > 0x7fffffffLL) return ENS_ESHAPE;
      if (a.size[i] > max_n) max_n = a.size[i];
    }
    for (int gq = 0; gq < n_groups; ++gq) a.step_size[gq] = (float)(lrs_host[gq] / bc1);
    a.m = exp_avg; a.v = exp_avg_sq;
    a.beta1 = (float)beta1; a.beta2 = (float)beta2;
    a.one_m_beta1 = (float)(1.0 - beta1); a.one_m_beta2 = (float)(1.0 - beta2);
    a.eps = (float)eps; a.beta1_d = beta1; a.beta2_d = beta2; a.dyn = dyn; a.n_groups = n_groups;
    if (!dyn) a.inv_bc2_sqrt = (float)(1.0 / sqrt(1.0 - pow(beta2, (double)step)));
    if (max_n == 0) continue;
    int gx = (max_n + 255) / 256;
    if (gx > 64) gx = 64;
    multi_adam_kernel<<<dim3(gx, cnt), 256, 0, (cudaStream_t)stream>>>(a);
    ENS_CHECK_CUDA();
  }
  return ENS_OK;
}

extern "C" int ens_grid_adam_step(const EnsAdamLevel *levels_host, int n_levels, double beta1, double beta2, double eps,
                                  int64_t step, const double *dyn, int clear_grad, ens_stream_t stream) {
  if (!levels_host || n_levels < 0 || n_levels > 4 || (step < 1 && !dyn)) return ENS_EINVAL;
  if (!(beta1 >= 0.0 && beta1 < 1.0) || !(beta2 >= 0.0 && beta2 < 1.0) || !(eps >= 0.0)) return ENS_EINVAL;
  AdamArgs a{};
  a.n_levels = n_levels;
  a.total_sel = 0;
  for (int l = 0; l < n_levels; ++l) {
    const EnsAdamLevel &h = levels_host[l];
    if (h.n_voxels < 0 || h.n_voxels > 0x7fffffffLL || h.n_selected < 0 || h.n_selected > h.n_voxels) return ENS_EINVAL;
    if (!h.voxel_index && h.n_selected != h.n_voxels) return ENS_ESHAPE;
    if (h.n_selected > 0) {
      if (!h.grid || !h.grad || !h.exp_avg || !h.exp_avg_sq) return ENS_EINVAL;
      const uintptr_t al = (uintptr_t)h.grid | (uintptr_t)h.grad | (uintptr_t)h.exp_avg | (uintptr_t)h.exp_avg_sq;
      if (al & 15) return ENS_EINVAL;                           // float4 accesses
    }
    const double bc1 = dyn ? 1.0 : 1.0 - pow(beta1, (double)step);
    a.lv[l] = AdamLevel{h.grid, h.grad, h.exp_avg, h.exp_avg_sq, h.voxel_index, (long long)h.n_selected, (float)(h.lr / bc1)};
    a.total_sel += h.n_selected;
  }
  if (a.total_sel == 0) return ENS_OK;
  a.beta1 = (float)beta1; a.beta2 = (float)beta2;
  a.one_m_beta1 = (float)(1.0 - beta1); a.one_m_beta2 = (float)(1.0 - beta2);
  a.eps = (float)eps;
  a.beta1_d = beta1; a.beta2_d = beta2;
  a.dyn = dyn;
  if (!dyn) a.inv_bc2_sqrt = (float)(1.0 / sqrt(1.0 - pow(beta2, (double)step)));
  a.clear_grad = clear_grad ? 1 : 0;
  // persistent grid: a multiple of the SM count, enough resident warps to cover HBM latency
  const int sms = sm_count();
  const long long need = (a.total_sel + (ADAM_THREADS / 8) * ADAM_UNROLL - 1) / ((ADAM_THREADS / 8) * ADAM_UNROLL);
  long long blocks = (long long)sms * 8;
  if (blocks > need) blocks = need;
  grid_adam_kernel<<<(unsigned)blocks, ADAM_THREADS, 0, (cudaStream_t)stream>>>(a);
  ENS_CHECK_CUDA();
  return ENS_OK;
}

// UNet input assembly of the event branch as ONE launch each way.
// Replaces the tensor shuffling at the head of src/event_net.py:67-99 (inference_event):
//     img1.permute(2,0,1), img2.permute(2,0,1) [, transforms.Resize(NEAREST) of both when scale_factor != 1],
//     torch.cat(dim 0), unsqueeze(0), .to(float32)
// img1 = the previous ground-truth colour image (float64 or float32, HWC), img2 = the colour image the renderer produced
// (float32, HWC, carries the gradient of the event loss back to the camera pose).  Output: [6][h][w] float32.
// Nearest-neighbour source index as ATen's upsample_nearest2d (what torchvision's Resize runs on tensors):
//     src = min((int)floorf(dst * scale), in - 1),  scale = (float)in / out        (identity when in == out)
#include "ens_common.cuh"

namespace ens {

__device__ __forceinline__ int nearest_src(int dst, int in, int out) {
  if (in == out) return dst;
  const float scale = (float)in / (float)out;
  const int s = (int)floorf((float)dst * scale);
  return s < in - 1 ? s : in - 1;
}

template <bool F64>
__global__ void __launch_bounds__(256) unet_input_kernel(const void *__restrict__ img1, int H1, int W1, const float *__restrict__ img2,
                                                         int H2, int W2, int h, int w, float *__restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n = (int64_t)h * w;
  if (t >= n) return;
  const int y = (int)(t / w), x = (int)(t % w);
  const int64_t s1 = ((int64_t)nearest_src(y, H1, h) * W1 + nearest_src(x, W1, w)) * 3;
  const int64_t s2 = ((int64_t)nearest_src(y, H2, h) * W2 + nearest_src(x, W2, w)) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    out[c * n + t] = F64 ? (float)reinterpret_cast<const double *>(img1)[s1 + c] : reinterpret_cast<const float *>(img1)[s1 + c];
    out[(3 + c) * n + t] = img2[s2 + c];
  }
}

// d loss / d img2 (HWC) from d loss / d out[3..5] (CHW); += because several outputs may share a source pixel when h > H2
__global__ void __launch_bounds__(256) unet_input_bwd_kernel(const float *__restrict__ g_out, int H2, int W2, int h, int w,
                                                             float *__restrict__ g_img2) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n = (int64_t)h * w;
  if (t >= n) return;
  const int y = (int)(t / w), x = (int)(t % w);
  const int64_t s2 = ((int64_t)nearest_src(y, H2, h) * W2 + nearest_src(x, W2, w)) * 3;
  const bool same = (H2 == h) && (W2 == w);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float g = g_out[(3 + c) * n + t];
    if (same) g_img2[s2 + c] = g;
    else atomicAdd(g_img2 + s2 + c, g);
  }
}

}  // namespace ens

using namespace ens;

extern "C" int ens_unet_input(const void *img1, int img1_is_f64, int H1, int W1, const float *img2, int H2, int W2, int h, int w,
                              float *out, ens_stream_t stream) {
  if (!img1 || !img2 || !out) return ENS_EINVAL;
  if (H1 < 1 || W1 < 1 || H2 < 1 || W2 < 1 || h < 1 || w < 1) return ENS_ESHAPE;
  const int64_t n = (int64_t)h * w;
  const unsigned nb = (unsigned)((n + 255) / 256);
  if (img1_is_f64) unet_input_kernel<true><<<nb, 256, 0, (cudaStream_t)stream>>>(img1, H1, W1, img2, H2, W2, h, w, out);
  else unet_input_kernel<false><<<nb, 256, 0, (cudaStream_t)stream>>>(img1, H1, W1, img2, H2, W2, h, w, out);
  ENS_CHECK_CUDA();
  return ENS_OK;
}

extern "C" int ens_unet_input_bwd(const float *g_out, int H2, int W2, int h, int w, float *g_img2, ens_stream_t stream) {
  if (!g_out || !g_img2) return ENS_EINVAL;
  if (H2 < 1 || W2 < 1 || h < 1 || w < 1) return ENS_ESHAPE;
  const int64_t n = (int64_t)h * w;
  if (!(H2 == h && W2 == w)) ENS_CUDA_CALL(cudaMemsetAsync(g_img2, 0, sizeof(float) * 3 * (size_t)H2 * W2, (cudaStream_t)stream));
  unet_input_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(g_out, H2, W2, h, w, g_img2);
  ENS_CHECK_CUDA();
  return ENS_OK;
}

// Element-wise pieces of the per-image feature stage (SlotAttnCNNEncoder, ocrs/common/models.py:96-107) that sit
// between the convolutions (cuDNN, library calls) in the bf16 inference path.  HBM-bound, one pass each.
//   ocrl_conv_bias_relu_bf16 : y = relu(y + bias[c]) in place on a channels-last bf16 tensor (Conv2dBlock,
//                              ocrs/common/networks.py:38-53) -- replaces a broadcast add + clamp (two passes).
//   ocrl_frames_to_nhwc_bf16 : obs [B,C,H,W] fp32 (utils/datasets.py:17 layout) -> [B,H,W,CP] bf16, channels zero-padded
//                              to CP (a multiple of 8, so that the first convolution takes a tensor-core kernel).
#include "common.cuh"

namespace ocrl {

template <int C>
__global__ void __launch_bounds__(256) bias_relu_bf16_kernel(__nv_bfloat16* __restrict__ y, const float* __restrict__ bias,
                                                             size_t nvec) {
  // one thread handles 8 consecutive channels (16 bytes); C % 8 == 0, so a vector never straddles a pixel
  __shared__ float sb[C];
  for (int i = threadIdx.x; i < C; i += blockDim.x) sb[i] = bias[i];
  __syncthreads();
  uint4* p = reinterpret_cast<uint4*>(y);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
    uint4 v = p[i];
    const int c0 = (int)((i * 8) % C);
    uint32_t* w = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[k]);
      const float a = fmaxf(__low2float(h) + sb[c0 + 2 * k], 0.f);
      const float b = fmaxf(__high2float(h) + sb[c0 + 2 * k + 1], 0.f);
      w[k] = pack_bf16x2(a, b);
    }
    p[i] = v;
  }
}

// one thread per pixel: gathers the C planes (coalesced across the warp) and writes CP bf16 channels (16 bytes for CP = 8)
template <int CP>
__global__ void __launch_bounds__(256) frames_to_nhwc_kernel(const float* __restrict__ obs, __nv_bfloat16* __restrict__ out,
                                                             int C, int HW, size_t npix) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / HW, pix = i % HW;
    float x[CP];
#pragma unroll
    for (int c = 0; c < CP; ++c) x[c] = (c < C) ? __ldg(obs + (b * C + c) * (size_t)HW + pix) : 0.f;
    uint32_t w[CP / 2];
#pragma unroll
    for (int c = 0; c < CP / 2; ++c) w[c] = pack_bf16x2(x[2 * c], x[2 * c + 1]);
    uint4* dst = reinterpret_cast<uint4*>(out + i * CP);
#pragma unroll
    for (int q = 0; q < CP / 8; ++q) dst[q] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
  }
}

}  // namespace ocrl

using namespace ocrl;

extern "C" {

int ocrl_conv_bias_relu_bf16(void* y, const float* bias, long long npixels, int channels, void* stream) {
  if (!y || !bias || (reinterpret_cast<uintptr_t>(y) & 15u)) {
    set_error("conv_bias_relu: null or unaligned pointer");
    return OCRL_E_ALIGN;
  }
  if (npixels <= 0) return OCRL_OK;
  const size_t nvec = (size_t)npixels * channels / 8;
  const int grid = (int)((nvec + 255) / 256 < 148 * 8 ? (nvec + 255) / 256 : 148 * 8);
  switch (channels) {
    case 64: bias_relu_bf16_kernel<64><<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<__nv_bfloat16*>(y), bias, nvec); ocrl::count_launch(); break;
    case 128: bias_relu_bf16_kernel<128><<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<__nv_bfloat16*>(y), bias, nvec); ocrl::count_launch(); break;
    default:
      set_error("conv_bias_relu: channels=%d not supported (64, 128)", channels);
      return OCRL_E_SHAPE;
  }
  OCRL_CHECK_CUDA(cudaGetLastError());
  return OCRL_OK;
}

int ocrl_frames_to_nhwc_bf16(const float* obs, void* out, int B, int C, int H, int W, int CP, void* stream) {
  if (!obs || !out || (reinterpret_cast<uintptr_t>(out) & 15u)) {
    set_error("frames_to_nhwc: null or unaligned pointer");
    return OCRL_E_ALIGN;
  }
  if (CP != 8 || C > CP || C < 1) {
    set_error("frames_to_nhwc: C=%d CP=%d not supported (C <= CP = 8)", C, CP);
    return OCRL_E_SHAPE;
  }
  const size_t npix = (size_t)B * H * W;
  if (npix == 0) return OCRL_OK;
  const int grid = (int)((npix + 255) / 256 < 148 * 8 ? (npix + 255) / 256 : 148 * 8);
  frames_to_nhwc_kernel<8><<<grid, 256, 0, (cudaStream_t)stream>>>(obs, reinterpret_cast<__nv_bfloat16*>(out), C, H * W, npix);
  ocrl::count_launch();
  OCRL_CHECK_CUDA(cudaGetLastError());
  return OCRL_OK;
}

}  // extern "C"

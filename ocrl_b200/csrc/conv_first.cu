// First layer of SlotAttnCNNEncoder (ocrs/common/models.py:96-107, Conv2dBlock of ocrs/common/networks.py:38-53):
//   y = relu(conv5x5(obs, W) + b),  obs [B,C,H,W] fp32 in [0,1] (C = 3), W [64,C,5,5], stride 1, zero padding 2,
// written as bf16 channels-last [B,H,W,64] -- the input layout of the following (cuDNN) layers.
//
// With C = 3 the layer is an implicit GEMM with K = 75 only: cuDNN runs it through the same 256x64x64 tile kernel as the
// 64 -> 64 layers and needs 129 us for 1/8 of their FLOPs.  Here it is HBM-write-bound (33.5 MB of output per 64 frames):
//   * a persistent CTA walks (image, output row) pairs; the 5 input rows x C planes of a row are staged in shared
//     memory as bf16 with the zero padding materialised, straight from the fp32 NCHW frames (no ingest pass);
//   * each warp owns 16 consecutive pixels of the row: A fragments (16 pixels x 16 taps) are gathered from the staged
//     rows, B fragments (the bf16 weights, K padded to 80) live in registers for the whole kernel, mma.sync m16n8k16
//     accumulates in fp32 -- same operand rounding as the bf16 autocast path of the reference modules;
//   * bias + ReLU + bf16 rounding in registers, then the 16 x 64 tile goes through shared memory so that the global
//     stores are full 128-byte pixel rows.
#include "common.cuh"

namespace ocrl {

namespace conv1 {

constexpr int CO = 64;       // output channels
constexpr int KS = 5;        // kernel size
constexpr int KT = 5;        // k-steps of 16 taps: K = C * 25 <= 80
constexpr int NW = 8;        // warps per CTA
constexpr int RB = 4;        // output rows per staging step (RB + 4 input rows are staged)
constexpr int XP = 8;        // left margin (elements) of a staged row: x = -2 sits at XP - 2, 16-byte aligned rows

// U8: the frames are uint8 HWC [B,H,W,C] as the HDF5 datasets / environments deliver them (utils/datasets.py:17); the
// kernel applies the reference's own ingest `obs / 255.0` (fp32 division) while staging, so the result is bit-identical
// to feeding the float CHW tensor and the frames cross PCIe / NVLink at a quarter of the bytes.
template <int C, bool U8>
__global__ void __launch_bounds__(NW * 32, 2) conv_first_kernel(const void* __restrict__ obs_any, const float* __restrict__ w,
                                                                const float* __restrict__ bias, __nv_bfloat16* __restrict__ out,
                                                                int B, int H, int W, int padded) {
  static_assert(C * KS * KS <= KT * 16, "taps must fit the padded K");
  extern __shared__ __align__(16) unsigned char smem[];
  const int RP = W + 2 * XP;  // staged row pitch (elements); columns [XP-2, XP+W+2) are read
  constexpr int NR = RB + KS - 1;
  uint2* bfrag = reinterpret_cast<uint2*>(smem);                                   // [KT][CO/8][32] B fragments per lane
  __nv_bfloat16* rows = reinterpret_cast<__nv_bfloat16*>(bfrag + KT * (CO / 8) * 32);  // [C][NR][RP]
  __nv_bfloat16* stage = rows + C * NR * RP;                                 // [NW][16][CO + 8] output staging
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g8 = lane >> 2, t4 = lane & 3;
  const float* obs = reinterpret_cast<const float*>(obs_any);
  const unsigned char* obs8 = reinterpret_cast<const unsigned char*>(obs_any);

  // ---- B fragments of the weights, once per CTA: B[k][n] = W[n][c][dy][dx], k = (c*5 + dy)*5 + dx, zero for k >= C*25
  for (int i = tid; i < KT * (CO / 8) * 32; i += NW * 32) {
    const int ln = i & 31, nt = (i >> 5) % (CO / 8), ks = i / (32 * (CO / 8));
    const int n = 8 * nt + (ln >> 2);
    uint32_t v[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k0 = 16 * ks + 8 * h + 2 * (ln & 3);
      const float w0 = (k0 < C * 25) ? __ldg(w + n * C * 25 + k0) : 0.f;
      const float w1 = (k0 + 1 < C * 25) ? __ldg(w + n * C * 25 + k0 + 1) : 0.f;
      v[h] = pack_bf16x2(w0, w1);
    }
    bfrag[i] = make_uint2(v[0], v[1]);
  }
  // per-thread element offsets of its four taps per k-step inside the staged rows
  int koff[KT][4];
#pragma unroll
  for (int ks = 0; ks < KT; ++ks)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = 16 * ks + 8 * (i >> 1) + 2 * t4 + (i & 1);
      if (k < C * 25) {
        const int c = k / 25, dy = (k % 25) / 5, dx = k % 5;
        koff[ks][i] = (c * NR + dy) * RP + XP - 2 + dx;
      } else {
        koff[ks][i] = XP;  // padding tap: its weight is zero, any finite staged value will do
      }
    }
  float bv[CO / 8][2];
#pragma unroll
  for (int nt = 0; nt < CO / 8; ++nt) {
    bv[nt][0] = __ldg(bias + 8 * nt + 2 * t4);
    bv[nt][1] = __ldg(bias + 8 * nt + 2 * t4 + 1);
  }
  // the margins are written once; the staging below only touches columns [XP, XP + W)
  for (int i = tid; i < C * NR * RP; i += NW * 32) rows[i] = __float2bfloat16_rn(0.f);

  if (padded) {  // padding positions of the output array: whole rows between the images, two columns either side of a row
    const int WPd = W + 4, rows_total = 2 + B * (H + 2);
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (int g = blockIdx.x * NW + warp; g < rows_total; g += gridDim.x * NW) {
      uint4* rowp = reinterpret_cast<uint4*>(out + (size_t)g * WPd * CO);
      const bool real = g >= 2 && ((g - 2) % (H + 2)) < H;
      if (!real) {
        for (int i = lane; i < WPd * 8; i += 32) rowp[i] = z;
      } else {
        rowp[lane < 16 ? lane : (W + 2) * 8 + (lane - 16)] = z;  // positions 0, 1 and W + 2, W + 3 (8 chunks of 16 B each)
      }
    }
  }
  const int yblocks = (H + RB - 1) / RB;
  const int nblocks = B * yblocks;
  const int W4 = W / 4;
  const int tiles_x = W / 16;
  for (int blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
    const int b = blk / yblocks, y0 = (blk % yblocks) * RB;
    __syncthreads();  // the previous block's readers are done with `rows`
    // ---- stage C planes x NR input rows (zero outside the image), fp32 -> bf16, 16-byte global loads
    if (!U8) {
      for (int i = tid; i < C * NR * W4; i += NW * 32) {
        const int x4 = i % W4, p = i / W4;
        const int c = p / NR, yy = y0 + p % NR - 2;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (yy >= 0 && yy < H) v = __ldg(reinterpret_cast<const float4*>(obs + ((size_t)(b * C + c) * H + yy) * W) + x4);
        *reinterpret_cast<uint2*>(rows + p * RP + XP + 4 * x4) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
      }
    } else {
      // a frame row is W*C contiguous bytes (W a multiple of 16: whole 32-bit words); byte e of the row belongs to
      // pixel e / C, plane e % C
      const int words = W * C / 4;
      for (int i = tid; i < NR * words; i += NW * 32) {
        const int wd = i % words, r = i / words, yy = y0 + r - 2;
        uint32_t v = 0u;
        if (yy >= 0 && yy < H) v = __ldg(reinterpret_cast<const uint32_t*>(obs8 + ((size_t)b * H + yy) * W * C) + wd);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int e = 4 * wd + q, x = e / C, c = e % C;
          rows[(c * NR + r) * RP + XP + x] = __float2bfloat16_rn(__fdiv_rn((float)((v >> (8 * q)) & 0xffu), 255.0f));
        }
      }
    }
    __syncthreads();
    for (int tile = warp; tile < RB * tiles_x; tile += NW) {
      const int ry = tile / tiles_x, x0 = 16 * (tile % tiles_x);
      if (y0 + ry >= H) continue;
      float acc[CO / 8][4];
#pragma unroll
      for (int nt = 0; nt < CO / 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
      const unsigned short* base = reinterpret_cast<const unsigned short*>(rows) + ry * RP + x0 + g8;
#pragma unroll
      for (int ks = 0; ks < KT; ++ks) {
        // A fragment: a0 = (pixel g8, taps 2t4, 2t4+1), a1 = (pixel g8+8, same), a2 / a3 = taps + 8
        uint32_t af[4];
        af[0] = (uint32_t)base[koff[ks][0]] | ((uint32_t)base[koff[ks][1]] << 16);
        af[1] = (uint32_t)base[koff[ks][0] + 8] | ((uint32_t)base[koff[ks][1] + 8] << 16);
        af[2] = (uint32_t)base[koff[ks][2]] | ((uint32_t)base[koff[ks][3]] << 16);
        af[3] = (uint32_t)base[koff[ks][2] + 8] | ((uint32_t)base[koff[ks][3] + 8] << 16);
#pragma unroll
        for (int nt = 0; nt < CO / 8; ++nt) {
          const uint2 bw = bfrag[(ks * (CO / 8) + nt) * 32 + lane];
          mma_bf16_16816(acc[nt], af, bw.x, bw.y);
        }
      }
      // ---- bias + ReLU + bf16, through shared memory so that every pixel leaves as one 128-byte row
      __nv_bfloat16* st = stage + warp * 16 * (CO + 8);
#pragma unroll
      for (int nt = 0; nt < CO / 8; ++nt) {
        const int n = 8 * nt + 2 * t4;
        *reinterpret_cast<uint32_t*>(st + g8 * (CO + 8) + n) =
            pack_bf16x2(fmaxf(acc[nt][0] + bv[nt][0], 0.f), fmaxf(acc[nt][1] + bv[nt][1], 0.f));
        *reinterpret_cast<uint32_t*>(st + (g8 + 8) * (CO + 8) + n) =
            pack_bf16x2(fmaxf(acc[nt][2] + bv[nt][0], 0.f), fmaxf(acc[nt][3] + bv[nt][1], 0.f));
      }
      __syncwarp();
      // 16 pixels x 128 bytes, contiguous; `padded`: the layout of conv_tc.cu (two zero rows between images, two zero
      // columns either side of a row -- the zeros are the caller's, only real pixels are written)
      __nv_bfloat16* dst = padded ? out + (((size_t)2 + (size_t)b * (H + 2) + y0 + ry) * (W + 4) + 2 + x0) * CO
                                  : out + (((size_t)b * H + y0 + ry) * W + x0) * CO;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int chunk = lane + 32 * i;         // 16-byte chunk of the 2 KB tile
        const int px = chunk >> 3, c8 = chunk & 7;
        *reinterpret_cast<uint4*>(dst + px * CO + 8 * c8) = *reinterpret_cast<const uint4*>(st + px * (CO + 8) + 8 * c8);
      }
      __syncwarp();
    }
  }
}

}  // namespace conv1
}  // namespace ocrl

using namespace ocrl;

static int conv_first_launch(const void* obs, const float* weight, const float* bias, void* out, int B, int C, int H, int W,
                             int CO, int padded, void* stream, bool u8 = false) {
  if (!obs || !weight || !bias || !out || (reinterpret_cast<uintptr_t>(out) & 15u)) {
    set_error("conv_first: null or unaligned pointer");
    return OCRL_E_ALIGN;
  }
  if (C != 3 || CO != conv1::CO || W % 16 != 0 || W > 512 || H < 1) {
    set_error("conv_first: C=%d CO=%d H=%d W=%d not supported (C = 3, CO = 64, W a multiple of 16 up to 512)", C, CO, H, W);
    return OCRL_E_SHAPE;
  }
  if (B <= 0) return OCRL_OK;
  const int RP = W + 2 * conv1::XP;
  constexpr int NR = conv1::RB + conv1::KS - 1;
  const size_t smem = sizeof(uint2) * conv1::KT * (conv1::CO / 8) * 32 +
                      sizeof(__nv_bfloat16) * ((size_t)C * NR * RP + conv1::NW * 16 * (conv1::CO + 8));
  auto kern = u8 ? conv1::conv_first_kernel<3, true> : conv1::conv_first_kernel<3, false>;
  if (smem > 48 * 1024) OCRL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int nblocks = B * ((H + conv1::RB - 1) / conv1::RB);
  const int grid = nblocks < 148 * 2 ? nblocks : 148 * 2;  // one wave at 2 CTAs (16 warps) per SM; the weight fragments are built once per CTA
  kern<<<grid, conv1::NW * 32, smem, (cudaStream_t)stream>>>(obs, weight, bias, reinterpret_cast<__nv_bfloat16*>(out), B, H, W,
                                                              padded);
  ocrl::count_launch();
  OCRL_CHECK_CUDA(cudaGetLastError());
  return OCRL_OK;
}

extern "C" int ocrl_conv_first_relu_bf16(const float* obs, const float* weight, const float* bias, void* out, int B, int C,
                                         int H, int W, int CO, void* stream) {
  return conv_first_launch(obs, weight, bias, out, B, C, H, W, CO, 0, stream);
}
extern "C" int ocrl_conv_first_relu_bf16p(const float* obs, const float* weight, const float* bias, void* out_padded, int B,
                                          int C, int H, int W, int CO, void* stream) {
  return conv_first_launch(obs, weight, bias, out_padded, B, C, H, W, CO, 1, stream);
}
extern "C" int ocrl_conv_first_relu_u8p(const unsigned char* frames_hwc, const float* weight, const float* bias,
                                        void* out_padded, int B, int C, int H, int W, int CO, void* stream) {
  return conv_first_launch(frames_hwc, weight, bias, out_padded, B, C, H, W, CO, 1, stream, true);
}

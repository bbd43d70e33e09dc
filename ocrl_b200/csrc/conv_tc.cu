// 5x5 / stride 1 / zero-padding 2 convolution, 64 -> 64 channels, as an implicit GEMM on the 5th-generation tensor
// cores: layers 2-4 of SlotAttnCNNEncoder (ocrs/common/models.py:96-107; Conv2dBlock = conv + bias + ReLU,
// ocrs/common/networks.py:38-53).  bf16 operands, fp32 accumulation in tensor memory -- the arithmetic of the module
// under bf16 autocast.
//
// Activations live in a PADDED channels-last layout: a flat array of positions [(2 + B (H + 2)) rows][W + 4][64] bf16
// with zeros in the two rows between images and in the two columns either side of a row, so that
//   * the 25 taps of an output position p are the positions p + dy (W + 4) + dx: every tap of a 128-position tile is the
//     same flat window shifted by a constant, and the zero padding of the convolution is simply stored;
//   * a tile's A operand for a tap is ONE K-major SWIZZLE_128B shared-memory descriptor whose start address is the
//     slab base advanced by whole 128-byte rows (the swizzle follows absolute address bits, so row-shifted starts read
//     what TMA wrote -- checked with scripts/umma_probe.cu): no im2col copies, the input slab is loaded once per unit.
// One persistent CTA per SM walks a contiguous range of 128-position tiles in units of G tiles:
//   warp 0: TMA producer -- the unit's input slab (128 G + 4 (W + 4) + 4 positions, double-buffered) and the 25 weight
//           taps ([64 out][64 in] bf16, 8 KB each, through a ring; every unit re-streams them from L2);
//   warp 1: tcgen05.mma issuer -- per tap and tile four K = 16 steps of M = 128 (positions) x N = 64 (output channels),
//           accumulators of the unit's G tiles in tensor memory (double-buffered: the epilogue of unit u overlaps the
//           MMAs of unit u + 1);
//   warps 2-5: epilogue -- tcgen05.ld of the thread's position, + bias, ReLU, zero at padding positions, bf16, swizzled
//           staging tile, TMA store of 16 KB contiguous positions.
//
// PAIRED TAPS (PAIR = 1).  With N = 64 output channels an M128 N64 K16 instruction reads 4 KB of positions and 2 KB of
// weights from shared memory for 32 cycles of tensor work: the kernel is bound by the 128 B/clk shared-memory port.  Two
// taps that sit side by side in a filter row, (dy, dx) and (dy, dx + 1), read windows one position apart -- so ONE
// instruction with N = 128 (the two taps' weights stacked: adjacent 8 KB blocks of the packed array) on the window of the
// first tap yields, in columns 0-63, tap A's contribution to output n and, in columns 64-127, tap B's contribution to
// output n - 1 (row n of the window is input n + s_A = (n - 1) + s_B).  The epilogue adds the two halves one row apart:
//   out[n] = acc0[n] + acc1[n + 1]
// (a shuffle; the row that crosses a warp goes through shared memory), and tiles advance by 127 positions so that row
// 127's missing neighbour is never needed.  Per tile 10 pair instructions-groups (8 KB of operands per MMA for twice the
// work) + 5 single taps (dx = 4) replace 25: 60 instead of 100 MMAs, 440 KB instead of 600 KB of operand reads.
#include <cuda.h>

#include "umma_common.cuh"

namespace ocrl {
namespace convtc {

static long long* g_conv_trace = nullptr;  // development aid (ocrl_dev_conv_trace)

constexpr int C = 64;          // channels in and out
constexpr int TAP_BYTES = C * C * 2;
constexpr int NT = 192;

template <int WP, int G, int NWST, int PAIR = 0>
struct Cfg {
  static constexpr int TS = PAIR ? 127 : 128;                 // output positions per tile (tile stride)
  static constexpr int ACCW = PAIR ? 2 * C : C;               // accumulator columns of a tile
  static constexpr int ST_BYTES = PAIR ? 2 * TAP_BYTES : TAP_BYTES;  // one weight stage (a pair of taps / a tap)
  static constexpr int NGRP = PAIR ? 15 : 25;                 // weight groups per unit
  static constexpr int NEED = 128 + TS * (G - 1) + 4 * WP + 4;  // positions a unit's taps touch
  static constexpr int BOX = ((NEED + 23) / 24) * 8;          // rows per TMA box (three boxes per slab, 1024-byte multiples)
  static constexpr int SLAB_ROWS = 3 * BOX;
  static constexpr int SLAB_BYTES = SLAB_ROWS * 128;
  static constexpr int OFF_SLAB = 0;
  static constexpr int OFF_W = OFF_SLAB + 2 * SLAB_BYTES;
  static constexpr int OFF_STG = OFF_W + NWST * ST_BYTES;     // epilogue staging tile [128][128 B]
  static constexpr int OFF_XROW = OFF_STG + 128 * 128;        // paired taps: row 0 of every epilogue warp's second half [4][64] fp32
  static constexpr int OFF_BIAS = OFF_XROW + (PAIR ? 4 * C * 4 : 0);
  static constexpr int OFF_BAR = OFF_BIAS + C * 4;            // slab_full[2] slab_empty[2] w_full[NWST] w_empty[NWST] acc_full[2] acc_empty[2]
  static constexpr int NBAR = 8 + 2 * NWST;
  static constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
  static constexpr int SMEM_BYTES = OFF_TMEM + 16 + 1024;     // + slack for the 1024-byte alignment of the base
  static constexpr uint32_t ACC_COLS = G * ACCW;              // accumulator columns of one unit
  static_assert(BOX <= 256 && 2 * ACC_COLS <= 512, "box rows / tensor memory");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

struct Params {
  const float* bias;   // [64] or null
  int relu;
  int B, H, W;
  long long p_first;   // flat position of the first tile
  long long n_pos;     // positions of the whole array (rows of the tensor maps)
  int n_tiles;
  long long* trace;    // development aid: clock64 stamps of CTA 0 (null in production)
};

template <int WP, int G, int NWST, int PAIR>
__global__ void __launch_bounds__(NT, 1)
conv5x5_tc_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_w,
                  const __grid_constant__ CUtensorMap tm_w2, const __grid_constant__ CUtensorMap tm_out, const Params p) {
  using CF = Cfg<WP, G, NWST, PAIR>;
  constexpr int TS = CF::TS;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* slab = sm + CF::OFF_SLAB;
  unsigned char* wst = sm + CF::OFF_W;
  unsigned char* stg = sm + CF::OFF_STG;
  float* s_bias = reinterpret_cast<float*>(sm + CF::OFF_BIAS);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + CF::OFF_BAR);
  uint64_t* slab_full = bars;
  uint64_t* slab_empty = bars + 2;
  uint64_t* w_full = bars + 4;
  uint64_t* w_empty = w_full + NWST;
  uint64_t* acc_full = w_empty + NWST;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + CF::OFF_TMEM);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (p.trace != nullptr && blockIdx.x == 0 && tid == 0) p.trace[0] = clock64();
  // every CTA: start / end on the global timer (ns), its SM (trace[128 + 4 b ..]): spread of the persistent CTAs
  if (p.trace != nullptr && tid == 0) {
    unsigned long long t;
    unsigned smid;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
    p.trace[128 + 4 * blockIdx.x] = (long long)t;
    p.trace[128 + 4 * blockIdx.x + 2] = (long long)smid;
  }

  if (tid < C) s_bias[tid] = p.bias ? __ldg(p.bias + tid) : 0.f;
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&slab_full[i], 1);
      mbar_init(&slab_empty[i], 1);
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 4);
    }
    for (int i = 0; i < NWST; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    mbar_fence_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_in) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_out) : "memory");
  }
  if (warp == 1) tc::tmem_alloc<512>(tmem_slot);
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tmem = *tmem_slot;

  // contiguous tile range of this CTA, walked in units of G tiles
  const int t_begin = (int)((long long)blockIdx.x * p.n_tiles / gridDim.x);
  const int t_end = (int)((long long)(blockIdx.x + 1) * p.n_tiles / gridDim.x);
  const int n_units = (t_end - t_begin + G - 1) / G;

  if (warp == 0) {
    // ================================================================ TMA producer (whole warp, one elected lane issues)
    const bool leader = tc::elect_one();
    const uint64_t pol_keep = [] { uint64_t q; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(q)); return q; }();
    auto load_slab = [&](int u) {  // slab of unit u into buffer u & 1
      const int sb = u & 1;
      if (u >= 2) mbar_wait(&slab_empty[sb], (uint32_t)(((u >> 1) - 1) & 1));
      const long long s0 = p.p_first + (long long)TS * (t_begin + u * G) - 2 * WP - 2;  // may be negative: rows outside read as zero
      if (leader) {
        mbar_expect_tx(&slab_full[sb], (uint32_t)CF::SLAB_BYTES);
#pragma unroll
        for (int j = 0; j < 3; ++j)
          tc::tma_load_2d(slab + sb * CF::SLAB_BYTES + j * CF::BOX * 128, &tm_in, 0, (int)(s0 + j * CF::BOX), &slab_full[sb]);
      }
      __syncwarp();
    };
    int wc = 0;
    if (n_units > 0) load_slab(0);
    for (int u = 0; u < n_units; ++u) {
      for (int grp = 0; grp < CF::NGRP; ++grp) {
        if (grp == (PAIR ? 4 : 6) && u + 1 < n_units) load_slab(u + 1);
        const int st = wc % NWST;
        if (wc >= NWST) mbar_wait(&w_empty[st], (uint32_t)(((wc / NWST) - 1) & 1));
        if (leader) {
          if (PAIR) {  // group = (filter row, 0..2): taps dx = 0|1, 2|3 (two adjacent blocks of the packed array: one box), 4
            const int tap0 = (grp / 3) * 5 + 2 * (grp % 3);
            const bool two = (grp % 3) < 2;
            mbar_expect_tx(&w_full[st], (uint32_t)(two ? 2 * TAP_BYTES : TAP_BYTES));
            tc::tma_load_2d_hint(wst + st * CF::ST_BYTES, two ? &tm_w2 : &tm_w, 0, tap0 * C, &w_full[st], pol_keep);
          } else {
            mbar_expect_tx(&w_full[st], (uint32_t)TAP_BYTES);
            tc::tma_load_2d_hint(wst + st * CF::ST_BYTES, &tm_w, 0, grp * C, &w_full[st], pol_keep);
          }
        }
        __syncwarp();
        ++wc;
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    const bool leader = tc::elect_one();
    constexpr uint32_t IDESC = tc::idesc_bf16(128, C), IDESC2 = tc::idesc_bf16(128, 2 * C);
    int wc = 0;
    const bool tr = (p.trace != nullptr && blockIdx.x == 0 && lane == 0);
    for (int u = 0; u < n_units; ++u) {
      const int sb = u & 1;
      const int g = min(G, t_end - (t_begin + u * G));
      if (tr && u < 8) p.trace[16 + u * 8 + 0] = clock64();
      mbar_wait(&slab_full[sb], (uint32_t)((u >> 1) & 1));
      if (tr && u < 8) p.trace[16 + u * 8 + 1] = clock64();
      if (u >= 2) mbar_wait(&acc_empty[sb], (uint32_t)(((u >> 1) - 1) & 1));
      if (tr && u < 8) p.trace[16 + u * 8 + 2] = clock64();
      long long wsum = 0;
      tc::fence_after();
      const uint32_t sa = smem_u32(slab + sb * CF::SLAB_BYTES);
      const uint32_t acc = tmem + (uint32_t)sb * CF::ACC_COLS;
      for (int grp = 0; grp < CF::NGRP; ++grp) {
        const int st = wc % NWST;
        const long long tw0 = tr ? clock64() : 0;
        mbar_wait(&w_full[st], (uint32_t)((wc / NWST) & 1));
        if (tr) wsum += clock64() - tw0;
        tc::fence_after();
        const uint32_t wa = smem_u32(wst + st * CF::ST_BYTES);
        // window of the group's first tap; a pair's second tap lands one row lower in columns 64-127
        const int tap = PAIR ? (grp / 3) * 5 + 2 * (grp % 3) : grp;
        const uint32_t rowoff = (uint32_t)((tap / 5) * WP + tap % 5) * 128u;
        const uint32_t idesc = (PAIR && (grp % 3) < 2) ? IDESC2 : IDESC;
        if (leader) {
          for (int i = 0; i < g; ++i) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              tc::mma_bf16(acc + (uint32_t)(i * CF::ACCW), tc::smem_desc(sa + (uint32_t)(i * TS * 128) + rowoff + ks * 32, 16, 1024, tc::SW_128),
                           tc::smem_desc(wa + ks * 32, 16, 1024, tc::SW_128), idesc, (uint32_t)((grp | ks) != 0));
          }
          tc::commit(&w_empty[st]);
        }
        __syncwarp();
        ++wc;
      }
      if (leader) {
        tc::commit(&acc_full[sb]);
        tc::commit(&slab_empty[sb]);
      }
      __syncwarp();
      if (tr && u < 8) { p.trace[16 + u * 8 + 3] = clock64(); p.trace[16 + u * 8 + 4] = wsum; }
    }
  } else {
    // ================================================================ epilogue: one output position per thread
    const int q = warp & 3, row = q * 32 + lane;
    const bool e0 = (warp == 2 && lane == 0);
    const int RI = p.H + 2;
    for (int u = 0; u < n_units; ++u) {
      const int ab = u & 1;
      const int g = min(G, t_end - (t_begin + u * G));
      const bool tre = (p.trace != nullptr && blockIdx.x == 0 && e0);
      if (tre && u < 8) p.trace[16 + u * 8 + 5] = clock64();
      mbar_wait(&acc_full[ab], (uint32_t)((u >> 1) & 1));
      if (tre && u < 8) p.trace[16 + u * 8 + 6] = clock64();
      tc::fence_after();
      for (int i = 0; i < g; ++i) {
        uint32_t r0[32], r1[32];
        const uint32_t tcol = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)ab * CF::ACC_COLS + (uint32_t)(i * CF::ACCW);
        tc::tmem_ld32_nowait(tcol, r0);
        tc::tmem_ld32_nowait(tcol + 32, r1);
        if (PAIR) {
          // second half (columns 64-127): the contribution of the pairs' second taps to the output ONE ROW UP.
          // out[n] = acc0[n] + acc1[n + 1]: a shuffle inside the warp, shared memory across the warp boundary
          uint32_t h0[32], h1[32];
          tc::tmem_ld32_nowait(tcol + 64, h0);
          tc::tmem_ld32_nowait(tcol + 96, h1);
          tc::tmem_ld_wait();
          float* xrow = reinterpret_cast<float*>(sm + CF::OFF_XROW);
          if (lane == 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j) { xrow[q * C + j] = __uint_as_float(h0[j]); xrow[q * C + 32 + j] = __uint_as_float(h1[j]); }
          }
          asm volatile("bar.sync 2, 128;" ::: "memory");
          const float* xn = xrow + ((q + 1) & 3) * C;  // (warp 3's lane 31 is row 127: never stored)
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float a = __shfl_down_sync(0xffffffffu, __uint_as_float(h0[j]), 1);
            float b = __shfl_down_sync(0xffffffffu, __uint_as_float(h1[j]), 1);
            if (lane == 31) { a = xn[j]; b = xn[32 + j]; }
            r0[j] = __float_as_uint(__uint_as_float(r0[j]) + a);
            r1[j] = __float_as_uint(__uint_as_float(r1[j]) + b);
          }
        } else {
          tc::tmem_ld_wait();
        }
        if (i == g - 1) {  // the unit's accumulators are in registers: the issuer may reuse the buffer
          tc::fence_before();
          __syncwarp();
          if (lane == 0) tc::arrive(&acc_empty[ab]);
        }
        const long long pos = p.p_first + (long long)TS * (t_begin + u * G + i) + row;
        const int grow = (int)(pos / WP), x = (int)(pos - (long long)grow * WP);
        const bool valid = grow >= 2 && grow < 2 + p.B * RI && ((grow - 2) % RI) < p.H && x >= 2 && x < p.W + 2;  // else padding
        uint4 o[8];
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
          uint32_t w4[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int ch = 8 * c8 + 2 * e;
            float a0 = __uint_as_float(ch < 32 ? r0[ch] : r1[ch - 32]) + s_bias[ch];
            float a1 = __uint_as_float(ch + 1 < 32 ? r0[ch + 1] : r1[ch + 1 - 32]) + s_bias[ch + 1];
            if (p.relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); }
            w4[e] = valid ? pack_bf16x2(a0, a1) : 0u;
          }
          o[c8] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
        }
        // the previous TMA store has read the staging tile (thread e0 waited for it before this barrier)
        asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) *reinterpret_cast<uint4*>(stg + row * 128 + ((c8 ^ (row & 7)) << 4)) = o[c8];
        fence_proxy_async();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (e0) {
          tc::tma_store_2d(&tm_out, stg, 0, (int)(p.p_first + (long long)TS * (t_begin + u * G + i)));
          tc::tma_store_commit();
          tc::tma_store_wait_read<0>();
        }
      }
      if (tre && u < 8) p.trace[16 + u * 8 + 7] = clock64();
    }
    if (e0) tc::tma_store_wait_all<0>();
    if (p.trace != nullptr && blockIdx.x == 0 && e0) p.trace[1] = clock64();
  }
  tc::fence_before();
  __syncthreads();
  if (p.trace != nullptr && tid == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    p.trace[128 + 4 * blockIdx.x + 1] = (long long)t;
    p.trace[128 + 4 * blockIdx.x + 3] = (long long)(t_end - t_begin);
  }
  if (warp == 1) tc::tmem_dealloc<512>(tmem);
}

// fp32 [64 out][64 in][5][5] -> bf16 [tap = dy * 5 + dx][out][in]: the K-major B operand of every tap
__global__ void conv5x5_pack_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 25 * C * C) return;
  const int ci = i % C, co = (i / C) % C, tap = i / (C * C);
  out[i] = __float2bfloat16_rn(__ldg(w + ((size_t)(co * C + ci) * 25) + tap));
}

template <int WP, int G, int NWST, int PAIR = 0>
static int launch(const void* in, const void* wpk, const float* bias, void* out, int B, int H, int W, int relu, cudaStream_t stream) {
  using CF = Cfg<WP, G, NWST, PAIR>;
  Params p;
  p.bias = bias; p.relu = relu; p.B = B; p.H = H; p.W = W;
  // the tiles cover EVERY position of the output array (padding positions are written as zeros), so the caller never has
  // to initialise an output buffer
  p.p_first = 0;
  p.n_pos = (2LL + (long long)B * (H + 2)) * WP;
  p.n_tiles = (int)((p.n_pos + CF::TS - 1) / CF::TS);
  p.trace = g_conv_trace;
  CUtensorMap tm_in, tm_w, tm_w2, tm_out;
  if (!tc::make_map_bf16_sw128(&tm_in, in, C, (uint64_t)p.n_pos, C * 2, CF::BOX) ||
      !tc::make_map_bf16_sw128(&tm_w, wpk, C, 25 * C, C * 2, C) ||
      !tc::make_map_bf16_sw128(&tm_w2, wpk, C, 25 * C, C * 2, 2 * C) ||
      !tc::make_map_bf16_sw128(&tm_out, out, C, (uint64_t)p.n_pos, C * 2, CF::TS)) {
    set_error("conv5x5_tc: cuTensorMapEncodeTiled failed");
    return OCRL_E_LAUNCH;
  }
  auto kern = conv5x5_tc_kernel<WP, G, NWST, PAIR>;
  static bool configured = false;
  if (!configured) {
    OCRL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::SMEM_BYTES));
    configured = true;
  }
  int dev = 0, sms = 148;
  OCRL_CHECK_CUDA(cudaGetDevice(&dev));
  OCRL_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int units = (p.n_tiles + G - 1) / G;
  const int grid = units < sms ? units : sms;
  kern<<<grid, NT, CF::SMEM_BYTES, stream>>>(tm_in, tm_w, tm_w2, tm_out, p);
  ocrl::count_launch();
  OCRL_CHECK_CUDA(cudaGetLastError());
  return OCRL_OK;
}

}  // namespace convtc
}  // namespace ocrl

using namespace ocrl;

extern "C" void ocrl_dev_conv_trace(long long* p) { convtc::g_conv_trace = p; }
static int g_conv_variant = 0;
extern "C" void ocrl_dev_conv_variant(int v) { g_conv_variant = v; }  // development knob (scripts/quick_conv.py), not in the header

extern "C" size_t ocrl_conv_padded_bytes(int B, int H, int W) {
  if (B < 0 || H < 1 || W < 1) return 0;
  return (size_t)(2 + (size_t)B * (H + 2)) * (W + 4) * convtc::C * 2;
}

extern "C" int ocrl_conv5x5_pack_weights(const float* weight, void* packed, int CO, int CI, void* stream) {
  if (!weight || !packed || (reinterpret_cast<uintptr_t>(packed) & 127u)) {
    set_error("conv5x5_pack_weights: null or unaligned pointer (packed needs 128-byte alignment)");
    return OCRL_E_ALIGN;
  }
  if (CO != convtc::C || CI != convtc::C) {
    set_error("conv5x5_pack_weights: %d -> %d channels not supported (64 -> 64)", CI, CO);
    return OCRL_E_SHAPE;
  }
  convtc::conv5x5_pack_kernel<<<(25 * convtc::C * convtc::C + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      weight, reinterpret_cast<__nv_bfloat16*>(packed));
  ocrl::count_launch();
  OCRL_CHECK_CUDA(cudaGetLastError());
  return OCRL_OK;
}

extern "C" int ocrl_conv5x5_c64_tc(const void* in_padded, const void* packed_w, const float* bias, void* out_padded, int B,
                                   int H, int W, int relu, void* stream) {
  if (!in_padded || !packed_w || !out_padded || (reinterpret_cast<uintptr_t>(in_padded) & 127u) ||
      (reinterpret_cast<uintptr_t>(out_padded) & 127u) || (reinterpret_cast<uintptr_t>(packed_w) & 127u)) {
    set_error("conv5x5_c64_tc: null or unaligned pointer (128-byte alignment)");
    return OCRL_E_ALIGN;
  }
  if (in_padded == out_padded) {
    set_error("conv5x5_c64_tc: in-place convolution is not possible");
    return OCRL_E_SHAPE;
  }
  if (B <= 0) return OCRL_OK;
  if (H < 1) {
    set_error("conv5x5_c64_tc: H=%d", H);
    return OCRL_E_SHAPE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  // large batches: paired taps (PAIR = 1, units of two 127-position tiles); small ones (rollout: 4 frames = 141 tiles, one per
  // CTA): a CTA's 25 weight taps are a latency chain through the ring, so units of one tile with a 12-stage ring.
  // Measured per layer at B = 64, 64 x 64 (us): three tiles / four stages 75.5, paired taps 64.1, cuDNN 60.7
  switch (W) {
    case 32: {
      const long long n_tiles = ((2LL + (long long)B * (H + 2)) * (W + 4) + 127) / 128;
      if (g_conv_variant == 0 && n_tiles > 2 * 148) return convtc::launch<36, 2, 4, 1>(in_padded, packed_w, bias, out_padded, B, H, W, relu, s);
      return convtc::launch<36, 3, 4>(in_padded, packed_w, bias, out_padded, B, H, W, relu, s);
    }
    case 64: {
      const long long n_tiles = ((2LL + (long long)B * (H + 2)) * (W + 4) + 127) / 128;
      if (g_conv_variant == 0 && n_tiles <= 2 * 148) return convtc::launch<68, 1, 12>(in_padded, packed_w, bias, out_padded, B, H, W, relu, s);
      if (g_conv_variant == 1) return convtc::launch<68, 3, 4>(in_padded, packed_w, bias, out_padded, B, H, W, relu, s);
      if (g_conv_variant == 2) return convtc::launch<68, 2, 8>(in_padded, packed_w, bias, out_padded, B, H, W, relu, s);
      if (g_conv_variant == 4) return convtc::launch<68, 2, 3, 1>(in_padded, packed_w, bias, out_padded, B, H, W, relu, s);
      return convtc::launch<68, 2, 4, 1>(in_padded, packed_w, bias, out_padded, B, H, W, relu, s);
    }
    case 128: return convtc::launch<132, 1, 4>(in_padded, packed_w, bias, out_padded, B, H, W, relu, s);
    default:
      set_error("conv5x5_c64_tc: W=%d not instantiated (32, 64, 128)", W);
      return OCRL_E_SHAPE;
  }
}

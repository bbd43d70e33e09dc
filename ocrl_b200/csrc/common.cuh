// Shared device helpers for the sm_100a slot-attention kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cooperative_groups.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ocrl_sa.h"

namespace cg = cooperative_groups;

namespace ocrl {

void set_error(const char* fmt, ...);

#define OCRL_CHECK_CUDA(expr)                                                         \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      ocrl::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return OCRL_E_LAUNCH;                                                           \
    }                                                                                 \
  } while (0)

constexpr unsigned FULL = 0xffffffffu;

// ---------------------------------------------------------------------------------------------
// mbarrier + 1-D bulk async copy (TMA engine, no tensor map needed for contiguous rows)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared::cta bulk copy; bytes must be a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// element access: two consecutive features as float2 from fp32 or bf16 storage
// ---------------------------------------------------------------------------------------------
template <typename T>
struct Elem;
template <>
struct Elem<float> {
  static constexpr int kDtype = OCRL_DT_F32;
  __device__ static __forceinline__ float2 load2(const float* p) { return *reinterpret_cast<const float2*>(p); }
  __device__ static __forceinline__ void store2(float* p, float a, float b) {
    *reinterpret_cast<float2*>(p) = make_float2(a, b);
  }
};
template <>
struct Elem<__nv_bfloat16> {
  static constexpr int kDtype = OCRL_DT_BF16;
  __device__ static __forceinline__ float2 load2(const __nv_bfloat16* p) {
    uint32_t u = *reinterpret_cast<const uint32_t*>(p);
    return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
  }
  __device__ static __forceinline__ void store2(__nv_bfloat16* p, float a, float b) {
    *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
  }
};

// ---------------------------------------------------------------------------------------------
// Transposed warp reduction.  Every lane holds NV partial sums v[0..NV); afterwards the sums over
// all 32 lanes are spread across the lanes: lane `l` holds `count` complete sums, the i-th being
// original index `base + i`.  Costs ~NV shuffles instead of 5*NV.  Deterministic.
// ---------------------------------------------------------------------------------------------
template <int NV, int OFF>
struct XReduce {
  static constexpr bool kHalve = (NV % 2 == 0);
  static constexpr int kNext = kHalve ? NV / 2 : NV;
  using Next = XReduce<kNext, OFF / 2>;
  static constexpr int kFinal = Next::kFinal;
  __device__ static __forceinline__ void run(float* v, int lane, int& base) {
    if constexpr (kHalve) {
      constexpr int h = NV / 2;
      const bool upper = (lane & OFF) != 0;
#pragma unroll
      for (int i = 0; i < h; ++i) {
        float keep = upper ? v[h + i] : v[i];
        float send = upper ? v[i] : v[h + i];
        v[i] = keep + __shfl_xor_sync(FULL, send, OFF);
      }
      base += upper ? h : 0;
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] += __shfl_xor_sync(FULL, v[i], OFF);
    }
    Next::run(v, lane, base);
  }
  // lanes for which this returns true hold the canonical copy (duplicates exist after odd steps)
  __device__ static __forceinline__ bool primary(int lane) {
    return (kHalve || (lane & OFF) == 0) && Next::primary(lane);
  }
};
template <int NV>
struct XReduce<NV, 0> {
  static constexpr int kFinal = NV;
  __device__ static __forceinline__ void run(float*, int, int&) {}
  __device__ static __forceinline__ bool primary(int) { return true; }
};

template <int NV>
__device__ __forceinline__ void xreduce(float (&v)[NV], int lane, int& base) {
  base = 0;
  XReduce<NV, 16>::run(v, lane, base);
}

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
  return x;
}

}  // namespace ocrl

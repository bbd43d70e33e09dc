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
void count_launch();  // one call per kernel launch of this library (ocrl_launch_count)

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

// ---------------------------------------------------------------------------------------------
// cp.async (16-byte, L2-only), ldmatrix and bf16 mma.sync wrappers
// ---------------------------------------------------------------------------------------------
// copies 16 bytes global -> shared; src_bytes = 0 zero-fills the destination instead
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src_gmem, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(smem_row)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(smem_row)));
}
__device__ __forceinline__ void ldmatrix_x2(uint32_t (&r)[2], const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(smem_u32(smem_row)));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t (&r)[2], const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];"
               : "=r"(r[0]), "=r"(r[1])
               : "r"(smem_u32(smem_row)));
}
// D(16x8, f32) += A(16x8, bf16, row) * B(8x8, bf16, col)
__device__ __forceinline__ void mma_bf16_1688(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(b0));
}
// D(16x8, f32) += A(16x16, bf16, row) * B(16x8, bf16, col)
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
  return x;
}

}  // namespace ocrl

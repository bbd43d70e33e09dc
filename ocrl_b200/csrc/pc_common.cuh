// Device / host helpers shared by the persistent-cluster iteration kernels (sa_iter_fwd_pc.cu, sa_iter_fwd_pipe.cu):
// cluster exchange primitives (st.async + mbarrier), single-box swizzled TMA tiles, the tensor-core
// matrix-vector job of the slot update.
#pragma once
#include <cuda.h>
#include <stdlib.h>

#include "slot_math.cuh"

namespace ocrl {
namespace pc {

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Wait for an exchange round.  The data behind the barrier was written into THIS CTA's shared memory by peers'
// st.async; its visibility is tied to the complete_tx the waiter observes, so the default CTA-scope acquire is
// enough (a cluster-scope acquire makes ptxas emit an L1 invalidate, CCTL.IVALL, after every wait: ~10 % of the
// kernel's stall samples in the first profile).
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAITC_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONEC_%=;\n"
      "bra WAITC_%=;\n"
      "DONEC_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_async_v4(uint32_t raddr, float4 v, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(raddr),
               "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)),
               "r"(__float_as_uint(v.w)), "r"(rbar)
               : "memory");
}
__device__ __forceinline__ void st_async_v2(uint32_t raddr, uint32_t x, uint32_t y, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];" ::"r"(raddr),
               "r"(x), "r"(y), "r"(rbar)
               : "memory");
}
__device__ __forceinline__ void st_async_b32(uint32_t raddr, uint32_t x, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(raddr), "r"(x),
               "r"(rbar)
               : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_u32(dst)),
      "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}
// same, with an L2 eviction-priority hint (createpolicy): evict_last for tiles that are read again by the next
// iteration, evict_first for the final pass
__device__ __forceinline__ void tma_load_3d_hint(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;" ::"r"(
          smem_u32(dst)),
      "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// Byte offset of (row, dbyte) inside a [TOK][D] bf16 tile that the TMA unit wrote as [row][D/64][64] with the
// 128-byte swizzle (16-byte chunk index xor-ed with the 128-byte line index mod 8); dbyte is a multiple of 16.
template <int D>
__device__ __forceinline__ int swz_off(int row, int dbyte) {
  const int line = row * (D / 64) + (dbyte >> 7);
  return line * 128 + ((((dbyte >> 4) & 7) ^ (line & 7)) << 4);
}
__device__ __forceinline__ int lds_volatile(const int* p) {
  int v;
  asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void sts_volatile(int* p, int v) {
  asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void group_sync(int grp) {
  asm volatile("bar.sync %0, 256;" ::"r"(grp + 1) : "memory");
}
__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t x) {
  uint32_t r;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(r) : "r"(x));
  return r;
}
__device__ __forceinline__ float ex2f(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void split_hilo(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
  const __nv_bfloat162 l = __floats2bfloat162_rn(x - __low2float(h), y - __high2float(h));
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// One MMA job: rows [16*mt, 16*mt+16) of the weight slice `W` (bf16, `wp` bytes per row, `R` valid rows;
// rows past R read the zero row) times the activations (bf16 hi/lo, [8][ap bytes]) over k-steps
// [ks0, ks1) -> P[(16*mt + row) * 8 + slot].
template <bool HILO = true>
__device__ __forceinline__ void mma_job(const unsigned char* W, int wp, int R, const unsigned char* zrow, int mt,
                                        int ks0, int ks1, const unsigned char* act_hi, const unsigned char* act_lo,
                                        int ap, float* P, int lane) {
  const int g8 = lane >> 2, t4 = lane & 3;
  const int row = 16 * mt + (lane & 7) + ((lane >> 3) & 1) * 8;
  const unsigned char* arow = (row < R ? W + (size_t)row * wp : zrow) + (lane >> 4) * 16;
  const unsigned char* bh = act_hi + g8 * ap + 4 * t4;
  const unsigned char* bl = act_lo + g8 * ap + 4 * t4;
  float ch0[4] = {0.f, 0.f, 0.f, 0.f}, ch1[4] = {0.f, 0.f, 0.f, 0.f};  // four independent accumulator chains
  float cl0[4] = {0.f, 0.f, 0.f, 0.f}, cl1[4] = {0.f, 0.f, 0.f, 0.f};
  auto step = [&](int ks, float (&h)[4], float (&l)[4]) {
    uint32_t af[4];
    ldmatrix_x4(af, arow + ks * 32);
    const uint32_t h0 = *reinterpret_cast<const uint32_t*>(bh + ks * 32);
    const uint32_t h1 = *reinterpret_cast<const uint32_t*>(bh + ks * 32 + 16);
    mma_bf16_16816(h, af, h0, h1);
    if constexpr (HILO) {
      const uint32_t l0 = *reinterpret_cast<const uint32_t*>(bl + ks * 32);
      const uint32_t l1 = *reinterpret_cast<const uint32_t*>(bl + ks * 32 + 16);
      mma_bf16_16816(l, af, l0, l1);
    }
  };
  int ks = ks0;
#pragma unroll 2
  for (; ks + 1 < ks1; ks += 2) {
    step(ks, ch0, cl0);
    step(ks + 1, ch1, cl1);
  }
  if (ks < ks1) step(ks, ch0, cl0);
  float* o = P + (16 * mt + g8) * 8 + 2 * t4;
  *reinterpret_cast<float2*>(o) = make_float2((ch0[0] + ch1[0]) + (cl0[0] + cl1[0]), (ch0[1] + ch1[1]) + (cl0[1] + cl1[1]));
  *reinterpret_cast<float2*>(o + 64) = make_float2((ch0[2] + ch1[2]) + (cl0[2] + cl1[2]), (ch0[3] + ch1[3]) + (cl0[3] + cl1[3]));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}
// [rows][D] bf16 viewed as [rows][D/64][64]: one box = `tok` rows, written to shared memory with the 128-byte swizzle
static bool make_kv_map(CUtensorMap* tm, const void* base, long long rows, int D, int tok) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[3] = {64, (cuuint64_t)(D / 64), (cuuint64_t)rows};
  cuuint64_t strides[2] = {128, (cuuint64_t)D * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)(D / 64), (cuuint32_t)tok};
  cuuint32_t estr[3] = {1, 1, 1};
  return fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace pc
}  // namespace ocrl

// Issue / completion timing of straight-line tcgen05.mma sequences (development aid; companion of umma_probe.cu).
// Every descriptor / accumulator offset is a compile-time constant added to a uniform base, so the timed region is
// the UTCHMMA instructions themselves.  Operands are zero-filled shared memory: only timing matters.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}

// NM MMAs; accumulator i % DW at column (i % DW) * DSTEP; A / B start advance (16-byte units) (i % 4) * ASTEP / BSTEP
template <int NM, int DW, int DSTEP, int ASTEP, int BSTEP>
__global__ void __launch_bounds__(128, 1) time_kernel(uint64_t da_rel, uint64_t db_rel, uint32_t idesc, long long* out) {
  extern __shared__ __align__(1024) unsigned char sm_raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(sm_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 196608 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    const uint64_t base = (uint64_t)(smem_u32(sm) >> 4);
    const bool leader = elect_one();
    for (int rep = 0; rep < 3; ++rep) {  // the last repetition is reported (warm instruction cache)
      long long t0 = 0, t1 = 0;
      if (leader) {
        const uint64_t da0 = da_rel + base, db0 = db_rel + base;
        t0 = clock64();
#pragma unroll
        for (int i = 0; i < NM; ++i) {
          const uint64_t da = da0 + (uint64_t)((i & 3) * ASTEP), db = db0 + (uint64_t)((i & 3) * BSTEP);
          const uint32_t dcol = tmem + (uint32_t)((i % DW) * DSTEP);
          if (i < DW)
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 0, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(dcol), "l"(da), "l"(db), "r"(idesc) : "memory");
          else
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(dcol), "l"(da), "l"(db), "r"(idesc) : "memory");
        }
        t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      }
      __syncwarp();
      asm volatile(
          "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(
              smem_u32(&bar)),
          "r"(rep & 1)
          : "memory");
      const long long t2 = clock64();
      if (leader) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

static uint64_t desc(uint32_t start, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((start >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)layout << 61);
}
static uint32_t idesc(int M, int N, int a_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
static long long* d_out;
constexpr uint32_t B_OFF = 131072;
constexpr int SMEM = 196608 + 2048;

template <int NM, int DW, int DSTEP, int ASTEP, int BSTEP>
static void run(const char* name, uint64_t da, uint64_t db, uint32_t id) {
  auto k = time_kernel<NM, DW, DSTEP, ASTEP, BSTEP>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  k<<<1, 128, SMEM>>>(da, db, id, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
  long long h[2];
  cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
  printf("%-58s %2d mma, %2d accumulators: issue %5lld cyc (%5.1f / mma)  complete %5lld cyc (%5.1f / mma)\n", name, NM, DW, h[0],
         (double)h[0] / NM, h[1], (double)h[1] / NM);
}

int main() {
  cudaMalloc(&d_out, 16);
  const uint64_t a_k = desc(0, 16, 1024, 2), a_mn = desc(0, 8192, 1024, 2);
  const uint64_t b8 = desc(B_OFF, 128, 128, 0), b16 = desc(B_OFF, 256, 128, 0), b64 = desc(B_OFF, 16, 1024, 2);
#define SWEEP(NAME, DSTEP, ASTEP, BSTEP, DA, DB, ID)            \
  run<48, 1, DSTEP, ASTEP, BSTEP>(NAME, DA, DB, ID);            \
  run<48, 2, DSTEP, ASTEP, BSTEP>(NAME, DA, DB, ID);            \
  run<48, 4, DSTEP, ASTEP, BSTEP>(NAME, DA, DB, ID);            \
  run<48, 8, DSTEP, ASTEP, BSTEP>(NAME, DA, DB, ID);
  SWEEP("M64  N8  K-major sw128 A", 8, 2, 16, a_k, b8, idesc(64, 8, 0))
  SWEEP("M128 N16 K-major sw128 A", 16, 2, 32, a_k, b16, idesc(128, 16, 0))
  SWEEP("M128 N16 MN-major sw128 A", 16, 128, 32, a_mn, b16, idesc(128, 16, 1))
  SWEEP("M64  N16 MN-major sw128 A", 16, 128, 32, a_mn, b16, idesc(64, 16, 1))
  SWEEP("M128 N64 K-major sw128 A, sw128 B", 64, 2, 2, a_k, b64, idesc(128, 64, 0))
  run<48, 4, 64, 2, 2>("M128 N64, A start shifted by 1 row (128 B)", desc(128, 16, 1024, 2), b64, idesc(128, 64, 0));
  run<48, 4, 64, 2, 2>("M128 N64, A start shifted by 5 rows", desc(640, 16, 1024, 2), b64, idesc(128, 64, 0));
  run<48, 4, 64, 2, 2>("M128 N64, A start shifted by 4 rows", desc(512, 16, 1024, 2), b64, idesc(128, 64, 0));
  run<48, 4, 64, 2, 2>("M128 N64, A start shifted by 68 rows", desc(68 * 128, 16, 1024, 2), b64, idesc(128, 64, 0));
  run<48, 4, 64, 2, 2>("M128 N64, A no swizzle (LBO 16 KB... K-chunk planes)", desc(0, 16384, 128, 0), b64, idesc(128, 64, 0));
  run<48, 2, 128, 2, 2>("M128 N128 K-major sw128 both", a_k, b64, idesc(128, 128, 0));
  run<48, 2, 256, 2, 2>("M128 N256 K-major sw128 both", a_k, b64, idesc(128, 256, 0));
  run<48, 1, 256, 2, 2>("M128 N256 K-major sw128 both", a_k, b64, idesc(128, 256, 0));
  return 0;
}

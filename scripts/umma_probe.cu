// Data-driven tcgen05.mma probe (development aid, not part of the product path).
//
//   umma_probe <case.bin> <out.bin>
//
// A case file holds a raw shared-memory image, a list of MMA operations (shared-memory descriptors whose start
// address field is relative to the image, an instruction descriptor, a tensor-memory offset, the accumulate flag)
// and the number of tensor-memory columns to dump.  The kernel fills the columns with a sentinel, issues the
// operations from one thread, and dumps all 128 lanes x ncols columns.  scripts/umma_cases.py generates the cases
// and decodes the dumps (operand layouts are discovered with address-coded operands against a selector operand).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

struct MmaOp {
  uint64_t da, db;
  uint32_t idesc, dcol, acc, pad;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1)
probe_kernel(const uint8_t* image, int image_bytes, const MmaOp* ops, int nops, float* out, int ncols, float sentinel,
             const uint32_t* timage, int tcols) {
  extern __shared__ __align__(1024) unsigned char sm_raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(sm_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < image_bytes / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = reinterpret_cast<const uint32_t*>(image)[i];
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
  const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
  {
    const uint32_t s = __float_as_uint(sentinel);
    for (int c = 0; c < ncols; c += 8) {
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(trow + c), "r"(s) : "memory");
    }
    // optional A-operand image: this thread's lane, columns [256, 256 + tcols)
    for (int c = 0; c < tcols; ++c) {
      const uint32_t w = timage[(size_t)tid * tcols + c];
      asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(trow + 256 + c), "r"(w) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) {
    const uint64_t base = (uint64_t)(smem_u32(sm) >> 4);
    for (int i = 0; i < nops; ++i) {
      const MmaOp o = ops[i];
      const uint64_t da = o.da + base, db = o.db + base;
      if (o.pad & 0x80000000u)
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem + o.dcol),
            "r"(tmem + (uint32_t)o.da), "l"(db), "r"(o.idesc), "r"(o.acc)
            : "memory");
      else
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem + o.dcol),
            "l"(da), "l"(db), "r"(o.idesc), "r"(o.acc)
            : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(
            smem_u32(&bar)),
        "r"(0)
        : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c = 0; c < ncols; c += 8) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(trow + c));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 8; ++i) out[(size_t)tid * ncols + c + i] = __uint_as_float(r[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e = (x);                                                           \
    if (e != cudaSuccess) {                                                        \
      fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e));                      \
      return 2;                                                                    \
    }                                                                              \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: umma_probe case.bin out.bin\n");
    return 1;
  }
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 1;
  int32_t hdr[4];
  if (fread(hdr, 4, 4, f) != 4) return 1;
  const int image_bytes = hdr[0], nops = hdr[1], ncols = hdr[2], tcols = hdr[3];
  std::vector<MmaOp> ops(nops);
  std::vector<uint8_t> image(image_bytes);
  if (fread(ops.data(), sizeof(MmaOp), nops, f) != (size_t)nops) return 1;
  if (fread(image.data(), 1, image_bytes, f) != (size_t)image_bytes) return 1;
  std::vector<uint32_t> timage((size_t)128 * (tcols > 0 ? tcols : 1));
  if (tcols > 0 && fread(timage.data(), 4, (size_t)128 * tcols, f) != (size_t)128 * tcols) return 1;
  fclose(f);
  uint32_t* d_timg;
  CK(cudaMalloc(&d_timg, timage.size() * 4));
  CK(cudaMemcpy(d_timg, timage.data(), timage.size() * 4, cudaMemcpyHostToDevice));
  uint8_t* d_img;
  MmaOp* d_ops;
  float* d_out;
  CK(cudaMalloc(&d_img, image_bytes));
  CK(cudaMalloc(&d_ops, nops * sizeof(MmaOp)));
  CK(cudaMalloc(&d_out, 128 * ncols * 4));
  CK(cudaMemcpy(d_img, image.data(), image_bytes, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_ops, ops.data(), nops * sizeof(MmaOp), cudaMemcpyHostToDevice));
  const int smem = image_bytes + 2048;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_kernel<<<1, 128, smem>>>(d_img, image_bytes, d_ops, nops, d_out, ncols, 12345.0f, d_timg, tcols);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> out(128 * ncols);
  CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
  FILE* g = fopen(argv[2], "wb");
  if (!g) return 1;
  fwrite(out.data(), 4, out.size(), g);
  fclose(g);
  return 0;
}

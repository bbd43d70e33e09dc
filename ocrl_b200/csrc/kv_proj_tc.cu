// Token stage on the 5th-generation tensor cores (bf16 mode of ocrl_kv_proj_fwd).
//
//   [pos-table add + NCHW -> token-major]      ocrs/common/utils.py:28-33, slate_module.py:132-133
//   [LayerNorm -> Linear+ReLU -> Linear]       SlotAttentionEncoder, slot_attn.py:125-129,151
//   LayerNorm(norm_inputs) -> [k | v] = x^ [s W_k ; W_v]^T     slot_attn.py:54-61
//
// One persistent CTA per SM, 128-token tiles (UMMA M = 128):
//   * warp 4, one elected lane: TMA producer (cp.async.bulk.tensor: 2-stage ring of fp32 token tiles, the
//     bf16 weight matrices once) and tcgen05.mma issuer;
//   * warps 0-3, one thread per token row: LayerNorm in registers -> bf16 A tile in the canonical 128-byte
//     swizzled K-major layout -> (MMA) -> tcgen05.ld of the thread's own accumulator row from TMEM ->
//     bias / ReLU / LayerNorm -> next A tile ... -> bf16 k|v tile staged in shared memory -> TMA store.
// The three GEMMs of a tile (64->64, 64->64, 64->2D) chain through tensor memory (64 + 64 + 2D <= 512
// columns); activations never return to HBM between the layers.  HBM-bound: 256 B in, 4D B out per token.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace ocrl {

constexpr int PT_TM = 128;   // tokens per tile
constexpr int PT_C = 64;     // token feature width
constexpr int PT_ROWT = 128; // row threads (warps 0-3)
constexpr int PT_NT = 160;   // + producer / MMA warp

struct ProjTcParams {
  const float* enc_ln_w; const float* enc_ln_b;   // null: no token MLP
  const float* b1; const float* b2;
  const float* in_ln_w; const float* in_ln_b;
  const float* pos;                               // [64][N] or null (token-major input)
  float* y_out;                                   // [M][64] or null
  long long* trace;                               // optional clock64 stamps of CTA 0 / thread 0 (development aid)
  int has_mlp, x_format, N, D;
  int pos_tiles;                                  // 1: the position table arrives as bf16 [N][64] tiles through TMA (tm_pos)
  int pad_w, pad_h;                               // > 0: bf16 tokens come from the padded layout of conv_tc.cu (frame W x H)
  int xhat_only;                                  // 1: stop after norm_inputs and store x^ [M][64] bf16 through tm_k (ocrl_xhat_fwd)
  long long M;
  int ntiles;
  float ln_eps;
};

// ---- PTX wrappers -----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tm), "r"(c0), "r"(c1),
               "r"(smem_u32(src))
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// 64 consecutive fp32 columns: both loads are issued before the single wait
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, float (&lo)[32], float (&hi)[32]) {
  uint32_t a[32], b[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31}, [%32];"
      : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]), "=r"(a[8]),
        "=r"(a[9]), "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15]), "=r"(a[16]),
        "=r"(a[17]), "=r"(a[18]), "=r"(a[19]), "=r"(a[20]), "=r"(a[21]), "=r"(a[22]), "=r"(a[23]), "=r"(a[24]),
        "=r"(a[25]), "=r"(a[26]), "=r"(a[27]), "=r"(a[28]), "=r"(a[29]), "=r"(a[30]), "=r"(a[31])
      : "r"(taddr));
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31}, [%32];"
      : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]), "=r"(b[4]), "=r"(b[5]), "=r"(b[6]), "=r"(b[7]), "=r"(b[8]),
        "=r"(b[9]), "=r"(b[10]), "=r"(b[11]), "=r"(b[12]), "=r"(b[13]), "=r"(b[14]), "=r"(b[15]), "=r"(b[16]),
        "=r"(b[17]), "=r"(b[18]), "=r"(b[19]), "=r"(b[20]), "=r"(b[21]), "=r"(b[22]), "=r"(b[23]), "=r"(b[24]),
        "=r"(b[25]), "=r"(b[26]), "=r"(b[27]), "=r"(b[28]), "=r"(b[29]), "=r"(b[30]), "=r"(b[31])
      : "r"(taddr + 32));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) { lo[i] = __uint_as_float(a[i]); hi[i] = __uint_as_float(b[i]); }
}

// K-major, 128-byte-swizzled shared-memory operand descriptor (rows of 64 bf16 = 128 B, 8-row groups 1024 B apart)
__device__ __forceinline__ uint64_t umma_desc_sw128(const void* smem) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_u32(smem) >> 4) & 0x3fffu);  // start address
  d |= (uint64_t)1 << 16;                             // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                             // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
  return d;
}
// bf16 x bf16 -> fp32, A and B K-major, M = 128
__host__ __device__ constexpr uint32_t umma_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// LayerNorm of one row held in registers (four independent partial sums: the 64-long dependent add chains were a
// tenth of a tile's time)
__device__ __forceinline__ void row_layer_norm(float (&x)[PT_C], const float* gw, const float* gb, float eps) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int i = 0; i < PT_C; i += 4) { s0 += x[i]; s1 += x[i + 1]; s2 += x[i + 2]; s3 += x[i + 3]; }
  const float mean = ((s0 + s1) + (s2 + s3)) * (1.f / PT_C);
  float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
  for (int i = 0; i < PT_C; i += 4) {
    x[i] -= mean; x[i + 1] -= mean; x[i + 2] -= mean; x[i + 3] -= mean;
    q0 = fmaf(x[i], x[i], q0); q1 = fmaf(x[i + 1], x[i + 1], q1);
    q2 = fmaf(x[i + 2], x[i + 2], q2); q3 = fmaf(x[i + 3], x[i + 3], q3);
  }
  const float rstd = rsqrtf(((q0 + q1) + (q2 + q3)) * (1.f / PT_C) + eps);
#pragma unroll
  for (int i = 0; i < PT_C; ++i) x[i] = x[i] * rstd * gw[i] + gb[i];
}

// write one row (64 values) as bf16 into a [128][64] K-major SWIZZLE_128B tile
__device__ __forceinline__ void store_row_bf16_sw128(unsigned char* tile, int row, const float (&x)[PT_C]) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint4 u;
    u.x = pack_bf16x2(x[8 * c + 0], x[8 * c + 1]);
    u.y = pack_bf16x2(x[8 * c + 2], x[8 * c + 3]);
    u.z = pack_bf16x2(x[8 * c + 4], x[8 * c + 5]);
    u.w = pack_bf16x2(x[8 * c + 6], x[8 * c + 7]);
    *reinterpret_cast<uint4*>(tile + row * 128 + ((c ^ (row & 7)) << 4)) = u;
  }
}

template <int D>
__global__ void __launch_bounds__(PT_NT, 1)
kv_proj_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w1,
                  const __grid_constant__ CUtensorMap tm_w2, const __grid_constant__ CUtensorMap tm_wkv,
                  const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v,
                  const __grid_constant__ CUtensorMap tm_pos, const ProjTcParams p) {
  constexpr int NKV = 2 * D;                       // output features of the projection
  constexpr int N_A = NKV > 256 ? 256 : NKV;       // first UMMA N
  constexpr int N_B = NKV - N_A;                   // second UMMA N (0 if none)
  constexpr uint32_t COL_G1 = 0, COL_G2 = 64, COL_KV = 128;
  static_assert(COL_KV + NKV <= 512, "tensor memory holds 512 fp32 columns");

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* sp = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* Wkv_s = sp; sp += NKV * 128;
  unsigned char* W1_s = sp; sp += 64 * 128;
  unsigned char* W2_s = sp; sp += 64 * 128;
  unsigned char* A_s = sp; sp += PT_TM * 128;
  unsigned char* O_s = sp; sp += 2 * PT_TM * 128;   // two staging tiles for the TMA stores
  unsigned char* X_s = sp; sp += 2 * PT_TM * PT_C * 4;  // two fp32 token tiles
  float* prm = reinterpret_cast<float*>(sp); sp += 6 * PT_C * sizeof(float);  // enc ln w,b | b1 | b2 | in ln w,b
  uint64_t* bars = reinterpret_cast<uint64_t*>(sp); sp += 12 * sizeof(uint64_t);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sp);
  uint64_t* full = bars;        // [2] token tile landed
  uint64_t* empty = bars + 2;   // [2] token tile consumed by the 128 row threads
  uint64_t* a_ready = bars + 4; // A tile written (128 arrivals)
  uint64_t* mma_done = bars + 5;
  uint64_t* w_full = bars + 6;
  uint64_t* o_ready = bars + 7;  // [2] staging tile written by the 128 row threads
  uint64_t* o_free = bars + 9;   // [2] the TMA store that used the staging tile has read it (producer lane arrives)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid < PT_C) {
    prm[tid] = p.has_mlp ? __ldg(p.enc_ln_w + tid) : 1.f;
    prm[64 + tid] = p.has_mlp ? __ldg(p.enc_ln_b + tid) : 0.f;
    prm[128 + tid] = p.has_mlp ? __ldg(p.b1 + tid) : 0.f;
    prm[192 + tid] = p.has_mlp ? __ldg(p.b2 + tid) : 0.f;
    prm[256 + tid] = __ldg(p.in_ln_w + tid);
    prm[320 + tid] = __ldg(p.in_ln_b + tid);
  }
  if (tid == 0) {
    mbar_init(&full[0], 1); mbar_init(&full[1], 1);
    mbar_init(&empty[0], PT_ROWT); mbar_init(&empty[1], PT_ROWT);
    mbar_init(a_ready, PT_ROWT);
    mbar_init(mma_done, 1);
    mbar_init(w_full, 1);
    mbar_init(&o_ready[0], PT_ROWT); mbar_init(&o_ready[1], PT_ROWT);
    mbar_init(&o_free[0], 1); mbar_init(&o_free[1], 1);
  }
  mbar_fence_init();
  if (warp == 4) {  // tensor memory: 512 columns for this CTA
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int ntiles = p.ntiles;
  const int my_tiles = (ntiles > (int)blockIdx.x) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  constexpr uint32_t X_BYTES = PT_TM * PT_C * 4;

  if (warp == 4) {
    if (lane == 0) {
      // ============================ TMA producer + MMA issuer ==========================================
      auto load_x = [&](int it) {
        const long long tile = blockIdx.x + (long long)it * gridDim.x;
        const int s = it & 1;
        unsigned char* dst = X_s + s * X_BYTES;
        mbar_expect_tx(&full[s], (p.x_format == OCRL_X_TOKENS_BF16 && !p.pos_tiles) ? X_BYTES / 2 : X_BYTES);
        if (p.x_format == OCRL_X_NCHW_F32) {
          const long long m0 = tile * PT_TM;
          const int b = (int)(m0 / p.N), n0 = (int)(m0 % p.N);
          tma_load_2d(dst, &tm_x, n0, b * PT_C, &full[s]);           // [64 ch][128 tokens] fp32
        } else if (p.x_format == OCRL_X_TOKENS_BF16) {
          if (p.pad_w > 0) {  // 128 tokens = 128 / W image rows of the padded feature map, one box of W positions each
            const long long m0 = tile * PT_TM;
            const int b = (int)(m0 / p.N), y0 = (int)(m0 % p.N) / p.pad_w;
            for (int j = 0; j < PT_TM / p.pad_w; ++j)
              tma_load_2d(dst + j * p.pad_w * 128, &tm_x, 0, (2 + b * (p.pad_h + 2) + y0 + j) * (p.pad_w + 4) + 2, &full[s]);
          } else {
            tma_load_2d(dst, &tm_x, 0, (int)(tile * PT_TM), &full[s]);  // [128 tokens][64] bf16, swizzled
          }
          if (p.pos_tiles)  // the matching rows of the bf16 position table ride in the unused half of the stage
            tma_load_2d(dst + X_BYTES / 2, &tm_pos, 0, (int)((tile * PT_TM) % p.N), &full[s]);
        } else {
          tma_load_2d(dst, &tm_x, 0, (int)(tile * PT_TM), &full[s]);  // cols 0-31, 128 rows, swizzled
          tma_load_2d(dst + X_BYTES / 2, &tm_x, 32, (int)(tile * PT_TM), &full[s]);
        }
      };
      // weights, once
      uint32_t wbytes = NKV * 128;
      if (p.has_mlp) wbytes += 2 * 64 * 128;
      mbar_expect_tx(w_full, wbytes);
#pragma unroll
      for (int r = 0; r < NKV / 128; ++r) tma_load_2d(Wkv_s + r * 128 * 128, &tm_wkv, 0, r * 128, w_full);
      if (p.has_mlp) {
        tma_load_2d(W1_s, &tm_w1, 0, 0, w_full);
        tma_load_2d(W2_s, &tm_w2, 0, 0, w_full);
      }
      for (int it = 0; it < 2 && it < my_tiles; ++it) load_x(it);
      mbar_wait(w_full, 0);

      uint32_t ar_phase = 0;
      auto gemm = [&](const unsigned char* b_tile, int n, uint32_t col) {
        mbar_wait(a_ready, ar_phase);
        ar_phase ^= 1;
        tc_fence_after();
        const uint64_t da = umma_desc_sw128(A_s), db = umma_desc_sw128(b_tile);
#pragma unroll
        for (int ks = 0; ks < PT_C / 16; ++ks)  // +32 bytes along K inside the swizzle atom per step
          umma_bf16(tmem + col, da + 2 * ks, db + 2 * ks, umma_idesc(n), ks > 0);
      };
      for (int it = 0; it < my_tiles; ++it) {
        if (p.has_mlp) {
          gemm(W1_s, 64, COL_G1);
          umma_commit(mma_done);
          gemm(W2_s, 64, COL_G2);
          umma_commit(mma_done);
        }
        gemm(Wkv_s, N_A, COL_KV);
        if (N_B > 0) {
          const uint64_t da = umma_desc_sw128(A_s), db = umma_desc_sw128(Wkv_s + N_A * 128);
#pragma unroll
          for (int ks = 0; ks < PT_C / 16; ++ks)
            umma_bf16(tmem + COL_KV + N_A, da + 2 * ks, db + 2 * ks, umma_idesc(N_B > 0 ? N_B : 16), ks > 0);
        }
        umma_commit(mma_done);
        // refill the token-tile stage this tile used, once its 128 readers are done with it
        if (it + 2 < my_tiles) {
          mbar_wait(&empty[it & 1], (uint32_t)((it >> 1) & 1));
          load_x(it + 2);
        }
        // the tile's k | v TMA stores are issued here, off the row threads' critical path: chunk c of this CTA's
        // store sequence uses staging tile c & 1; a tile is handed back once its store has read it
        const long long tile = blockIdx.x + (long long)it * gridDim.x;
        for (int ch = 0; ch < NKV / 64; ++ch) {
          const int sq = it * (NKV / 64) + ch, buf = sq & 1;
          mbar_wait(&o_ready[buf], (uint32_t)((sq >> 1) & 1));
          const int f0 = 64 * ch;
          if (f0 < D) tma_store_2d(&tm_k, O_s + buf * (PT_TM * 128), f0, (int)(tile * PT_TM));
          else tma_store_2d(&tm_v, O_s + buf * (PT_TM * 128), f0 - D, (int)(tile * PT_TM));
          tma_store_commit();
          if (sq >= 1) {  // the previous store (other staging tile) has been read: that tile is free again
            tma_store_wait_read<1>();
            mbar_arrive(&o_free[buf ^ 1]);
          }
        }
      }
      tma_store_wait_all<0>();
    }
  } else {
    // ================================ row threads: one token each ==========================================
    const int row = tid;  // TMEM lane = accumulator row = token inside the tile
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    uint32_t md_phase = 0;
    int store_seq = 0;  // TMA stores issued so far by thread 0 (selects the staging tile)
    for (int it = 0; it < my_tiles; ++it) {
      const long long tile = blockIdx.x + (long long)it * gridDim.x;
      const long long m = tile * PT_TM + row;
      const int s = it & 1;
      float x[PT_C];
      const bool tr = (p.trace != nullptr && blockIdx.x == 0 && tid == 0 && it >= 2 && it < 10);
      if (tr) p.trace[(it - 2) * 8 + 0] = clock64();
      mbar_wait(&full[s], (uint32_t)((it >> 1) & 1));
      if (tr) p.trace[(it - 2) * 8 + 1] = clock64();
      const unsigned char* xs = X_s + s * X_BYTES;
      if (p.x_format == OCRL_X_NCHW_F32) {
#pragma unroll
        for (int c = 0; c < PT_C; ++c) x[c] = reinterpret_cast<const float*>(xs)[c * PT_TM + row];
      } else if (p.x_format == OCRL_X_TOKENS_BF16) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 u = *reinterpret_cast<const uint4*>(xs + row * 128 + ((c ^ (row & 7)) << 4));
          const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            x[8 * c + 2 * i] = __uint_as_float(w4[i] << 16);
            x[8 * c + 2 * i + 1] = __uint_as_float(w4[i] & 0xffff0000u);
          }
          if (p.pos_tiles) {
            const uint4 q = *reinterpret_cast<const uint4*>(xs + X_BYTES / 2 + row * 128 + ((c ^ (row & 7)) << 4));
            const uint32_t q4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              x[8 * c + 2 * i] += __uint_as_float(q4[i] << 16);
              x[8 * c + 2 * i + 1] += __uint_as_float(q4[i] & 0xffff0000u);
            }
          }
        }
      } else {
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 v4 = *reinterpret_cast<const float4*>(xs + h * (X_BYTES / 2) + row * 128 + ((c ^ (row & 7)) << 4));
            x[32 * h + 4 * c + 0] = v4.x; x[32 * h + 4 * c + 1] = v4.y;
            x[32 * h + 4 * c + 2] = v4.z; x[32 * h + 4 * c + 3] = v4.w;
          }
      }
      mbar_arrive(&empty[s]);
      if (p.pos != nullptr && !p.pos_tiles) {  // position table [64][N]: coalesced across the 128 row threads
        const int n = (int)((tile * PT_TM + row) % p.N);
#pragma unroll
        for (int c = 0; c < PT_C; ++c) x[c] += __ldg(p.pos + (size_t)c * p.N + n);
      }

      if (tr) p.trace[(it - 2) * 8 + 2] = clock64();
      if (p.has_mlp) {
        row_layer_norm(x, prm, prm + 64, p.ln_eps);
        store_row_bf16_sw128(A_s, row, x);
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(a_ready);
        if (tr) p.trace[(it - 2) * 8 + 3] = clock64();
        // hidden layer: relu(acc + b1)
        mbar_wait(mma_done, md_phase); md_phase ^= 1;
        tc_fence_after();
        {
          float acc[32];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            tmem_ld32(trow + COL_G1 + 32 * h, acc);
#pragma unroll
            for (int i = 0; i < 32; ++i) x[32 * h + i] = fmaxf(acc[i] + prm[128 + 32 * h + i], 0.f);
          }
        }
        store_row_bf16_sw128(A_s, row, x);
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(a_ready);
        if (tr) p.trace[(it - 2) * 8 + 4] = clock64();
        // output layer: acc + b2  (= the token-MLP output y)
        mbar_wait(mma_done, md_phase); md_phase ^= 1;
        tc_fence_after();
        {
          float acc[32];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            tmem_ld32(trow + COL_G2 + 32 * h, acc);
#pragma unroll
            for (int i = 0; i < 32; ++i) x[32 * h + i] = acc[i] + prm[192 + 32 * h + i];
          }
        }
      }
      if (p.y_out != nullptr && m < p.M) {
#pragma unroll
        for (int c = 0; c < PT_C / 4; ++c)
          *reinterpret_cast<float4*>(p.y_out + m * PT_C + 4 * c) = make_float4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]);
      }
      row_layer_norm(x, prm + 256, prm + 320, p.ln_eps);
      store_row_bf16_sw128(A_s, row, x);
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(a_ready);

      if (tr) p.trace[(it - 2) * 8 + 5] = clock64();
      // ---- k | v epilogue: TMEM -> bf16 -> swizzled staging tile -> TMA store, 64 features at a time ----
      mbar_wait(mma_done, md_phase); md_phase ^= 1;
      if (tr) p.trace[(it - 2) * 8 + 6] = clock64();
      tc_fence_after();
#pragma unroll 1
      for (int ch = 0; ch < NKV / 64; ++ch) {
        float lo[32], hi[32];
        tmem_ld64(trow + COL_KV + 64 * ch, lo, hi);
        const int buf = store_seq & 1;
        unsigned char* ot = O_s + buf * (PT_TM * 128);
        // staging tile `buf` was last used by store store_seq - 2; its hand-back is phase (store_seq / 2 - 1)
        if (store_seq >= 2) mbar_wait(&o_free[buf], (uint32_t)(((store_seq >> 1) - 1) & 1));
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float* src = (c < 4) ? (lo + 8 * c) : (hi + 8 * (c - 4));
          uint4 u;
          u.x = pack_bf16x2(src[0], src[1]); u.y = pack_bf16x2(src[2], src[3]);
          u.z = pack_bf16x2(src[4], src[5]); u.w = pack_bf16x2(src[6], src[7]);
          *reinterpret_cast<uint4*>(ot + row * 128 + ((c ^ (row & 7)) << 4)) = u;
        }
        fence_proxy_async();
        mbar_arrive(&o_ready[buf]);
        ++store_seq;
      }
      tc_fence_before();
      if (tr) p.trace[(it - 2) * 8 + 7] = clock64();
    }
  }
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Two row groups per CTA.  The kernel above runs ONE serial chain per SM (LayerNorm -> MMA -> tcgen05.ld -> ... ->
// epilogue, 9.1 k cycles per 128-token tile of which the tensor pipe works 0.5 k): every step waits for the previous
// one.  Here two independent chains share an SM: row warps 0-3 + producer warp 8 work on the CTA's even tiles, row warps
// 4-7 + producer warp 9 on the odd ones, each with its own token stage, A tile, staging tiles, barriers and half of the
// tensor memory (256 columns: 64 shared by the two MLP accumulators, 192 for the projection, which therefore runs as a
// k pass and a v pass through the same columns for D > 64).  The weights in shared memory are shared.
constexpr int PT2_NT = 320;  // 2 x 4 row warps + 2 producer / MMA warps

template <int D>
__global__ void __launch_bounds__(PT2_NT, 1)
kv_proj_tc2_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w1,
                   const __grid_constant__ CUtensorMap tm_w2, const __grid_constant__ CUtensorMap tm_wkv,
                   const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v,
                   const __grid_constant__ CUtensorMap tm_pos, const ProjTcParams p) {
  constexpr int NKV = 2 * D;
  constexpr int PASS_N = (D == 64) ? 128 : D;     // output features per projection pass (k | v together for D = 64)
  constexpr int NPASS = NKV / PASS_N;             // 1 or 2
  constexpr int CPP = PASS_N / 64;                // 64-feature chunks per pass
  constexpr uint32_t X_BYTES = PT_TM * PT_C * 4;
  constexpr uint32_t GRP_BYTES = X_BYTES + PT_TM * 128 + 2 * PT_TM * 128;  // token stage, A tile, two staging tiles
  static_assert(64 + PASS_N <= 256, "a group owns 256 tensor-memory columns");

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* sp = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* Wkv_s = sp; sp += NKV * 128;
  unsigned char* W1_s = sp; sp += 64 * 128;
  unsigned char* W2_s = sp; sp += 64 * 128;
  unsigned char* grp_s = sp; sp += 2 * GRP_BYTES;
  float* prm = reinterpret_cast<float*>(sp); sp += 6 * PT_C * sizeof(float);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sp); sp += 20 * sizeof(uint64_t);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sp);
  uint64_t* w_full = bars;  // shared; then per group: full, empty, a_ready, mma_done, kv_free, o_ready[2], o_free[2]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = (warp < 8) ? (warp >> 2) : (warp - 8);  // group of this warp
  uint64_t* gb = bars + 1 + 9 * g;
  uint64_t* full = gb;
  uint64_t* empty = gb + 1;
  uint64_t* a_ready = gb + 2;
  uint64_t* mma_done = gb + 3;
  uint64_t* kv_free = gb + 4;
  uint64_t* o_ready = gb + 5;
  uint64_t* o_free = gb + 7;
  unsigned char* X_s = grp_s + g * GRP_BYTES;
  unsigned char* A_s = X_s + X_BYTES;
  unsigned char* O_s = A_s + PT_TM * 128;

  if (tid < PT_C) {
    prm[tid] = p.has_mlp ? __ldg(p.enc_ln_w + tid) : 1.f;
    prm[64 + tid] = p.has_mlp ? __ldg(p.enc_ln_b + tid) : 0.f;
    prm[128 + tid] = p.has_mlp ? __ldg(p.b1 + tid) : 0.f;
    prm[192 + tid] = p.has_mlp ? __ldg(p.b2 + tid) : 0.f;
    prm[256 + tid] = __ldg(p.in_ln_w + tid);
    prm[320 + tid] = __ldg(p.in_ln_b + tid);
  }
  if (tid == 0) {
    mbar_init(w_full, 1);
    for (int q = 0; q < 2; ++q) {
      uint64_t* b = bars + 1 + 9 * q;
      mbar_init(b + 0, 1);
      mbar_init(b + 1, PT_ROWT);
      mbar_init(b + 2, PT_ROWT);
      mbar_init(b + 3, 1);
      mbar_init(b + 4, PT_ROWT);
      mbar_init(b + 5, PT_ROWT); mbar_init(b + 6, PT_ROWT);
      mbar_init(b + 7, 1); mbar_init(b + 8, 1);
    }
  }
  mbar_fence_init();
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot + 256u * (uint32_t)g;
  constexpr uint32_t COL_G = 0, COL_KV = 64;

  const int cta_tiles = (p.ntiles > (int)blockIdx.x) ? (p.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int my_tiles = (cta_tiles > g) ? (cta_tiles - g + 1) / 2 : 0;       // the CTA's tiles g, g + 2, ...
  auto tile_of = [&](int it) { return (long long)blockIdx.x + (long long)(2 * it + g) * gridDim.x; };

  if (warp >= 8) {
    if (lane == 0) {
      // ============================ TMA producer + MMA issuer of group g =================================
      auto load_x = [&](int it) {
        const long long tile = tile_of(it);
        unsigned char* dst = X_s;
        mbar_expect_tx(full, (p.x_format == OCRL_X_TOKENS_BF16 && !p.pos_tiles) ? X_BYTES / 2 : X_BYTES);
        if (p.x_format == OCRL_X_NCHW_F32) {
          const long long m0 = tile * PT_TM;
          const int b = (int)(m0 / p.N), n0 = (int)(m0 % p.N);
          tma_load_2d(dst, &tm_x, n0, b * PT_C, full);
        } else if (p.x_format == OCRL_X_TOKENS_BF16) {
          if (p.pad_w > 0) {
            const long long m0 = tile * PT_TM;
            const int b = (int)(m0 / p.N), y0 = (int)(m0 % p.N) / p.pad_w;
            for (int j = 0; j < PT_TM / p.pad_w; ++j)
              tma_load_2d(dst + j * p.pad_w * 128, &tm_x, 0, (2 + b * (p.pad_h + 2) + y0 + j) * (p.pad_w + 4) + 2, full);
          } else {
            tma_load_2d(dst, &tm_x, 0, (int)(tile * PT_TM), full);
          }
          if (p.pos_tiles) tma_load_2d(dst + X_BYTES / 2, &tm_pos, 0, (int)((tile * PT_TM) % p.N), full);
        } else {
          tma_load_2d(dst, &tm_x, 0, (int)(tile * PT_TM), full);
          tma_load_2d(dst + X_BYTES / 2, &tm_x, 32, (int)(tile * PT_TM), full);
        }
      };
      if (g == 0) {  // weights, once per CTA
        uint32_t wbytes = p.xhat_only ? 0u : NKV * 128;
        if (p.has_mlp) wbytes += 2 * 64 * 128;
        if (wbytes == 0) mbar_arrive(w_full);
        else mbar_expect_tx(w_full, wbytes);
        if (!p.xhat_only) {
#pragma unroll
          for (int r = 0; r < NKV / 128; ++r) tma_load_2d(Wkv_s + r * 128 * 128, &tm_wkv, 0, r * 128, w_full);
        }
        if (p.has_mlp) {
          tma_load_2d(W1_s, &tm_w1, 0, 0, w_full);
          tma_load_2d(W2_s, &tm_w2, 0, 0, w_full);
        }
      }
      if (my_tiles > 0) load_x(0);
      mbar_wait(w_full, 0);

      uint32_t ar_phase = 0, kf_phase = 0;
      int sq = 0;  // stores issued by this producer (selects the staging tile)
      auto mma_tile = [&](const unsigned char* b_tile, int n, uint32_t col) {
        tc_fence_after();
        const uint64_t da = umma_desc_sw128(A_s), db = umma_desc_sw128(b_tile);
#pragma unroll
        for (int ks = 0; ks < PT_C / 16; ++ks) umma_bf16(tmem + col, da + 2 * ks, db + 2 * ks, umma_idesc(n), ks > 0);
        umma_commit(mma_done);
      };
      for (int it = 0; it < my_tiles; ++it) {
        if (p.has_mlp) {
          mbar_wait(a_ready, ar_phase); ar_phase ^= 1;
          mma_tile(W1_s, 64, COL_G);
          mbar_wait(a_ready, ar_phase); ar_phase ^= 1;   // (the row threads have read the first accumulator by now)
          mma_tile(W2_s, 64, COL_G);
        }
        const long long tile = tile_of(it);
        if (p.xhat_only) {  // no projection: the row threads stage x^ itself, one 128 x 64 tile per token tile
          if (it + 1 < my_tiles) {
            mbar_wait(empty, (uint32_t)(it & 1));
            load_x(it + 1);
          }
          const int buf = sq & 1;
          mbar_wait(&o_ready[buf], (uint32_t)((sq >> 1) & 1));
          tma_store_2d(&tm_k, O_s + buf * (PT_TM * 128), 0, (int)(tile * PT_TM));
          tma_store_commit();
          if (sq >= 1) {
            tma_store_wait_read<1>();
            mbar_arrive(&o_free[buf ^ 1]);
          }
          ++sq;
          continue;
        }
#pragma unroll 1
        for (int ps = 0; ps < NPASS; ++ps) {
          if (ps == 0) { mbar_wait(a_ready, ar_phase); ar_phase ^= 1; }
          else { mbar_wait(kv_free, kf_phase); kf_phase ^= 1; }   // the k pass has been read out of the columns
          mma_tile(Wkv_s + ps * PASS_N * 128, PASS_N, COL_KV);
          if (ps == 0 && it + 1 < my_tiles) {  // the token stage is free as soon as the rows sit in registers
            mbar_wait(empty, (uint32_t)(it & 1));
            load_x(it + 1);
          }
          for (int chl = 0; chl < CPP; ++chl, ++sq) {
            const int buf = sq & 1;
            mbar_wait(&o_ready[buf], (uint32_t)((sq >> 1) & 1));
            const int f0 = 64 * (ps * CPP + chl);
            if (f0 < D) tma_store_2d(&tm_k, O_s + buf * (PT_TM * 128), f0, (int)(tile * PT_TM));
            else tma_store_2d(&tm_v, O_s + buf * (PT_TM * 128), f0 - D, (int)(tile * PT_TM));
            tma_store_commit();
            if (sq >= 1) {
              tma_store_wait_read<1>();
              mbar_arrive(&o_free[buf ^ 1]);
            }
          }
        }
      }
      tma_store_wait_all<0>();
    }
  } else {
    // ================================ row threads of group g: one token each ================================
    const int row = tid & 127;
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t md_phase = 0;
    int store_seq = 0;
    for (int it = 0; it < my_tiles; ++it) {
      const long long tile = tile_of(it);
      const long long m = tile * PT_TM + row;
      float x[PT_C];
      mbar_wait(full, (uint32_t)(it & 1));
      const unsigned char* xs = X_s;
      if (p.x_format == OCRL_X_NCHW_F32) {
#pragma unroll
        for (int c = 0; c < PT_C; ++c) x[c] = reinterpret_cast<const float*>(xs)[c * PT_TM + row];
      } else if (p.x_format == OCRL_X_TOKENS_BF16) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 u = *reinterpret_cast<const uint4*>(xs + row * 128 + ((c ^ (row & 7)) << 4));
          const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            x[8 * c + 2 * i] = __uint_as_float(w4[i] << 16);
            x[8 * c + 2 * i + 1] = __uint_as_float(w4[i] & 0xffff0000u);
          }
          if (p.pos_tiles) {
            const uint4 q = *reinterpret_cast<const uint4*>(xs + X_BYTES / 2 + row * 128 + ((c ^ (row & 7)) << 4));
            const uint32_t q4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              x[8 * c + 2 * i] += __uint_as_float(q4[i] << 16);
              x[8 * c + 2 * i + 1] += __uint_as_float(q4[i] & 0xffff0000u);
            }
          }
        }
      } else {
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 v4 = *reinterpret_cast<const float4*>(xs + h * (X_BYTES / 2) + row * 128 + ((c ^ (row & 7)) << 4));
            x[32 * h + 4 * c + 0] = v4.x; x[32 * h + 4 * c + 1] = v4.y;
            x[32 * h + 4 * c + 2] = v4.z; x[32 * h + 4 * c + 3] = v4.w;
          }
      }
      mbar_arrive(empty);
      if (p.pos != nullptr && !p.pos_tiles) {
        const int n = (int)((tile * PT_TM + row) % p.N);
#pragma unroll
        for (int c = 0; c < PT_C; ++c) x[c] += __ldg(p.pos + (size_t)c * p.N + n);
      }
      if (p.has_mlp) {
        row_layer_norm(x, prm, prm + 64, p.ln_eps);
        store_row_bf16_sw128(A_s, row, x);
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(a_ready);
        mbar_wait(mma_done, md_phase); md_phase ^= 1;
        tc_fence_after();
        {
          float acc[32];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            tmem_ld32(trow + COL_G + 32 * h, acc);
#pragma unroll
            for (int i = 0; i < 32; ++i) x[32 * h + i] = fmaxf(acc[i] + prm[128 + 32 * h + i], 0.f);
          }
        }
        store_row_bf16_sw128(A_s, row, x);
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(a_ready);
        mbar_wait(mma_done, md_phase); md_phase ^= 1;
        tc_fence_after();
        {
          float acc[32];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            tmem_ld32(trow + COL_G + 32 * h, acc);
#pragma unroll
            for (int i = 0; i < 32; ++i) x[32 * h + i] = acc[i] + prm[192 + 32 * h + i];
          }
        }
      }
      if (p.y_out != nullptr && m < p.M) {
#pragma unroll
        for (int c = 0; c < PT_C / 4; ++c)
          *reinterpret_cast<float4*>(p.y_out + m * PT_C + 4 * c) = make_float4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]);
      }
      row_layer_norm(x, prm + 256, prm + 320, p.ln_eps);
      if (p.xhat_only) {  // x^ as bf16 into a staging tile (the layout of the A tiles = the TMA store's swizzle)
        const int buf = store_seq & 1;
        if (store_seq >= 2) mbar_wait(&o_free[buf], (uint32_t)(((store_seq >> 1) - 1) & 1));
        store_row_bf16_sw128(O_s + buf * (PT_TM * 128), row, x);
        fence_proxy_async();
        mbar_arrive(&o_ready[buf]);
        ++store_seq;
        tc_fence_before();
        continue;
      }
      store_row_bf16_sw128(A_s, row, x);
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(a_ready);
      // ---- k pass, v pass: TMEM -> bf16 -> swizzled staging tile -> TMA store (by the producer), 64 features at a time
#pragma unroll 1
      for (int ps = 0; ps < NPASS; ++ps) {
        mbar_wait(mma_done, md_phase); md_phase ^= 1;
        tc_fence_after();
#pragma unroll 1
        for (int chl = 0; chl < CPP; ++chl) {
          float lo[32], hi[32];
          tmem_ld64(trow + COL_KV + 64 * chl, lo, hi);
          if (chl == CPP - 1 && ps + 1 < NPASS) {  // the columns may take the next pass
            tc_fence_before();
            mbar_arrive(kv_free);
          }
          const int buf = store_seq & 1;
          unsigned char* ot = O_s + buf * (PT_TM * 128);
          if (store_seq >= 2) mbar_wait(&o_free[buf], (uint32_t)(((store_seq >> 1) - 1) & 1));
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float* src = (c < 4) ? (lo + 8 * c) : (hi + 8 * (c - 4));
            uint4 u;
            u.x = pack_bf16x2(src[0], src[1]); u.y = pack_bf16x2(src[2], src[3]);
            u.z = pack_bf16x2(src[4], src[5]); u.w = pack_bf16x2(src[6], src[7]);
            *reinterpret_cast<uint4*>(ot + row * 128 + ((c ^ (row & 7)) << 4)) = u;
          }
          fence_proxy_async();
          mbar_arrive(&o_ready[buf]);
          ++store_seq;
        }
      }
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 8) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(*tmem_slot) : "memory");
  }
}

// fp32 weights -> bf16 copies in the layout the tensor maps describe ([rows][64], k rows pre-scaled by D^-1/2)
__global__ void proj_tc_prep_kernel(const float* __restrict__ w1, const float* __restrict__ w2,
                                    const float* __restrict__ wk, const float* __restrict__ wv, __nv_bfloat16* w1b,
                                    __nv_bfloat16* w2b, __nv_bfloat16* wkvb, int D, float kscale,
                                    const float* __restrict__ pos, __nv_bfloat16* posb, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (posb != nullptr)  // position table [64][N] fp32 -> token-major bf16 [N][64]: coalesced reads, 16-byte writes
    for (int j = i; j < 8 * N; j += gridDim.x * blockDim.x) {
      const int n = j % N, c8 = j / N;
      uint32_t q[4];
#pragma unroll
      for (int e = 0; e < 4; ++e)
        q[e] = pack_bf16x2(__ldg(pos + (size_t)(8 * c8 + 2 * e) * N + n), __ldg(pos + (size_t)(8 * c8 + 2 * e + 1) * N + n));
      *reinterpret_cast<uint4*>(posb + (size_t)n * 64 + 8 * c8) = make_uint4(q[0], q[1], q[2], q[3]);
    }
  if (i < 64 * 64) {
    if (w1 != nullptr) { w1b[i] = __float2bfloat16_rn(w1[i]); w2b[i] = __float2bfloat16_rn(w2[i]); }
  }
  if (wk != nullptr && i < D * 64) {
    wkvb[i] = __float2bfloat16_rn(wk[i] * kscale);
    wkvb[D * 64 + i] = __float2bfloat16_rn(wv[i]);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
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

static bool make_map(CUtensorMap* tm, CUtensorMapDataType dt, int esz, const void* base, uint64_t inner, uint64_t outer,
                     uint64_t row_bytes, uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle sw) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  (void)esz;
  return fn(tm, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int g_proj_variant = 0;  // 0: two row groups per CTA, 1: the single-chain kernel (development knob, not in the header)
}  // namespace ocrl
extern "C" void ocrl_dev_proj_variant(int v) { ocrl::g_proj_variant = v; }
namespace ocrl {

size_t kv_proj_tc_workspace(const ocrl_sa_dims* d) {
  return sizeof(__nv_bfloat16) * ((size_t)2 * 64 * 64 + (size_t)2 * d->D * 64 + (size_t)d->N * 64) + 256 + 1024;  // + bf16 position table, trace slots
}

// returns OCRL_E_SHAPE when this shape has to take the FFMA kernel instead
int kv_proj_tc_launch(const ocrl_sa_dims* d, const void* x, const float* pos, const ocrl_token_weights* w, float* y_out,
                      void* k_out, void* v_out, void* workspace, cudaStream_t stream, void* xhat_out) {
  const int D = d->D;
  const bool xhat_only = (xhat_out != nullptr);
  const long long M = (long long)d->B * d->N;
  if (d->C_in != PT_C || (D != 64 && D != 128 && D != 192) || workspace == nullptr) return OCRL_E_SHAPE;
  if (d->x_format == OCRL_X_NCHW_F32 && (d->N % PT_TM) != 0) return OCRL_E_SHAPE;  // a tile must not straddle two images
  const bool padded = (d->x_format == OCRL_X_PADDED_BF16);
  const int pw = padded ? d->frame_w : 0, ph = (padded && pw > 0) ? d->N / pw : 0;
  if (padded && ((pw != 32 && pw != 64 && pw != 128) || d->N % pw != 0 || d->N % PT_TM != 0)) {
    set_error("kv_proj(tensor): padded feature map needs frame_w in {32, 64, 128} and N a multiple of 128 (frame_w=%d, N=%d)", pw, d->N);
    return OCRL_E_SHAPE;
  }
  if (!encode_fn()) return OCRL_E_SHAPE;
  __nv_bfloat16* w1b = reinterpret_cast<__nv_bfloat16*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  __nv_bfloat16* w2b = w1b + 64 * 64;
  __nv_bfloat16* wkvb = w2b + 64 * 64;
  const bool has_mlp = w->mlp_w1 != nullptr;
  // bf16 tokens with a position table whose tiles never straddle two images: the table goes through TMA as well
  const bool pos_tiles = ((d->x_format == OCRL_X_TOKENS_BF16 || padded) && pos != nullptr && d->N % PT_TM == 0);
  __nv_bfloat16* posb = wkvb + (size_t)2 * D * 64;
  const int prep_threads = (pos_tiles && 8 * d->N > D * 64) ? 8 * d->N : D * 64;
  proj_tc_prep_kernel<<<(prep_threads + 255) / 256, 256, 0, stream>>>(w->mlp_w1, w->mlp_w2, w->wk, w->wv, w1b, w2b, wkvb, D,
                                                                 1.0f / sqrtf((float)D), pos, pos_tiles ? posb : nullptr,
                                                                 d->N);
  ocrl::count_launch();
  OCRL_CHECK_CUDA(cudaGetLastError());

  CUtensorMap tm_x, tm_w1, tm_w2, tm_wkv, tm_k, tm_v, tm_pos;
  bool ok = true;
  if (d->x_format == OCRL_X_NCHW_F32)  // NCHW feature map viewed as [B*64][N]; box = 128 tokens x 64 channels
    ok &= make_map(&tm_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x, (uint64_t)d->N, (uint64_t)d->B * PT_C,
                   (uint64_t)d->N * 4, PT_TM, PT_C, CU_TENSOR_MAP_SWIZZLE_NONE);
  else if (padded)  // flat positions [(2 + B (H + 2)) (W + 4)][64] bf16: boxes of W positions (one image row)
    ok &= make_map(&tm_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, x, PT_C, (uint64_t)(2 + (uint64_t)d->B * (ph + 2)) * (pw + 4),
                   PT_C * 2, PT_C, (uint32_t)pw, CU_TENSOR_MAP_SWIZZLE_128B);
  else if (d->x_format == OCRL_X_TOKENS_BF16)  // bf16 tokens [M][64]: one 128-byte-swizzled box of 128 rows
    ok &= make_map(&tm_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, x, PT_C, (uint64_t)M, PT_C * 2, PT_C, PT_TM,
                   CU_TENSOR_MAP_SWIZZLE_128B);
  else  // tokens [M][64]; two boxes of 32 features (128 B) x 128 rows, 128-byte swizzle
    ok &= make_map(&tm_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x, PT_C, (uint64_t)M, PT_C * 4, 32, PT_TM,
                   CU_TENSOR_MAP_SWIZZLE_128B);
  ok &= make_map(&tm_w1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w1b, 64, 64, 128, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B);
  ok &= make_map(&tm_w2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w2b, 64, 64, 128, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B);
  ok &= make_map(&tm_wkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, wkvb, 64, (uint64_t)2 * D, 128, 64, 128,
                 CU_TENSOR_MAP_SWIZZLE_128B);
  if (xhat_only) {  // x^ [M][64] bf16
    ok &= make_map(&tm_k, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, xhat_out, PT_C, (uint64_t)M, PT_C * 2, PT_C, PT_TM,
                   CU_TENSOR_MAP_SWIZZLE_128B);
    tm_v = tm_k;
  } else {
    ok &= make_map(&tm_k, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, k_out, (uint64_t)D, (uint64_t)M, (uint64_t)D * 2, 64, PT_TM,
                   CU_TENSOR_MAP_SWIZZLE_128B);
    ok &= make_map(&tm_v, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, v_out, (uint64_t)D, (uint64_t)M, (uint64_t)D * 2, 64, PT_TM,
                   CU_TENSOR_MAP_SWIZZLE_128B);
  }
  if (pos_tiles)
    ok &= make_map(&tm_pos, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, posb, PT_C, (uint64_t)d->N, PT_C * 2, PT_C, PT_TM,
                   CU_TENSOR_MAP_SWIZZLE_128B);
  else
    tm_pos = tm_x;  // unused
  if (!ok) {
    set_error("kv_proj(tensor): cuTensorMapEncodeTiled failed");
    return OCRL_E_LAUNCH;
  }
  ProjTcParams p;
  p.enc_ln_w = w->enc_ln_w; p.enc_ln_b = w->enc_ln_b; p.b1 = w->mlp_b1; p.b2 = w->mlp_b2;
  p.in_ln_w = w->in_ln_w; p.in_ln_b = w->in_ln_b; p.pos = pos; p.y_out = y_out;
  p.trace = nullptr;
  p.pos_tiles = pos_tiles ? 1 : 0;
  p.has_mlp = has_mlp ? 1 : 0; p.x_format = padded ? OCRL_X_TOKENS_BF16 : d->x_format; p.N = d->N; p.D = D; p.M = M;
  p.pad_w = pw; p.pad_h = ph;
  p.xhat_only = xhat_only ? 1 : 0;
  p.ntiles = (int)((M + PT_TM - 1) / PT_TM);
  p.ln_eps = d->ln_eps;
  int dev = 0, sms = 148;
  OCRL_CHECK_CUDA(cudaGetDevice(&dev));
  OCRL_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (g_proj_variant == 0 || xhat_only) {  // two row groups per CTA (default; the only kernel with the x^ mode)
    const int grid2 = (p.ntiles + 1) / 2 < sms ? (p.ntiles + 1) / 2 : sms;
    const size_t smem2 = 1024 + (size_t)2 * D * 128 + 2 * 64 * 128 + 2 * (PT_TM * PT_C * 4 + 3 * PT_TM * 128) + 6 * PT_C * 4 +
                         20 * 8 + 16;
#define OCRL_LAUNCH_PT2(DD)                                                                                      \
  do {                                                                                                           \
    OCRL_CHECK_CUDA(cudaFuncSetAttribute(kv_proj_tc2_kernel<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                         (int)smem2));                                                           \
    kv_proj_tc2_kernel<DD><<<grid2, PT2_NT, smem2, stream>>>(tm_x, tm_w1, tm_w2, tm_wkv, tm_k, tm_v, tm_pos, p); ocrl::count_launch(); \
  } while (0)
    if (D == 64) OCRL_LAUNCH_PT2(64);
    else if (D == 128) OCRL_LAUNCH_PT2(128);
    else OCRL_LAUNCH_PT2(192);
#undef OCRL_LAUNCH_PT2
    OCRL_CHECK_CUDA(cudaGetLastError());
    return OCRL_OK;
  }
  const int grid = p.ntiles < sms ? p.ntiles : sms;
  const size_t smem = 1024 + (size_t)2 * D * 128 + 2 * 64 * 128 + PT_TM * 128 + 2 * PT_TM * 128 + 2 * PT_TM * PT_C * 4 +
                      6 * PT_C * 4 + 12 * 8 + 16;
#define OCRL_LAUNCH_PT(DD)                                                                                       \
  do {                                                                                                           \
    OCRL_CHECK_CUDA(cudaFuncSetAttribute(kv_proj_tc_kernel<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                         (int)smem));                                                            \
    kv_proj_tc_kernel<DD><<<grid, PT_NT, smem, stream>>>(tm_x, tm_w1, tm_w2, tm_wkv, tm_k, tm_v, tm_pos, p); ocrl::count_launch(); \
  } while (0)
  if (D == 64) OCRL_LAUNCH_PT(64);
  else if (D == 128) OCRL_LAUNCH_PT(128);
  else OCRL_LAUNCH_PT(192);
#undef OCRL_LAUNCH_PT
  OCRL_CHECK_CUDA(cudaGetLastError());
  return OCRL_OK;
}

}  // namespace ocrl

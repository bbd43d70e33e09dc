// Device helpers shared by the fused forward and backward iteration kernels: small distributed
// matrix-vector building blocks over the K slot vectors held in shared memory.
#pragma once
#include "common.cuh"

namespace ocrl {

// Layout of the per-(image, iteration) state the forward keeps for the backward (floats).
struct SavedLayout {
  int K, D, H;
  __host__ __device__ SavedLayout(int k, int d, int h) : K(k), D(d), H(h) {}
  __host__ __device__ int off_h() const { return 0; }            // slots entering the iteration [K][D]
  __host__ __device__ int off_q() const { return K * D; }        // queries                      [K][D]
  __host__ __device__ int off_u() const { return 2 * K * D; }    // normalised updates           [K][D]
  __host__ __device__ int off_r() const { return 3 * K * D; }    // GRU reset gate               [K][D]
  __host__ __device__ int off_z() const { return 4 * K * D; }    // GRU update gate              [K][D]
  __host__ __device__ int off_n() const { return 5 * K * D; }    // GRU candidate                [K][D]
  __host__ __device__ int off_ghn() const { return 6 * K * D; }  // W_hn h + b_hn                [K][D]
  __host__ __device__ int off_hp() const { return 7 * K * D; }   // GRU output h'                [K][D]
  __host__ __device__ int off_pre() const { return 8 * K * D; }  // MLP pre-activation           [K][H]
  __host__ __device__ int off_s() const { return 8 * K * D + K * H; }  // sum_n (a+eps)            [K]
  __host__ __device__ int stride() const { return (8 * K * D + K * H + K + 3) & ~3; }  // 16-byte multiple
};

struct IterFwdArgs {
  const void* k;
  const void* v;
  const float* slots0;
  ocrl_sa_weights w;
  float* slots_out;
  float* attn_out;  // may be null
  float* saved;     // may be null: [B][T][SavedLayout::stride()]
  int B, N, D, H, K, T, CL;
  float eps, ln_eps;
  long long* trace;  // optional: clock64() at phase boundaries of CTA 0 (development aid), else null
  const __nv_bfloat16* wb16;  // optional bf16 copies [wq | w_ih | w_hh | w1 | w2] for the tensor-core slot update
  void* workspace = nullptr;         // caller's fwd workspace (ocrl_sa_query_workspace) and its size
  size_t workspace_bytes = 0;
  const uint32_t* wprep = nullptr;   // tcgen05 kernel: prepared weight words / folded-LayerNorm constants (inside the workspace)
  const float* wprep_consts = nullptr;
  int max_clusters = 0;  // ocrl_sa_launch_opts: cap on the resident clusters of the persistent kernels (0 = launcher's choice)
  int lanes = 0;         // ocrl_sa_launch_opts: images in flight per cluster (0 = default)
  int prepared = 0;      // ocrl_sa_launch_opts: the workspace already holds the prepared weights
  // factored pass (ocrl_sa_iter_fwd_xhat): the normalised tokens x^ [B,N,F] bf16 and the projections they stand for
  const void* xhat = nullptr;
  const float* wk = nullptr;  // project_k.weight [D,F]
  const float* wv = nullptr;  // project_v.weight [D,F]
  int F = 0;
};

// out[j*ldo + out_off + row] = dot(W[row0+row, 0:L], vec[j, 0:L]) for row < nrows, j < KP.
// One warp handles RB rows at a time; weights stream from global (L2-resident, shared by the batch).
template <int KP, int RB>
__device__ __forceinline__ void rows_dot(const float* __restrict__ W, int L, int row0, int nrows,
                                         const float* vec, float* out, int ldo, int out_off, int warp,
                                         int lane, int nwarps) {
  constexpr int NV = RB * KP;
  const int nb = (nrows + RB - 1) / RB;
  for (int b = warp; b < nb; b += nwarps) {
    float acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = 0.f;
    for (int c = 0; c < L / 64; ++c) {
      float2 wv[RB];
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const int row = b * RB + r;
        wv[r] = (row < nrows)
                    ? __ldg(reinterpret_cast<const float2*>(W + (size_t)(row0 + row) * L + 64 * c + 2 * lane))
                    : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < KP; ++j) {
        const float2 x = *reinterpret_cast<const float2*>(vec + j * L + 64 * c + 2 * lane);
#pragma unroll
        for (int r = 0; r < RB; ++r) acc[r * KP + j] = fmaf(wv[r].x, x.x, fmaf(wv[r].y, x.y, acc[r * KP + j]));
      }
    }
    int base;
    xreduce<NV>(acc, lane, base);
    if (XReduce<NV, 16>::primary(lane)) {
#pragma unroll
      for (int i = 0; i < XReduce<NV, 16>::kFinal; ++i) {
        const int idx = base + i;
        const int r = idx / KP, j = idx % KP;
        const int row = b * RB + r;
        if (row < nrows) out[j * ldo + out_off + row] = acc[i];
      }
    }
  }
}

// Several rows_dot jobs (same row length 64*NCHL and row count) flattened into one batch list so that
// all warps stay busy.  A warp takes two batches at a time and issues all of their weight loads
// before the first FMA, so one L2 round trip is paid per pair instead of one per 64-feature chunk.
// JobFn: void(int job, const float*& W, int& row0, const float*& vec, float*& out)
template <int KP, int RB, int NCHL, typename JobFn>
__device__ __forceinline__ void rows_dot_jobs(int njobs, int nrows, int ldo, JobFn job_of, int warp, int lane,
                                              int nwarps) {
  constexpr int NV = RB * KP;
  constexpr int L = 64 * NCHL;
  const int nb = (nrows + RB - 1) / RB;
  const int total = njobs * nb;
  auto load = [&](int b, float2(&dst)[RB][NCHL]) {
    const float* W; const float* vec; float* out; int row0;
    job_of(b / nb, W, row0, vec, out);
    const int lb = b % nb;
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      const int row = lb * RB + r;
#pragma unroll
      for (int c = 0; c < NCHL; ++c)
        dst[r][c] = (row < nrows)
                        ? __ldg(reinterpret_cast<const float2*>(W + (size_t)(row0 + row) * L + 64 * c + 2 * lane))
                        : make_float2(0.f, 0.f);
    }
  };
  auto compute = [&](int b, const float2(&wv)[RB][NCHL]) {
    const float* W; const float* vec; float* out; int row0;
    job_of(b / nb, W, row0, vec, out);
    const int lb = b % nb;
    float acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = 0.f;
#pragma unroll
    for (int c = 0; c < NCHL; ++c) {
#pragma unroll
      for (int j = 0; j < KP; ++j) {
        const float2 x = *reinterpret_cast<const float2*>(vec + j * L + 64 * c + 2 * lane);
#pragma unroll
        for (int r = 0; r < RB; ++r)
          acc[r * KP + j] = fmaf(wv[r][c].x, x.x, fmaf(wv[r][c].y, x.y, acc[r * KP + j]));
      }
    }
    int base;
    xreduce<NV>(acc, lane, base);
    if (XReduce<NV, 16>::primary(lane)) {
#pragma unroll
      for (int i = 0; i < XReduce<NV, 16>::kFinal; ++i) {
        const int idx = base + i;
        const int r = idx / KP, j = idx % KP;
        const int row = lb * RB + r;
        if (row < nrows) out[j * ldo + row] = acc[i];
      }
    }
  };
#pragma unroll 1
  for (int b0 = 2 * warp; b0 < total; b0 += 2 * nwarps) {
    float2 wa[RB][NCHL], wb[RB][NCHL];
    const bool two = (b0 + 1 < total);
    load(b0, wa);
    if (two) load(b0 + 1, wb);
    compute(b0, wa);
    if (two) compute(b0 + 1, wb);
  }
}

// Descriptor form of rows_dot_jobs for a compact, NOT inlined routine (keeps the kernels' code size
// -- and with it instruction-cache misses in the once-per-iteration slot update -- small).
// job j: W = (j < split ? W0 : W1), vec = (j < split ? vec0 : vec1), rows [row_base + (j % split_mod) * row_stride, +nrows),
//        out = out0 + j * out_stride.
struct DotDesc {
  const float* W0; const float* W1;
  const float* vec0; const float* vec1;
  float* out0;
  int njobs, split, split_mod, row_base, row_stride, nrows, out_stride, ldo;
};

// GW = true: weights in global memory (read-only path); false: weights staged in shared memory.
template <int KP, int RB, int NCHL, int NB, bool GW = true>
__device__ __noinline__ void rows_dot_desc(const DotDesc d, int warp, int lane, int nwarps) {
  constexpr int NV = RB * KP;
  constexpr int L = 64 * NCHL;
  const int nb = (d.nrows + RB - 1) / RB;
  const int total = d.njobs * nb;
#pragma unroll 1
  for (int b0 = NB * warp; b0 < total; b0 += NB * nwarps) {
    float2 w[NB][RB][NCHL];
#pragma unroll
    for (int q = 0; q < NB; ++q) {
      const int b = b0 + q;
      const int job = b / nb, lb = b % nb;
      const float* W = (job < d.split) ? d.W0 : d.W1;
      const int row0 = d.row_base + (job % d.split_mod) * d.row_stride;
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const int row = lb * RB + r;
#pragma unroll
        for (int c = 0; c < NCHL; ++c)
        {
          const float2* wp = reinterpret_cast<const float2*>(W + (size_t)(row0 + row) * L + 64 * c + 2 * lane);
          const bool ok = (b < total && row < d.nrows);
          if constexpr (GW) w[q][r][c] = ok ? __ldg(wp) : make_float2(0.f, 0.f);
          else w[q][r][c] = ok ? *wp : make_float2(0.f, 0.f);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < NB; ++q) {
      const int b = b0 + q;
      if (b < total) {
        const int job = b / nb, lb = b % nb;
        const float* vec = (job < d.split) ? d.vec0 : d.vec1;
        float* out = d.out0 + job * d.out_stride;
        float acc[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) acc[i] = 0.f;
#pragma unroll
        for (int c = 0; c < NCHL; ++c) {
#pragma unroll
          for (int j = 0; j < KP; ++j) {
            const float2 x = *reinterpret_cast<const float2*>(vec + j * L + 64 * c + 2 * lane);
#pragma unroll
            for (int r = 0; r < RB; ++r)
              acc[r * KP + j] = fmaf(w[q][r][c].x, x.x, fmaf(w[q][r][c].y, x.y, acc[r * KP + j]));
          }
        }
        int base;
        xreduce<NV>(acc, lane, base);
        if (XReduce<NV, 16>::primary(lane)) {
#pragma unroll
          for (int i = 0; i < XReduce<NV, 16>::kFinal; ++i) {
            const int idx = base + i;
            const int r = idx / KP, j = idx % KP;
            const int row = lb * RB + r;
            if (row < d.nrows) out[j * d.ldo + row] = acc[i];
          }
        }
      }
    }
  }
}

// runtime row length (multiple of 64, <= 256)
template <int KP, int RB, int NB, bool GW = true>
__device__ __forceinline__ void rows_dot_desc_len(const DotDesc& d, int L, int warp, int lane, int nwarps) {
  switch (L / 64) {
    case 1: rows_dot_desc<KP, RB, 1, NB, GW>(d, warp, lane, nwarps); break;
    case 2: rows_dot_desc<KP, RB, 2, NB, GW>(d, warp, lane, nwarps); break;
    case 3: rows_dot_desc<KP, RB, 3, NB, GW>(d, warp, lane, nwarps); break;
    default: rows_dot_desc<KP, RB, 4, NB, GW>(d, warp, lane, nwarps); break;
  }
}

// LayerNorm of `rows` rows of compile-time length D in shared memory, affine parameters in shared
// memory too; every lane keeps its D/32 features in registers (one pass over the row).
template <int D>
__device__ __forceinline__ void ln_rows_fast(const float* src, const float* gw_s, const float* gb_s, float* dst,
                                             int rows, float eps, int warp, int lane, int nwarps) {
  constexpr int NCH = D / 64;
  for (int j = warp; j < rows; j += nwarps) {
    float2 x[NCH];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      x[c] = *reinterpret_cast<const float2*>(src + j * D + 64 * c + 2 * lane);
      s += x[c].x + x[c].y;
    }
    const float mean = warp_sum(s) * (1.f / D);
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      x[c].x -= mean;
      x[c].y -= mean;
      q = fmaf(x[c].x, x[c].x, fmaf(x[c].y, x[c].y, q));
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / D) + eps);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const float2 g = *reinterpret_cast<const float2*>(gw_s + 64 * c + 2 * lane);
      const float2 b = *reinterpret_cast<const float2*>(gb_s + 64 * c + 2 * lane);
      *reinterpret_cast<float2*>(dst + j * D + 64 * c + 2 * lane) =
          make_float2(x[c].x * rstd * g.x + b.x, x[c].y * rstd * g.y + b.y);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Tensor-core form of the slot-update matvecs (bf16 mode).
//   out[j][row] = sum_l W[row][l] * vec[j][l]      W: bf16 in global memory (L2-resident), 16 rows per MMA
//   vec is given as a bf16 hi/lo pair (hi + lo reproduces the fp32 value to ~2^-17), so only the weights
//   carry bf16 rounding.  A fragments are 128-bit global loads (the contraction index is permuted so that a
//   lane's 8 consecutive weights feed two k16 steps), B fragments are hoisted for the whole call.
// Jobs are flattened over the warps: job u covers rows [row_base + u*row_stride, +nrows), writes
// out0 + u*out_job_stride.  vhi / vlo: shared memory [KP][LP] bf16 (rows >= K zero).
// ---------------------------------------------------------------------------------------------
template <int KP, int NCH32>
__device__ __noinline__ void rows_mma_jobs(const __nv_bfloat16* __restrict__ W, int row_base, int row_stride,
                                              int njobs, int nrows, const __nv_bfloat16* vhi,
                                              const __nv_bfloat16* vlo, int LP, float* out0, int out_job_stride,
                                              int ldo, int warp, int lane, int nwarps) {
  constexpr int L = 32 * NCH32, NSL = KP / 8;
  const int g8 = lane >> 2, t4 = lane & 3;
  uint2 bh[NSL][NCH32][2], bl[NSL][NCH32][2];
#pragma unroll
  for (int sl = 0; sl < NSL; ++sl)
#pragma unroll
    for (int c = 0; c < NCH32; ++c)
#pragma unroll
      for (int st = 0; st < 2; ++st) {
        const int off = (8 * sl + g8) * LP + 32 * c + 8 * t4 + 4 * st;
        bh[sl][c][st] = *reinterpret_cast<const uint2*>(vhi + off);
        bl[sl][c][st] = *reinterpret_cast<const uint2*>(vlo + off);
      }
  const int tpj = (nrows + 15) / 16;
  const int total = njobs * tpj;
  uint4 a0[NCH32], a1[NCH32];
  auto load_tile = [&](int u, uint4(&x0)[NCH32], uint4(&x1)[NCH32]) {
    const int job = u / tpj, mt = u % tpj;
    const int r0 = min(mt * 16 + g8, nrows - 1), r1 = min(mt * 16 + g8 + 8, nrows - 1);
    const __nv_bfloat16* base = W + (size_t)(row_base + job * row_stride) * L + 8 * t4;
#pragma unroll
    for (int c = 0; c < NCH32; ++c) {
      x0[c] = __ldg(reinterpret_cast<const uint4*>(base + (size_t)r0 * L + 32 * c));
      x1[c] = __ldg(reinterpret_cast<const uint4*>(base + (size_t)r1 * L + 32 * c));
    }
  };
  int u = warp;
  if (u < total) load_tile(u, a0, a1);
#pragma unroll 1
  for (; u < total; u += nwarps) {
    uint4 n0[NCH32], n1[NCH32];
    const bool more = (u + nwarps) < total;
    if (more) load_tile(u + nwarps, n0, n1);
    float acc[NSL][4];
#pragma unroll
    for (int sl = 0; sl < NSL; ++sl)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[sl][i] = 0.f;
#pragma unroll
    for (int c = 0; c < NCH32; ++c) {
      const uint32_t f0[4] = {a0[c].x, a1[c].x, a0[c].y, a1[c].y};
      const uint32_t f1[4] = {a0[c].z, a1[c].z, a0[c].w, a1[c].w};
#pragma unroll
      for (int sl = 0; sl < NSL; ++sl) {
        mma_bf16_16816(acc[sl], f0, bh[sl][c][0].x, bh[sl][c][0].y);
        mma_bf16_16816(acc[sl], f0, bl[sl][c][0].x, bl[sl][c][0].y);
        mma_bf16_16816(acc[sl], f1, bh[sl][c][1].x, bh[sl][c][1].y);
        mma_bf16_16816(acc[sl], f1, bl[sl][c][1].x, bl[sl][c][1].y);
      }
    }
    const int job = u / tpj, mt = u % tpj;
    float* out = out0 + job * out_job_stride;
    const int r0 = mt * 16 + g8, r1 = r0 + 8;
#pragma unroll
    for (int sl = 0; sl < NSL; ++sl) {
      const int j0 = 8 * sl + 2 * t4;
      if (r0 < nrows) { out[j0 * ldo + r0] = acc[sl][0]; out[(j0 + 1) * ldo + r0] = acc[sl][1]; }
      if (r1 < nrows) { out[j0 * ldo + r1] = acc[sl][2]; out[(j0 + 1) * ldo + r1] = acc[sl][3]; }
    }
    if (more) {
#pragma unroll
      for (int c = 0; c < NCH32; ++c) { a0[c] = n0[c]; a1[c] = n1[c]; }
    }
  }
}

template <int KP>
__device__ __forceinline__ void rows_mma_jobs_len(int L, const __nv_bfloat16* W, int row_base, int row_stride, int njobs,
                                                  int nrows, const __nv_bfloat16* vhi, const __nv_bfloat16* vlo, int LP,
                                                  float* out0, int out_job_stride, int ldo, int warp, int lane,
                                                  int nwarps) {
  switch (L / 32) {
    case 2: rows_mma_jobs<KP, 2>(W, row_base, row_stride, njobs, nrows, vhi, vlo, LP, out0, out_job_stride, ldo, warp, lane, nwarps); break;
    case 4: rows_mma_jobs<KP, 4>(W, row_base, row_stride, njobs, nrows, vhi, vlo, LP, out0, out_job_stride, ldo, warp, lane, nwarps); break;
    case 6: rows_mma_jobs<KP, 6>(W, row_base, row_stride, njobs, nrows, vhi, vlo, LP, out0, out_job_stride, ldo, warp, lane, nwarps); break;
    default: rows_mma_jobs<KP, 8>(W, row_base, row_stride, njobs, nrows, vhi, vlo, LP, out0, out_job_stride, ldo, warp, lane, nwarps); break;
  }
}

// fp32 vectors [K][L] in shared memory -> bf16 hi / lo pair [KP][LP] (rows >= K zero)
__device__ __forceinline__ void stage_vec_hilo(const float* src, int K, int KPr, int L, __nv_bfloat16* hi,
                                               __nv_bfloat16* lo, int LP, int tid, int nthreads) {
  for (int e = tid; e < KPr * (L / 2); e += nthreads) {
    const int j = e / (L / 2), l2 = e % (L / 2);
    float2 x = make_float2(0.f, 0.f);
    if (j < K) x = *reinterpret_cast<const float2*>(src + j * L + 2 * l2);
    const __nv_bfloat162 h = __floats2bfloat162_rn(x.x, x.y);
    const __nv_bfloat162 l = __floats2bfloat162_rn(x.x - __low2float(h), x.y - __high2float(h));
    *reinterpret_cast<__nv_bfloat162*>(hi + j * LP + 2 * l2) = h;
    *reinterpret_cast<__nv_bfloat162*>(lo + j * LP + 2 * l2) = l;
  }
}

// LayerNorm of `rows` rows of length L held in shared memory (warp per row).
__device__ __forceinline__ void ln_rows(const float* src, const float* __restrict__ gw,
                                        const float* __restrict__ gb, float* dst, int rows, int L, float eps,
                                        int warp, int lane, int nwarps) {
  for (int j = warp; j < rows; j += nwarps) {
    float s = 0.f;
    for (int d = lane; d < L; d += 32) s += src[j * L + d];
    const float mean = warp_sum(s) / (float)L;
    float q = 0.f;
    for (int d = lane; d < L; d += 32) {
      const float t = src[j * L + d] - mean;
      q = fmaf(t, t, q);
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)L + eps);
    for (int d = lane; d < L; d += 32) dst[j * L + d] = (src[j * L + d] - mean) * rstd * __ldg(gw + d) + __ldg(gb + d);
  }
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

}  // namespace ocrl

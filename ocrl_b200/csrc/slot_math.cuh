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

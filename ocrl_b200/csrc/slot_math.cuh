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

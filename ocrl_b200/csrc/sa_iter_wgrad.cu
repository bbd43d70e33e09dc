// Weight gradients of the T-iteration loop from the factor record of the cluster kernel (FLog, sa_iter_bwd.cuh).
//
// Every matrix gradient is a product over all M = B T K recorded rows,  dW[r][c] = sum_m A[m][r] Bf[m][c]:
//     dW2  = dsn^T relu(pre)      dW1  = dpre^T mhat      dWq = dq^T shat
//     dW_ih = [dr dz dn]^T u      dW_hh = [dr dz dn*r]^T h
// (u, h, pre from the forward's saved state), every bias gradient a column sum of an A factor, and the LayerNorm
// gradients sums of the per-(image, iteration) vectors.  wgrad_gemm_kernel forms 16 x 64 tiles of the nine products for
// one of WG_SPLITS shares of the rows, wgrad_finish_kernel adds the shares in order and does the column sums: fixed
// summation order, no atomics -- the result does not depend on the grid.
#include "sa_iter_bwd.cuh"

namespace ocrl {
namespace wgrad {

constexpr int NT = 256, ROWS = 16, COLS = 64, KC = 32, WG_SPLITS = 4, NPROD = 9;

struct Prod {
  int rows, cols;       // of this product
  int a_fact;           // FLog factor holding A
  int b_fact;           // FLog factor holding B, or -1: from the saved state at b_off with row pitch b_pitch
  int b_off, b_pitch, b_relu;
  int out_off;          // offset of the product inside the concatenated gradient buffer
  int tile0;            // first tile index
};

struct Args {
  const float* flog;
  const float* saved;
  float* part;          // [WG_SPLITS][total] split partial sums
  Prod prod[NPROD];
  int ntiles, total;
  int B, T, K, D, H;
  ocrl_sa_weight_grads dw;
};

__global__ void __launch_bounds__(NT) wgrad_gemm_kernel(const Args a) {
  // rows of a chunk = NB whole (image, iteration) records x K slot rows; their offsets inside a record are the same for
  // every chunk, so the index arithmetic is done once per block
  __shared__ float as[KC][ROWS];
  __shared__ float bs[KC][COLS];
  __shared__ int offa[KC], offb[KC];
  const int tid = threadIdx.x;
  int p = 0;
#pragma unroll
  for (int i = 1; i < NPROD; ++i)
    if ((int)blockIdx.x >= a.prod[i].tile0) p = i;
  const Prod P = a.prod[p];
  const int ctiles = (P.cols + COLS - 1) / COLS;
  const int tile = blockIdx.x - P.tile0;
  const int r0 = (tile / ctiles) * ROWS, c0 = (tile % ctiles) * COLS;
  const FLog FL(a.K, a.D, a.H);
  const SavedLayout SL(a.K, a.D, a.H);
  const int K = a.K, NB = KC / K, rows = NB * K;  // records and rows per chunk
  const int BT = a.B * a.T;
  const int per = (BT + WG_SPLITS - 1) / WG_SPLITS;
  const int bt_begin = blockIdx.y * per, bt_end = min(BT, bt_begin + per);
  const size_t a_stride = FL.stride(), b_stride = (P.b_fact >= 0) ? FL.stride() : (size_t)SL.stride();
  const float* a_base = a.flog + FL.fact(P.a_fact) + r0;
  const float* b_base = (P.b_fact >= 0 ? a.flog + FL.fact(P.b_fact) : a.saved + P.b_off) + c0;
  if (tid < rows) {
    const int blk = tid / K, j = tid % K;
    offa[tid] = (int)(blk * a_stride + (size_t)j * FL.L);
    offb[tid] = (int)(blk * b_stride + (size_t)j * (P.b_fact >= 0 ? FL.L : P.b_pitch));
  }
  __syncthreads();
  const int c = tid % COLS, rq = tid / COLS;  // rows r0 + 4 rq .. + 3
  const int la_kk = tid / ROWS, la_rr = tid % ROWS;  // this thread's A elements: rows la_kk, la_kk + 16
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int bt0 = bt_begin; bt0 < bt_end; bt0 += NB) {
    const int live = min(NB, bt_end - bt0) * K;  // rows of this chunk that exist
    const float* ap = a_base + (size_t)bt0 * a_stride;
    const float* bp = b_base + (size_t)bt0 * b_stride;
#pragma unroll
    for (int h = 0; h < KC * ROWS / NT; ++h) {
      const int kk = la_kk + h * (NT / ROWS);
      as[kk][la_rr] = (kk < live && r0 + la_rr < P.rows) ? __ldg(ap + offa[kk] + la_rr) : 0.f;
    }
#pragma unroll
    for (int h = 0; h < KC * COLS / NT; ++h) {
      const int kk = rq + h * (NT / COLS);
      float v = (kk < live && c0 + c < P.cols) ? __ldg(bp + offb[kk] + c) : 0.f;
      bs[kk][c] = P.b_relu ? fmaxf(v, 0.f) : v;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < KC; ++kk) {
      const float bv = bs[kk][c];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(as[kk][rq * 4 + i], bv, acc[i]);
    }
    __syncthreads();
  }
  float* out = a.part + (size_t)blockIdx.y * a.total + P.out_off;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + rq * 4 + i;
    if (r < P.rows && c0 + c < P.cols) out[(size_t)r * P.cols + c0 + c] = acc[i];
  }
}

// blocks [0, nmat): the matrix gradients (sum of the split partials);  blocks [nmat, nmat + nvec): 64 columns of one
// vector gradient each -- 16 row groups per block sum every 16th recorded row, then a fixed-order combine
__global__ void __launch_bounds__(1024) wgrad_finish_kernel(const Args a, int nmat) {
  const int tid = threadIdx.x;
  const int D = a.D, H = a.H, K = a.K;
  if ((int)blockIdx.x < nmat) {
    const int i = blockIdx.x * 1024 + tid;
    if (i >= a.total) return;
    float s = 0.f;
#pragma unroll
    for (int sp = 0; sp < WG_SPLITS; ++sp) s += a.part[(size_t)sp * a.total + i];
    // concatenated layout: W2 [D][H] | W1 [H][D] | W_ih [3D][D] | W_hh [3D][D] | Wq [D][D]
    int o = i;
    if (o < D * H) { a.dw.w2[o] = s; return; }
    o -= D * H;
    if (o < H * D) { a.dw.w1[o] = s; return; }
    o -= H * D;
    if (o < 3 * D * D) { a.dw.w_ih[o] = s; return; }
    o -= 3 * D * D;
    if (o < 3 * D * D) { a.dw.w_hh[o] = s; return; }
    o -= 3 * D * D;
    a.dw.wq[o] = s;
    return;
  }
  __shared__ float red[16][64];
  const FLog FL(K, D, H);
  // vector outputs, 64 columns per block: b2 [D] | b1 [H] | b_ih [3D] | b_hh [3D] | ln_m_w | ln_m_b | ln_s_w | ln_s_b [D each]
  int col = (blockIdx.x - nmat) * 64 + (tid % 64);
  const int g = tid / 64;  // 0..15
  float* dst = nullptr;
  int fact = -1, vec = -1, len = D;
  if (col < D) { dst = a.dw.b2; fact = FLog::DSN; }
  else if ((col -= D) < H) { dst = a.dw.b1; fact = FLog::DPRE; len = H; }
  else if ((col -= H) < 3 * D) { dst = a.dw.b_ih + (col / D) * D; fact = FLog::DR + col / D; col %= D; }
  else if ((col -= 3 * D) < 3 * D) { dst = a.dw.b_hh + (col / D) * D; fact = (col / D == 2) ? (int)FLog::DNR : FLog::DR + col / D; col %= D; }
  else if ((col -= 3 * D) < 4 * D) {
    vec = col / D;
    dst = vec == 0 ? a.dw.ln_mlp_w : vec == 1 ? a.dw.ln_mlp_b : vec == 2 ? a.dw.ln_slots_w : a.dw.ln_slots_b;
    col %= D;
  } else {
    col = -1;
  }
  float s = 0.f;
  if (col >= 0 && col < len) {
    if (fact >= 0) {
      const int M = a.B * a.T * K;
      for (int m = g; m < M; m += 16)
        s += __ldg(a.flog + (size_t)(m / K) * FL.stride() + FL.fact(fact) + (size_t)(m % K) * FL.L + col);
    } else {
      const int M = a.B * a.T;
      for (int m = g; m < M; m += 16) s += __ldg(a.flog + (size_t)m * FL.stride() + FL.vec(vec) + col);
    }
  }
  red[g][tid % 64] = s;
  __syncthreads();
  if (g == 0 && col >= 0 && col < len) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) t += red[i][tid % 64];
    dst[col] = t;
  }
}

}  // namespace wgrad

size_t sa_iter_wgrad_part_floats(const ocrl_sa_dims* d) {
  const size_t D = d->D, H = d->H_mlp;
  return (size_t)wgrad::WG_SPLITS * (2 * D * H + 7 * D * D);
}

int sa_iter_wgrad_launch(const ocrl_sa_dims* d, const float* flog, const float* saved, float* part,
                         const ocrl_sa_weight_grads* dw, cudaStream_t stream) {
  using namespace wgrad;
  const int D = d->D, H = d->H_mlp, K = d->K;
  const SavedLayout SL(K, D, H);
  Args a;
  a.flog = flog; a.saved = saved; a.part = part; a.dw = *dw;
  a.B = d->B; a.T = d->T; a.K = K; a.D = D; a.H = H;
  int tile = 0, off = 0;
  auto add = [&](int i, int rows, int cols, int a_fact, int b_fact, int b_off, int b_pitch, int b_relu) {
    Prod& P = a.prod[i];
    P.rows = rows; P.cols = cols; P.a_fact = a_fact; P.b_fact = b_fact; P.b_off = b_off; P.b_pitch = b_pitch;
    P.b_relu = b_relu; P.out_off = off; P.tile0 = tile;
    off += rows * cols;
    tile += ((rows + ROWS - 1) / ROWS) * ((cols + COLS - 1) / COLS);
  };
  add(0, D, H, FLog::DSN, -1, SL.off_pre(), H, 1);          // W2
  add(1, H, D, FLog::DPRE, FLog::MHAT, 0, 0, 0);            // W1
  add(2, D, D, FLog::DR, -1, SL.off_u(), D, 0);             // W_ih, rows r | z | n
  add(3, D, D, FLog::DZ, -1, SL.off_u(), D, 0);
  add(4, D, D, FLog::DN, -1, SL.off_u(), D, 0);
  add(5, D, D, FLog::DR, -1, SL.off_h(), D, 0);             // W_hh
  add(6, D, D, FLog::DZ, -1, SL.off_h(), D, 0);
  add(7, D, D, FLog::DNR, -1, SL.off_h(), D, 0);
  add(8, D, D, FLog::DQ, FLog::SHAT, 0, 0, 0);              // Wq
  a.ntiles = tile;
  a.total = off;
  wgrad_gemm_kernel<<<dim3(tile, WG_SPLITS), NT, 0, stream>>>(a);
  ocrl::count_launch();
  OCRL_CHECK_CUDA(cudaGetLastError());
  const int nmat = (off + 1023) / 1024;
  const int nvec = (D + H + 6 * D + 4 * D + 63) / 64 + 4;  // segments are not 64-aligned in general: a few spare blocks
  wgrad_finish_kernel<<<nmat + nvec, 1024, 0, stream>>>(a, nmat);
  ocrl::count_launch();
  OCRL_CHECK_CUDA(cudaGetLastError());
  return OCRL_OK;
}

}  // namespace ocrl

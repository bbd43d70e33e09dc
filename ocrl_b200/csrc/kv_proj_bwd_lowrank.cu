// Backward of norm_inputs + project_k / project_v (autograd of ocrs/common/slot_attn.py:54-61) straight from the
// rank-(2 T K) coefficients of the fused iteration backward: dk and dv never exist in memory.
//
// The iteration backward leaves, per token n and iteration t, the 2K coefficients  dl^t_nj  and  w^t_nj = a^t_nj + eps
// (coef [B][N][T][2K]) with
//     dk_n = sum_{t,j} dl^t_nj q^t_j              dv_n = sum_{t,j} w^t_nj gm^t_j        (gm = dU / S, [B][T][K][D])
// so with xh_n = LayerNorm(x_n) (64 wide), k = s W_k xh, v = W_v xh and the per-image matrix
//     M[(t,0,j)] = s W_k^T q^t_j ,   M[(t,1,j)] = W_v^T gm^t_j          (R = 2 T K rows of 64)
// the input gradient is   dxh_n = sum_r coef_nr M_r   and the weight gradients are
//     dW_k = s sum_{b,t,j} q^t_j (x) G[(t,0,j)] ,  dW_v = sum_{b,t,j} gm^t_j (x) G[(t,1,j)] ,   G = coef^T xh  (R x 64 per image).
// Per token that is 2 R 64 multiply-adds instead of 4 D 64, and the traffic is x + coef in, dx out (0.66 KB per
// token at K = 6, T = 3) instead of the 2 x 1.5 KB fp32 round trip of dk, dv.  fp32 FFMA throughout (the 2e-4 gradient
// tolerance of the parity mode rules out bf16 / tf32 operands); four launches:
//   lowrank_m_kernel     M per image                                        (B T blocks)
//   lowrank_main_kernel  dxh, LayerNorm backward, dx; partial G, d gamma, d beta per (image, token range)
//   lowrank_dw_kernel    split-K partials of dW_k, dW_v from q / gm and the partial G
//   lowrank_finish_kernel  dW_k, dW_v, d gamma, d beta                        (fixed summation order, no atomics)
#include "slot_math.cuh"

namespace ocrl {
namespace lowrank {

constexpr int C = 64;    // input width of the projection (C_in)
constexpr int TT = 64;   // tokens per tile
constexpr int NT = 256;

struct Args {
  const float* x;        // [B][N][64] input of norm_inputs
  const float* coef;     // [B][N][R]
  const float* gm;       // [B][T][K][D]
  const float* saved;    // forward state (queries)
  const float* ln_w; const float* ln_b;
  const float* wk; const float* wv;  // [D][64]
  float* dx;             // [B][N][64]
  float* M;              // [B][R][64]
  float* Gp;             // [B*S][R][64] partial coef^T xh
  float* lnp;            // [B*S][2][64] partial d gamma, d beta
  float* dWp;            // [DW_SPLITS][2][D][64] split-K partials of dW_k, dW_v
  float* dwk; float* dwv; float* d_ln_w; float* d_ln_b;
  int B, N, D, H, K, T, R, S, chunk;
  float ln_eps, kscale;
};

// M[b][t*2K + which*K + j][c] = sum_d W[d][c] vec[d],  vec = s q^t_j (which = 0) or gm^t_j (which = 1)
__global__ void __launch_bounds__(NT) lowrank_m_kernel(const Args a) {
  extern __shared__ float vec[];  // [2K][D]
  const int b = blockIdx.x / a.T, t = blockIdx.x % a.T;
  const int tid = threadIdx.x, K = a.K, D = a.D;
  const SavedLayout SL(K, D, a.H);
  const float* q = a.saved + ((size_t)b * a.T + t) * SL.stride() + SL.off_q();
  const float* g = a.gm + ((size_t)b * a.T + t) * K * D;
  for (int e = tid; e < K * D; e += NT) {
    vec[e] = a.kscale * __ldg(q + e);
    vec[K * D + e] = __ldg(g + e);
  }
  __syncthreads();
  const int c = tid % C;
  for (int row = tid / C; row < 2 * K; row += NT / C) {
    const float* W = (row < K) ? a.wk : a.wv;
    const float* vv = vec + row * D;
    float acc = 0.f;
    for (int d = 0; d < D; ++d) acc = fmaf(__ldg(W + (size_t)d * C + c), vv[d], acc);
    a.M[((size_t)b * a.R + t * 2 * K + row) * C + c] = acc;
  }
}

// One CTA per (image, token range).  RQ: accumulators of G per thread (R <= 4 RQ).
template <int RQ>
__global__ void __launch_bounds__(NT, 2) lowrank_main_kernel(const Args a) {
  extern __shared__ __align__(16) float sm[];
  const int R = a.R, RS = (R + 3) & ~3;  // coefficient row stride in shared memory (float4 loads)
  float* xs = sm;                  // [TT][64] x, then the normalised x
  float* ds = xs + TT * C;         // [TT][64] dxh
  float* cs = ds + TT * C;         // [TT][RS] coefficients
  float* Ms = cs + TT * RS;        // [R][64]
  float* rstd_s = Ms + R * C;      // [TT]
  float* red = rstd_s + TT;        // [8][2][64] d gamma / d beta of the warps
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int img = blockIdx.x / a.S, part = blockIdx.x % a.S;
  const int n_begin = part * a.chunk, n_end = min(a.N, n_begin + a.chunk);

  for (int e = tid; e < R * C; e += NT) Ms[e] = __ldg(a.M + (size_t)img * R * C + e);
  const int gc = tid % C, rg = tid / C;  // G: this thread owns column gc of rows rg*RQ .. rg*RQ + RQ - 1
  const float gam_c = __ldg(a.ln_w + gc), bet_c = __ldg(a.ln_b + gc);
  const float g0 = __ldg(a.ln_w + lane), g1 = __ldg(a.ln_w + lane + 32);
  float G[RQ];
#pragma unroll
  for (int i = 0; i < RQ; ++i) G[i] = 0.f;
  float dgam0 = 0.f, dgam1 = 0.f, dbet0 = 0.f, dbet1 = 0.f;  // columns lane, lane + 32 of this warp's tokens
  const int tn = tid / 16, tc = tid % 16;  // dxh micro-tile: tokens 4 tn .. +3, columns 4 tc .. +3
  __syncthreads();

  for (int n0 = n_begin; n0 < n_end; n0 += TT) {
    const int valid = min(TT, n_end - n0);
    const float* xg = a.x + ((size_t)img * a.N + n0) * C;
    const float* cg = a.coef + ((size_t)img * a.N + n0) * R;
    for (int e = tid; e < TT * C / 4; e += NT) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e / (C / 4) < valid) v = __ldg(reinterpret_cast<const float4*>(xg) + e);
      reinterpret_cast<float4*>(xs)[e] = v;
    }
    for (int e = tid; e < TT * RS; e += NT) {  // rows past the range read as zero: they add nothing to G or dxh
      const int n = e / RS, r = e % RS;
      cs[e] = (n < valid && r < R) ? __ldg(cg + (size_t)n * R + r) : 0.f;
    }
    __syncthreads();
    // LayerNorm statistics, warp per token
    for (int n = warp; n < TT; n += NT / 32) {
      float x0 = xs[n * C + lane], x1 = xs[n * C + lane + 32];
      const float mean = warp_sum(x0 + x1) * (1.f / C);
      x0 -= mean;
      x1 -= mean;
      const float rstd = rsqrtf(warp_sum(x0 * x0 + x1 * x1) * (1.f / C) + a.ln_eps);
      const bool live = n < valid;
      xs[n * C + lane] = live ? x0 * rstd : 0.f;
      xs[n * C + lane + 32] = live ? x1 * rstd : 0.f;
      if (lane == 0) rstd_s[n] = rstd;
    }
    // dxh = coef M  (4 x 4 outputs per thread)
    {
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      for (int r = 0; r < R; ++r) {
        const float4 m4 = *reinterpret_cast<const float4*>(Ms + r * C + 4 * tc);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float cv = cs[(4 * tn + i) * RS + r];
          acc[i][0] = fmaf(cv, m4.x, acc[i][0]);
          acc[i][1] = fmaf(cv, m4.y, acc[i][1]);
          acc[i][2] = fmaf(cv, m4.z, acc[i][2]);
          acc[i][3] = fmaf(cv, m4.w, acc[i][3]);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4*>(ds + (4 * tn + i) * C + 4 * tc) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
    __syncthreads();
    // G += coef^T xh  (xh = xn gamma + beta; rows past the range have zero coefficients)
    for (int n = 0; n < TT; ++n) {
      const float xh = fmaf(xs[n * C + gc], gam_c, bet_c);
      const float* cr = cs + n * RS + rg * RQ;
#pragma unroll
      for (int i = 0; i < RQ; ++i)
        if (rg * RQ + i < R) G[i] = fmaf(cr[i], xh, G[i]);
    }
    // LayerNorm backward, warp per token; d gamma / d beta column sums
    for (int n = warp; n < valid; n += NT / 32) {
      const float d0 = ds[n * C + lane], d1 = ds[n * C + lane + 32];
      const float xn0 = xs[n * C + lane], xn1 = xs[n * C + lane + 32];
      dgam0 = fmaf(d0, xn0, dgam0);
      dgam1 = fmaf(d1, xn1, dgam1);
      dbet0 += d0;
      dbet1 += d1;
      const float y0 = d0 * g0, y1 = d1 * g1;
      const float s1 = warp_sum(y0 + y1) * (1.f / C);
      const float s2 = warp_sum(y0 * xn0 + y1 * xn1) * (1.f / C);
      const float rstd = rstd_s[n];
      float* dst = a.dx + ((size_t)img * a.N + n0 + n) * C;
      dst[lane] = rstd * (y0 - s1 - xn0 * s2);
      dst[lane + 32] = rstd * (y1 - s1 - xn1 * s2);
    }
    __syncthreads();
  }
  float* Gout = a.Gp + (size_t)blockIdx.x * R * C;
#pragma unroll
  for (int i = 0; i < RQ; ++i)
    if (rg * RQ + i < R) Gout[(rg * RQ + i) * C + gc] = G[i];
  red[(warp * 2 + 0) * C + lane] = dgam0;
  red[(warp * 2 + 0) * C + lane + 32] = dgam1;
  red[(warp * 2 + 1) * C + lane] = dbet0;
  red[(warp * 2 + 1) * C + lane + 32] = dbet1;
  __syncthreads();
  if (tid < 2 * C) {
    const int which = tid / C, c = tid % C;
    float s = 0.f;
    for (int w = 0; w < NT / 32; ++w) s += red[(w * 2 + which) * C + c];
    a.lnp[((size_t)blockIdx.x * 2 + which) * C + c] = s;
  }
}

// dW[d][c] = sum_{b,t,j} vec_{b,t,j}[d] * (sum_s G[b][s][t*2K + which*K + j][c]) as a split-K product: block
// (row tile, which, k split) sums its share of the (b, t, j) triples into dWp[split][which][d][c]; lowrank_finish_kernel
// adds the splits in order (and the per-CTA LayerNorm partials): fixed summation order, no atomics.
constexpr int DW_ROWS = 16, DW_KC = 32, DW_SPLITS = 16;
__global__ void __launch_bounds__(NT) lowrank_dw_kernel(const Args a) {
  __shared__ float vs[DW_KC][DW_ROWS];
  __shared__ float gs[DW_KC][C];
  const int tid = threadIdx.x;
  const int which = blockIdx.y, d0 = blockIdx.x * DW_ROWS, K = a.K, D = a.D, T = a.T;
  const SavedLayout SL(K, D, a.H);
  const int total = a.B * T * K;  // (b, t, j) triples
  const int per = (total + DW_SPLITS - 1) / DW_SPLITS;
  const int k_begin = blockIdx.z * per, k_end = min(total, k_begin + per);
  const int c = tid % C, rq = tid / C;  // rows d0 + rq*4 .. +3
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = k_begin; k0 < k_end; k0 += DW_KC) {
    for (int e = tid; e < DW_KC * DW_ROWS; e += NT) {
      const int kk = e / DW_ROWS, dd = e % DW_ROWS, idx = k0 + kk;
      float v = 0.f;
      if (idx < k_end && d0 + dd < D) {
        const int b = idx / (T * K), t = (idx / K) % T, j = idx % K;
        v = (which == 0) ? __ldg(a.saved + ((size_t)b * T + t) * SL.stride() + SL.off_q() + j * D + d0 + dd)
                         : __ldg(a.gm + (((size_t)b * T + t) * K + j) * D + d0 + dd);
      }
      vs[kk][dd] = v;
    }
    for (int e = tid; e < DW_KC * C; e += NT) {
      const int kk = e / C, cc = e % C, idx = k0 + kk;
      float g = 0.f;
      if (idx < k_end) {
        const int b = idx / (T * K), t = (idx / K) % T, j = idx % K;
        const int r = t * 2 * K + which * K + j;
        for (int s = 0; s < a.S; ++s) g += __ldg(a.Gp + (((size_t)b * a.S + s) * a.R + r) * C + cc);
      }
      gs[kk][cc] = g;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < DW_KC; ++kk) {
      const float g = gs[kk][c];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(vs[kk][rq * 4 + i], g, acc[i]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (d0 + rq * 4 + i < D) a.dWp[(((size_t)blockIdx.z * 2 + which) * D + d0 + rq * 4 + i) * C + c] = acc[i];
}

__global__ void __launch_bounds__(NT) lowrank_finish_kernel(const Args a) {
  const int i = blockIdx.x * NT + threadIdx.x;
  const int nw = a.D * C;
  if (i < 2 * nw) {
    const int which = i / nw, e = i % nw;
    float s = 0.f;
#pragma unroll
    for (int sp = 0; sp < DW_SPLITS; ++sp) s += a.dWp[((size_t)sp * 2 + which) * nw + e];
    (which == 0 ? a.dwk : a.dwv)[e] = (which == 0 ? a.kscale : 1.f) * s;
  } else if (i < 2 * nw + 2 * C) {  // d gamma, d beta: sum of the per-CTA partials
    const int which = (i - 2 * nw) / C, c = (i - 2 * nw) % C;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    const int U = a.B * a.S;
    int u = 0;
    for (; u + 3 < U; u += 4) {
      s0 += a.lnp[((size_t)(u + 0) * 2 + which) * C + c];
      s1 += a.lnp[((size_t)(u + 1) * 2 + which) * C + c];
      s2 += a.lnp[((size_t)(u + 2) * 2 + which) * C + c];
      s3 += a.lnp[((size_t)(u + 3) * 2 + which) * C + c];
    }
    for (; u < U; ++u) s0 += a.lnp[((size_t)u * 2 + which) * C + c];
    (which == 0 ? a.d_ln_w : a.d_ln_b)[c] = (s0 + s1) + (s2 + s3);
  }
}

static size_t align_up(size_t x, size_t al) { return (x + al - 1) / al * al; }

struct Plan {
  int R, S, chunk;
  size_t m_off, g_off, ln_off, dw_off, total;
};
static Plan plan(const ocrl_sa_dims* d) {
  Plan p;
  p.R = 2 * d->T * d->K;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles = (d->N + TT - 1) / TT;
  int S = d->B > 0 ? (2 * sms) / d->B : 1;  // two CTAs per SM, one wave
  if (S < 1) S = 1;
  if (S > tiles) S = tiles;
  p.S = S;
  p.chunk = ((tiles + S - 1) / S) * TT;
  p.S = (d->N + p.chunk - 1) / p.chunk;  // ranges that hold at least one token
  p.m_off = 0;
  p.g_off = align_up(sizeof(float) * (size_t)d->B * p.R * C, 256);
  p.ln_off = align_up(p.g_off + sizeof(float) * (size_t)d->B * p.S * p.R * C, 256);
  p.dw_off = align_up(p.ln_off + sizeof(float) * (size_t)d->B * p.S * 2 * C, 256);
  p.total = align_up(p.dw_off + sizeof(float) * (size_t)DW_SPLITS * 2 * d->D * C, 256);
  return p;
}

}  // namespace lowrank

void sa_iter_bwd_ws_offsets(const ocrl_sa_dims* d, size_t* coef_off, size_t* gm_off);

size_t kv_proj_bwd_lowrank_workspace(const ocrl_sa_dims* d) { return lowrank::plan(d).total; }

int kv_proj_bwd_lowrank_launch(const ocrl_sa_dims* d, const float* x, const ocrl_token_weights* w, const float* saved,
                               const void* iter_bwd_ws, float* dx, float* d_ln_w, float* d_ln_b, float* dwk, float* dwv,
                               void* ws, cudaStream_t stream) {
  using namespace lowrank;
  if (d->C_in != C) {
    set_error("kv_proj_bwd_lowrank: C_in=%d not supported (64)", d->C_in);
    return OCRL_E_SHAPE;
  }
  const Plan p = plan(d);
  if (p.R > 224) {
    set_error("kv_proj_bwd_lowrank: 2*T*K=%d not supported (<= 224)", p.R);
    return OCRL_E_SHAPE;
  }
  size_t coef_off = 0, gm_off = 0;
  sa_iter_bwd_ws_offsets(d, &coef_off, &gm_off);
  const unsigned char* ib = reinterpret_cast<const unsigned char*>(iter_bwd_ws);
  unsigned char* base = reinterpret_cast<unsigned char*>(ws);
  Args a;
  a.x = x; a.coef = reinterpret_cast<const float*>(ib + coef_off); a.gm = reinterpret_cast<const float*>(ib + gm_off);
  a.saved = saved; a.ln_w = w->in_ln_w; a.ln_b = w->in_ln_b; a.wk = w->wk; a.wv = w->wv;
  a.dx = dx; a.M = reinterpret_cast<float*>(base + p.m_off); a.Gp = reinterpret_cast<float*>(base + p.g_off);
  a.lnp = reinterpret_cast<float*>(base + p.ln_off);
  a.dWp = reinterpret_cast<float*>(base + p.dw_off);
  a.dwk = dwk; a.dwv = dwv; a.d_ln_w = d_ln_w; a.d_ln_b = d_ln_b;
  a.B = d->B; a.N = d->N; a.D = d->D; a.H = d->H_mlp; a.K = d->K; a.T = d->T; a.R = p.R; a.S = p.S; a.chunk = p.chunk;
  a.ln_eps = d->ln_eps; a.kscale = 1.0f / sqrtf((float)d->D);

  lowrank_m_kernel<<<d->B * d->T, NT, sizeof(float) * 2 * d->K * d->D, stream>>>(a);
  ocrl::count_launch();
  OCRL_CHECK_CUDA(cudaGetLastError());
  const int RS = (p.R + 3) & ~3;
  const size_t smem = sizeof(float) * ((size_t)2 * TT * C + (size_t)TT * RS + (size_t)p.R * C + TT + 8 * 2 * C);
#define OCRL_LR_MAIN(RQ)                                                                                          \
  do {                                                                                                            \
    OCRL_CHECK_CUDA(cudaFuncSetAttribute(lowrank_main_kernel<RQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    lowrank_main_kernel<RQ><<<d->B * p.S, NT, smem, stream>>>(a);                                                 \
  } while (0)
  if (p.R <= 36) OCRL_LR_MAIN(9);
  else if (p.R <= 64) OCRL_LR_MAIN(16);
  else if (p.R <= 112) OCRL_LR_MAIN(28);
  else OCRL_LR_MAIN(56);
#undef OCRL_LR_MAIN
  ocrl::count_launch();
  OCRL_CHECK_CUDA(cudaGetLastError());
  lowrank_dw_kernel<<<dim3((d->D + DW_ROWS - 1) / DW_ROWS, 2, DW_SPLITS), NT, 0, stream>>>(a);
  ocrl::count_launch();
  OCRL_CHECK_CUDA(cudaGetLastError());
  lowrank_finish_kernel<<<(2 * d->D * C + 2 * C + NT - 1) / NT, NT, 0, stream>>>(a);
  ocrl::count_launch();
  OCRL_CHECK_CUDA(cudaGetLastError());
  return OCRL_OK;
}

}  // namespace ocrl

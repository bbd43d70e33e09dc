// Backward of norm_inputs + project_k / project_v (autograd of ocrs/common/slot_attn.py:54-61),
// fp32 FFMA.  Persistent CTAs over 128-token tiles; [s*W_k; W_v]^T stays in shared memory.
//   dxh = s * dk W_k + dv W_v            (xh = LayerNorm output)
//   dW_k = s * dk^T xh,  dW_v = dv^T xh  (accumulated in registers across the CTA's tiles)
//   dx = LayerNorm backward(dxh),  d gamma / d beta column sums
// Per-CTA partial sums go to the workspace and are reduced by a second small kernel (no atomics).
#include "common.cuh"

namespace ocrl {

constexpr int PB_C = 64;
constexpr int PB_LD = PB_C + 4;
constexpr int PB_TM = 128;
constexpr int PB_NT = 256;

struct ProjBwdArgs {
  const float* x;
  ocrl_token_weights w;
  const float* dk;
  const float* dv;
  float* dx;
  float* partial;  // [grid][2*D*64 + 128]
  int D;
  long long M;
  float ln_eps, kscale;
};

template <int D>
__global__ void __launch_bounds__(PB_NT, 1) kv_proj_bwd_kernel(const ProjBwdArgs a) {
  constexpr int NCH = 2 * D / 64;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* Wt = reinterpret_cast<float*>(smem_raw);      // [NCH][64 c][68]: Wt[ch][c][dd] = W[ch*64+dd][c]
  float* Xh = Wt + NCH * PB_C * PB_LD;                 // [128][68] LayerNorm output
  float* X0 = Xh + PB_TM * PB_LD;                      // [128][68] normalised, before the affine
  float* Gc = X0 + PB_TM * PB_LD;                      // [128][68] dk / dv chunk, then dxh
  float* rstd_s = Gc + PB_TM * PB_LD;                  // [128]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ty = tid >> 4, tx = tid & 15;

  for (int e = tid; e < 2 * D * PB_C; e += PB_NT) {
    const int row = e / PB_C, c = e % PB_C;  // row: output feature of [W_k; W_v]
    const float wv = (row < D) ? a.kscale * __ldg(a.w.wk + e) : __ldg(a.w.wv + (e - D * PB_C));
    Wt[((row / 64) * PB_C + c) * PB_LD + (row % 64)] = wv;
  }
  const float g0 = __ldg(a.w.in_ln_w + lane), g1 = __ldg(a.w.in_ln_w + lane + 32);
  const float b0 = __ldg(a.w.in_ln_b + lane), b1 = __ldg(a.w.in_ln_b + lane + 32);

  float dW[NCH][4][4];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) dW[ch][i][jj] = 0.f;
  float dgam = 0.f, dbet = 0.f;  // threads 0..63 own one column each
  __syncthreads();

  const long long ntiles = (a.M + PB_TM - 1) / PB_TM;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long m0 = tile * PB_TM;
    for (int e = tid; e < PB_TM * (PB_C / 4); e += PB_NT) {
      const int r = e / (PB_C / 4), c4 = e % (PB_C / 4);
      const long long m = m0 + r;
      float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < a.M) xv = __ldg(reinterpret_cast<const float4*>(a.x + m * PB_C) + c4);
      *reinterpret_cast<float4*>(X0 + r * PB_LD + 4 * c4) = xv;
    }
    __syncthreads();
    for (int r = warp; r < PB_TM; r += PB_NT / 32) {
      float x0 = X0[r * PB_LD + lane], x1 = X0[r * PB_LD + lane + 32];
      const float mean = warp_sum(x0 + x1) * (1.f / PB_C);
      x0 -= mean;
      x1 -= mean;
      const float rstd = rsqrtf(warp_sum(x0 * x0 + x1 * x1) * (1.f / PB_C) + a.ln_eps);
      const bool live = (m0 + r) < a.M;
      x0 = live ? x0 * rstd : 0.f;
      x1 = live ? x1 * rstd : 0.f;
      X0[r * PB_LD + lane] = x0;
      X0[r * PB_LD + lane + 32] = x1;
      Xh[r * PB_LD + lane] = live ? x0 * g0 + b0 : 0.f;
      Xh[r * PB_LD + lane + 32] = live ? x1 * g1 + b1 : 0.f;
      if (lane == 0) rstd_s[r] = rstd;
    }
    float acc[8][4] = {};
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      __syncthreads();  // previous chunk fully consumed (and, for ch == 0, the LayerNorm tile is ready)
      const float* src = (ch * 64 < D) ? a.dk + ch * 64 : a.dv + (ch * 64 - D);
      for (int e = tid; e < PB_TM * (PB_C / 4); e += PB_NT) {
        const int r = e / (PB_C / 4), c4 = e % (PB_C / 4);
        const long long m = m0 + r;
        float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < a.M) gv = __ldg(reinterpret_cast<const float4*>(src + m * D) + c4);
        *reinterpret_cast<float4*>(Gc + r * PB_LD + 4 * c4) = gv;
      }
      __syncthreads();
      // dxh[m][c] += sum_dd G[m][dd] * Wt[ch][c][dd]
      const float* Ws = Wt + ch * PB_C * PB_LD;
#pragma unroll 4
      for (int c = 0; c < PB_C; c += 4) {
        float4 av[8], wv[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) av[i] = *reinterpret_cast<const float4*>(Gc + (ty * 8 + i) * PB_LD + c);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) wv[jj] = *reinterpret_cast<const float4*>(Ws + (tx + 16 * jj) * PB_LD + c);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            acc[i][jj] = fmaf(av[i].x, wv[jj].x, acc[i][jj]);
            acc[i][jj] = fmaf(av[i].y, wv[jj].y, acc[i][jj]);
            acc[i][jj] = fmaf(av[i].z, wv[jj].z, acc[i][jj]);
            acc[i][jj] = fmaf(av[i].w, wv[jj].w, acc[i][jj]);
          }
      }
      // dW[ch*64 + 4*ty + i][4*tx + jj] += sum_m G[m][4*ty+i] * Xh[m][4*tx+jj]
#pragma unroll 4
      for (int m = 0; m < PB_TM; ++m) {
        const float4 gv = *reinterpret_cast<const float4*>(Gc + m * PB_LD + 4 * ty);
        const float4 xv = *reinterpret_cast<const float4*>(Xh + m * PB_LD + 4 * tx);
        const float gg[4] = {gv.x, gv.y, gv.z, gv.w};
        const float xx[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) dW[ch][i][jj] = fmaf(gg[i], xx[jj], dW[ch][i][jj]);
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) Gc[(ty * 8 + i) * PB_LD + tx + 16 * jj] = acc[i][jj];
    __syncthreads();
    if (tid < PB_C) {
      for (int r = 0; r < PB_TM; ++r) {
        const float gval = Gc[r * PB_LD + tid];
        dgam = fmaf(gval, X0[r * PB_LD + tid], dgam);
        dbet += gval;
      }
    }
    for (int r = warp; r < PB_TM; r += PB_NT / 32) {
      const long long m = m0 + r;
      const float ga = Gc[r * PB_LD + lane] * g0, gb = Gc[r * PB_LD + lane + 32] * g1;
      const float xa = X0[r * PB_LD + lane], xb = X0[r * PB_LD + lane + 32];
      const float m1 = warp_sum(ga + gb) * (1.f / PB_C);
      const float m2 = warp_sum(ga * xa + gb * xb) * (1.f / PB_C);
      if (m < a.M) {
        const float rs = rstd_s[r];
        a.dx[m * PB_C + lane] = rs * (ga - m1 - xa * m2);
        a.dx[m * PB_C + lane + 32] = rs * (gb - m1 - xb * m2);
      }
    }
    __syncthreads();
  }
  float* part = a.partial + (size_t)blockIdx.x * (2 * D * PB_C + 128);
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) part[(ch * 64 + 4 * ty + i) * PB_C + 4 * tx + jj] = dW[ch][i][jj];
  if (tid < PB_C) {
    part[2 * D * PB_C + tid] = dgam;
    part[2 * D * PB_C + 64 + tid] = dbet;
  }
}

__global__ void kv_proj_bwd_reduce_kernel(const float* __restrict__ partial, int ncopies, int D, float kscale,
                                          float* dwk, float* dwv, float* dgam, float* dbet) {
  const int total = 2 * D * PB_C + 128;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int c = 0; c < ncopies; ++c) s += partial[(size_t)c * total + i];
    if (i < D * PB_C) dwk[i] = s * kscale;
    else if (i < 2 * D * PB_C) dwv[i - D * PB_C] = s;
    else if (i < 2 * D * PB_C + 64) dgam[i - 2 * D * PB_C] = s;
    else dbet[i - 2 * D * PB_C - 64] = s;
  }
}

static int pb_grid(const ocrl_sa_dims* d) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long ntiles = ((long long)d->B * d->N + PB_TM - 1) / PB_TM;
  return (int)(ntiles < sms ? (ntiles < 1 ? 1 : ntiles) : sms);
}

size_t kv_proj_bwd_workspace(const ocrl_sa_dims* d) {
  return sizeof(float) * (size_t)148 * (2 * (size_t)d->D * PB_C + 128);
}

int kv_proj_bwd_launch(const ocrl_sa_dims* d, const float* x, const ocrl_token_weights* w, const float* dk,
                       const float* dv, float* dx, float* d_ln_w, float* d_ln_b, float* dwk, float* dwv, void* ws,
                       cudaStream_t stream) {
  if (d->C_in != PB_C) {
    set_error("kv_proj_bwd: C_in=%d not supported (64)", d->C_in);
    return OCRL_E_SHAPE;
  }
  ProjBwdArgs a;
  a.x = x; a.w = *w; a.dk = dk; a.dv = dv; a.dx = dx; a.partial = reinterpret_cast<float*>(ws);
  a.D = d->D; a.M = (long long)d->B * d->N; a.ln_eps = d->ln_eps; a.kscale = 1.0f / sqrtf((float)d->D);
  const int grid = pb_grid(d) > 148 ? 148 : pb_grid(d);
  const size_t smem = sizeof(float) * ((size_t)(2 * d->D / 64) * PB_C * PB_LD + 3 * PB_TM * PB_LD + PB_TM);
  switch (d->D) {
    case 64:
      OCRL_CHECK_CUDA(cudaFuncSetAttribute(kv_proj_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kv_proj_bwd_kernel<64><<<grid, PB_NT, smem, stream>>>(a);
      ocrl::count_launch();
      break;
    case 128:
      OCRL_CHECK_CUDA(cudaFuncSetAttribute(kv_proj_bwd_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kv_proj_bwd_kernel<128><<<grid, PB_NT, smem, stream>>>(a);
      ocrl::count_launch();
      break;
    case 192:
      OCRL_CHECK_CUDA(cudaFuncSetAttribute(kv_proj_bwd_kernel<192>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kv_proj_bwd_kernel<192><<<grid, PB_NT, smem, stream>>>(a);
      ocrl::count_launch();
      break;
    default:
      set_error("kv_proj_bwd: slot_size=%d not supported (64, 128, 192)", d->D);
      return OCRL_E_SHAPE;
  }
  OCRL_CHECK_CUDA(cudaGetLastError());
  kv_proj_bwd_reduce_kernel<<<32, 256, 0, stream>>>(a.partial, grid, d->D, a.kscale, dwk, dwv, d_ln_w, d_ln_b);
  ocrl::count_launch();
  OCRL_CHECK_CUDA(cudaGetLastError());
  return OCRL_OK;
}

}  // namespace ocrl

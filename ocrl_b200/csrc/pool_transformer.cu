// Fused forward of the PPO consumer's slot pooling (poolings/common/transformer.py:9-33 as configured by
// configs/pooling/transformer.yaml and used at sb3s/ocr_extractor.py:45):
//     x = Linear(slots);  x = cat([cls, x]);  y = TransformerEncoderLayer(x)  (post-norm, ReLU, one layer);  out = y[cls]
// for the rollout / evaluation path (no autograd graph, dropout inactive).  Only the CLS row of the layer's output is
// returned, so only the CLS query, its attention over the S + 1 tokens, and the CLS row of the feed-forward block are
// computed: keys / values need every token, everything after the attention one row -- 4.5 x less arithmetic than the
// module and one launch instead of ~20 (the rollout batch is 4 .. 32 images: launch latency is the cost there).
// One CTA per image, fp32 FFMA (parity 1e-4), weights stream from L2 (2.4 MB, shared by the batch).
#include "common.cuh"

namespace ocrl {
namespace pool {

constexpr int NT = 1024, NW = NT / 32;  // 32 warps: the matrix-vector products are bound by loads in flight, not by math
constexpr int MAXT = 17;   // S + 1 tokens (num_slots <= 16)
constexpr int WC = 32;     // weight columns staged per step of the multi-token products

struct Args {
  const float* slots;       // [B][S][Din]
  ocrl_pool_weights w;
  float* out;               // [B][dm]
  int B, S, Din, dm, nhead, dff;
  float ln_eps;
};

// out[r] = act(dot(W[r, 0:C], in[0:C]) + b[r]) for r < R: one warp per row, coalesced float4 loads, RB rows in flight
template <int RB>
__device__ __forceinline__ void matvec_rows(const float* __restrict__ W, const float* __restrict__ b, const float* in, float* out,
                                            int R, int C, bool relu, int warp, int lane) {
  for (int r0 = warp * RB; r0 < R; r0 += NW * RB) {
    float acc[RB];
#pragma unroll
    for (int i = 0; i < RB; ++i) acc[i] = 0.f;
    for (int c = 4 * lane; c < C; c += 128) {
      const float4 x = *reinterpret_cast<const float4*>(in + c);
      float4 wv[RB];
#pragma unroll
      for (int i = 0; i < RB; ++i)
        wv[i] = (r0 + i < R) ? __ldg(reinterpret_cast<const float4*>(W + (size_t)(r0 + i) * C + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < RB; ++i)
        acc[i] = fmaf(wv[i].x, x.x, fmaf(wv[i].y, x.y, fmaf(wv[i].z, x.z, fmaf(wv[i].w, x.w, acc[i]))));
    }
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      const float s = warp_sum(acc[i]);
      if (lane == 0 && r0 + i < R) {
        const float v = s + (b ? __ldg(b + r0 + i) : 0.f);
        out[r0 + i] = relu ? fmaxf(v, 0.f) : v;
      }
    }
  }
}

// out[t][r] = dot(W[r, 0:C], in[t][0:C]) + b[r]  for r < R (R <= NT), t < Tn: the weight matrix goes through shared memory
// WC columns at a time (coalesced loads, transposed so that thread r reads its own row without bank conflicts); the
// token values are broadcast reads.  ldin / ldout: row pitches of in / out.
template <int TMAX>
__device__ __forceinline__ void matmul_tokens(const float* __restrict__ W, const float* __restrict__ b, const float* in, int ldin,
                                              float* out, int ldout, int R, int C, int Tn, float* wst, int tid) {
  float acc[TMAX];
#pragma unroll
  for (int t = 0; t < TMAX; ++t) acc[t] = 0.f;
  const int pitch = R + 1;
  for (int c0 = 0; c0 < C; c0 += WC) {
    __syncthreads();
    for (int e = tid; e < R * WC; e += NT) {
      const int r = e / WC, c = e % WC;
      wst[c * pitch + r] = __ldg(W + (size_t)r * C + c0 + c);
    }
    __syncthreads();
    if (tid < R) {
#pragma unroll 4
      for (int c = 0; c < WC; ++c) {
        const float wv = wst[c * pitch + tid];
#pragma unroll
        for (int t = 0; t < TMAX; ++t)
          if (t < Tn) acc[t] = fmaf(wv, in[t * ldin + c0 + c], acc[t]);
      }
    }
  }
  if (tid < R) {
    const float bv = b ? __ldg(b + tid) : 0.f;
#pragma unroll
    for (int t = 0; t < TMAX; ++t)
      if (t < Tn) out[t * ldout + tid] = acc[t] + bv;
  }
  __syncthreads();
}

// LayerNorm of one row of length n (n <= 1024) by the whole CTA: out = (x - mean) * rstd * g + b
__device__ __forceinline__ void layer_norm_row(const float* x, const float* __restrict__ g, const float* __restrict__ b, float* out,
                                               int n, float eps, float* red, int tid) {
  const int warp = tid >> 5, lane = tid & 31;
  float s = 0.f;
  for (int i = tid; i < n; i += NT) s += x[i];
  s = warp_sum(s);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  float mean = 0.f;
  for (int w = 0; w < NW; ++w) mean += red[w];
  mean /= (float)n;
  __syncthreads();
  float q = 0.f;
  for (int i = tid; i < n; i += NT) {
    const float d = x[i] - mean;
    q = fmaf(d, d, q);
  }
  q = warp_sum(q);
  if (lane == 0) red[warp] = q;
  __syncthreads();
  float var = 0.f;
  for (int w = 0; w < NW; ++w) var += red[w];
  const float rstd = rsqrtf(var / (float)n + eps);
  for (int i = tid; i < n; i += NT) out[i] = (x[i] - mean) * rstd * __ldg(g + i) + __ldg(b + i);
  __syncthreads();
}

__global__ void __launch_bounds__(NT) pool_transformer_kernel(const Args a) {
  extern __shared__ __align__(16) float sm[];
  const int S = a.S, T1 = S + 1, Din = a.Din, dm = a.dm, dff = a.dff, hd = dm / a.nhead;
  float* xin = sm;                       // [S][Din]
  float* x = xin + S * Din;              // [T1][dm]  tokens (row 0 = CLS)
  float* kv = x + T1 * dm;               // [T1][2 dm] keys | values
  float* q = kv + T1 * 2 * dm;           // [dm]
  float* o = q + dm;                     // [dm] attention output of the CLS query
  float* y = o + dm;                     // [dm]
  float* x1 = y + dm;                    // [dm]
  float* h = x1 + dm;                    // [dff]
  float* prob = h + dff;                 // [nhead][T1]
  float* red = prob + a.nhead * T1;      // [NW]
  float* wst = red + NW;                 // [WC][max(dm, 2 dm) + 1] weight staging
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int img = blockIdx.x;

  const float* src = a.slots + (size_t)img * S * Din;
  for (int e = tid; e < S * Din; e += NT) xin[e] = __ldg(src + e);
  for (int e = tid; e < dm; e += NT) x[e] = __ldg(a.w.cls + e);  // poolings/common/transformer.py:26-28
  __syncthreads();
  // x[1..S] = Linear(slots)                                          poolings/common/transformer.py:23-24
  matmul_tokens<MAXT - 1>(a.w.lin_w, a.w.lin_b, xin, Din, x + dm, dm, dm, Din, S, wst, tid);
  // keys and values of every token: rows dm .. 3 dm of in_proj          nn.MultiheadAttention (in_proj_weight [3 dm, dm])
  matmul_tokens<MAXT>(a.w.in_proj_w + (size_t)dm * dm, a.w.in_proj_b + dm, x, dm, kv, 2 * dm, 2 * dm, dm, T1, wst, tid);
  // the CLS query
  matvec_rows<4>(a.w.in_proj_w, a.w.in_proj_b, x, q, dm, dm, false, warp, lane);
  __syncthreads();
  // scores and softmax per head over the T1 tokens, scaled by head_dim^-1/2
  const float scale = rsqrtf((float)hd);
  for (int e = tid; e < a.nhead * T1; e += NT) {
    const int hh = e / T1, t = e % T1;
    float s = 0.f;
    for (int d = 0; d < hd; ++d) s = fmaf(q[hh * hd + d], kv[t * 2 * dm + hh * hd + d], s);
    prob[e] = s * scale;
  }
  __syncthreads();
  if (tid < a.nhead) {
    float* p = prob + tid * T1;
    float m = p[0];
    for (int t = 1; t < T1; ++t) m = fmaxf(m, p[t]);
    float sum = 0.f;
    for (int t = 0; t < T1; ++t) { p[t] = __expf(p[t] - m); sum += p[t]; }
    const float inv = 1.f / sum;
    for (int t = 0; t < T1; ++t) p[t] *= inv;
  }
  __syncthreads();
  for (int e = tid; e < dm; e += NT) {
    const int hh = e / hd;
    float s = 0.f;
    for (int t = 0; t < T1; ++t) s = fmaf(prob[hh * T1 + t], kv[t * 2 * dm + dm + e], s);
    o[e] = s;
  }
  __syncthreads();
  // out_proj, residual, norm1 (post-norm layer: x = norm1(x + sa(x)))
  matvec_rows<4>(a.w.out_proj_w, a.w.out_proj_b, o, y, dm, dm, false, warp, lane);
  __syncthreads();
  for (int e = tid; e < dm; e += NT) y[e] += x[e];
  __syncthreads();
  layer_norm_row(y, a.w.norm1_w, a.w.norm1_b, x1, dm, a.ln_eps, red, tid);
  // feed-forward of the CLS row: x = norm2(x + linear2(relu(linear1(x))))
  matvec_rows<8>(a.w.lin1_w, a.w.lin1_b, x1, h, dff, dm, true, warp, lane);
  __syncthreads();
  matvec_rows<4>(a.w.lin2_w, a.w.lin2_b, h, y, dm, dff, false, warp, lane);
  __syncthreads();
  for (int e = tid; e < dm; e += NT) y[e] += x1[e];
  __syncthreads();
  layer_norm_row(y, a.w.norm2_w, a.w.norm2_b, o, dm, a.ln_eps, red, tid);
  for (int e = tid; e < dm; e += NT) a.out[(size_t)img * dm + e] = o[e];
}

}  // namespace pool
}  // namespace ocrl

using namespace ocrl;

extern "C" int ocrl_pool_transformer_fwd(const float* slots, const ocrl_pool_weights* w, float* out, int B, int S, int Din,
                                         int d_model, int nhead, int dff, float ln_eps, void* stream) {
  if (!slots || !w || !out || !w->lin_w || !w->in_proj_w || !w->out_proj_w || !w->lin1_w || !w->lin2_w || !w->cls) {
    set_error("pool_transformer: null pointer");
    return OCRL_E_ALIGN;
  }
  if (S < 1 || S + 1 > pool::MAXT || Din % 32 || d_model % 32 || d_model > 128 || nhead < 1 || d_model % nhead || dff % 128 ||
      dff > 4096 || Din > 512) {
    set_error("pool_transformer: S=%d Din=%d d_model=%d nhead=%d dff=%d not supported (S <= 16, Din %% 32 == 0 and <= 512, "
              "d_model %% 32 == 0 and <= 128, dff %% 128 == 0 and <= 4096)", S, Din, d_model, nhead, dff);
    return OCRL_E_SHAPE;
  }
  uintptr_t al = reinterpret_cast<uintptr_t>(slots) | reinterpret_cast<uintptr_t>(w->lin_w) | reinterpret_cast<uintptr_t>(w->in_proj_w) |
                 reinterpret_cast<uintptr_t>(w->out_proj_w) | reinterpret_cast<uintptr_t>(w->lin1_w) | reinterpret_cast<uintptr_t>(w->lin2_w);
  if (al & 15u) {
    set_error("pool_transformer: slots and weight matrices must be 16-byte aligned");
    return OCRL_E_ALIGN;
  }
  if (B <= 0) return OCRL_OK;
  pool::Args a;
  a.slots = slots; a.w = *w; a.out = out; a.B = B; a.S = S; a.Din = Din; a.dm = d_model; a.nhead = nhead; a.dff = dff;
  a.ln_eps = ln_eps;
  const int T1 = S + 1;
  const size_t floats = (size_t)S * Din + (size_t)T1 * d_model + (size_t)T1 * 2 * d_model + 4 * (size_t)d_model + dff +
                        (size_t)nhead * T1 + pool::NW + (size_t)pool::WC * (2 * d_model + 1);
  const size_t smem = floats * sizeof(float);
  OCRL_CHECK_CUDA(cudaFuncSetAttribute(pool::pool_transformer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pool::pool_transformer_kernel<<<B, pool::NT, smem, (cudaStream_t)stream>>>(a);
  ocrl::count_launch();
  OCRL_CHECK_CUDA(cudaGetLastError());
  return OCRL_OK;
}

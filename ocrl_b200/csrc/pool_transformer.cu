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


// ---------------------------------------------------------------------------------------------------------------------
// Cluster form: CL = 8 CTAs per image.  At the rollout batch (4 environments) one CTA per image leaves 144 SMs idle while
// four stream 2.4 MB of weights each through dependent L2 round trips (70 us).  Here every CTA of the cluster owns one
// attention head and an eighth of every other product -- 16 output columns of the input projection and of out_proj, 256
// rows of linear1 and the matching 256 columns of linear2 -- so a CTA streams 0.3 MB; the pieces meet through distributed
// shared memory (remote stores + cluster barrier, four exchanges).  Same arithmetic, same summation order per output
// except the feed-forward reduction, which adds the eight partial sums in rank order.
constexpr int CL = 8, CNT = 256, CNW = CNT / 32;

// out[t * ldout + r] = dot(W[r * ld + 0:C], in[t * ldin + 0:C]) + b[r]  for r < R, t < Tn (relu optional): one warp per
// row, the row's weights in registers (C <= 256), one warp reduction per (row, token)
__device__ __forceinline__ void rows_tokens(const float* __restrict__ W, int ld, const float* __restrict__ b, const float* in,
                                            int ldin, float* out, int ldout, int R, int C, int Tn, bool relu, int warp, int lane) {
  for (int r = warp; r < R; r += CNW) {
    float wv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) wv[i] = (lane + 32 * i < C) ? __ldg(W + (size_t)r * ld + lane + 32 * i) : 0.f;
    const float bv = b ? __ldg(b + r) : 0.f;
    for (int t = 0; t < Tn; ++t) {
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (lane + 32 * i < C) acc = fmaf(wv[i], in[t * ldin + lane + 32 * i], acc);
      acc = warp_sum(acc);
      if (lane == 0) {
        const float v = acc + bv;
        out[t * ldout + r] = relu ? fmaxf(v, 0.f) : v;
      }
    }
  }
}

__global__ void __launch_bounds__(CNT) pool_transformer_cluster_kernel(const Args a) {
  extern __shared__ __align__(16) float sm[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int img = blockIdx.x / CL;
  const int S = a.S, T1 = S + 1, Din = a.Din, dm = a.dm, dff = a.dff, hd = dm / a.nhead;
  const int HPC = a.nhead / CL;            // heads per CTA
  const int DSL = dm / CL, FSL = dff / CL; // this CTA's columns of d_model, rows of linear1
  float* xin = sm;                         // [S][Din]
  float* x = xin + S * Din;                // [T1][dm] all tokens (gathered)
  float* qkv = x + T1 * dm;                // [T1][3 DSL]: q | k | v of this CTA's heads (q: token 0 only)
  float* prob = qkv + T1 * 3 * DSL;        // [HPC][T1]
  float* o = prob + HPC * T1;              // [dm] attention output, gathered
  float* y = o + dm;                       // [dm] x0 + out_proj(o), gathered
  float* x1 = y + dm;                      // [dm] after norm1 (replicated)
  float* h = x1 + dm;                      // [FSL] this CTA's rows of the hidden layer
  float* part = h + FSL;                   // [CL][dm] feed-forward partial sums (used in rank 0)
  float* red = part + CL * dm;             // [CNW]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const float* src = a.slots + (size_t)img * S * Din;
  for (int e = tid; e < S * Din; e += CNT) xin[e] = __ldg(src + e);
  __syncthreads();
  // ---- x[1..S][rank's DSL columns] = Linear(slots); x[0] = cls; all-gather
  rows_tokens(a.w.lin_w + (size_t)rank * DSL * Din, Din, a.w.lin_b + rank * DSL, xin, Din, qkv, DSL, DSL, Din, S, false, warp, lane);
  __syncthreads();
  for (int e = tid; e < T1 * DSL; e += CNT) {
    const int t = e / DSL, c = e % DSL;
    const float v = (t == 0) ? __ldg(a.w.cls + rank * DSL + c) : qkv[(t - 1) * DSL + c];
    for (int dst = 0; dst < CL; ++dst) cluster.map_shared_rank(x, dst)[t * dm + rank * DSL + c] = v;
  }
  cluster.sync();
  // ---- this CTA's heads: q (CLS token), k, v (all tokens); rows rank*DSL.. of each third of in_proj
  rows_tokens(a.w.in_proj_w + (size_t)(rank * DSL) * dm, dm, a.w.in_proj_b + rank * DSL, x, dm, qkv, 3 * DSL, DSL, dm, 1, false, warp, lane);
  rows_tokens(a.w.in_proj_w + (size_t)(dm + rank * DSL) * dm, dm, a.w.in_proj_b + dm + rank * DSL, x, dm, qkv + DSL, 3 * DSL, DSL, dm,
              T1, false, warp, lane);
  rows_tokens(a.w.in_proj_w + (size_t)(2 * dm + rank * DSL) * dm, dm, a.w.in_proj_b + 2 * dm + rank * DSL, x, dm, qkv + 2 * DSL,
              3 * DSL, DSL, dm, T1, false, warp, lane);
  __syncthreads();
  const float scale = rsqrtf((float)hd);
  for (int e = tid; e < HPC * T1; e += CNT) {
    const int hh = e / T1, t = e % T1;
    float s = 0.f;
    for (int d = 0; d < hd; ++d) s = fmaf(qkv[hh * hd + d], qkv[t * 3 * DSL + DSL + hh * hd + d], s);
    prob[e] = s * scale;
  }
  __syncthreads();
  if (tid < HPC) {
    float* pr = prob + tid * T1;
    float m = pr[0];
    for (int t = 1; t < T1; ++t) m = fmaxf(m, pr[t]);
    float sum = 0.f;
    for (int t = 0; t < T1; ++t) { pr[t] = __expf(pr[t] - m); sum += pr[t]; }
    const float inv = 1.f / sum;
    for (int t = 0; t < T1; ++t) pr[t] *= inv;
  }
  __syncthreads();
  for (int e = tid; e < DSL; e += CNT) {
    const int hh = e / hd;
    float s = 0.f;
    for (int t = 0; t < T1; ++t) s = fmaf(prob[hh * T1 + t], qkv[t * 3 * DSL + 2 * DSL + e], s);
    for (int dst = 0; dst < CL; ++dst) cluster.map_shared_rank(o, dst)[rank * DSL + e] = s;
  }
  cluster.sync();
  // ---- out_proj rows of this CTA, residual, all-gather; norm1 replicated
  rows_tokens(a.w.out_proj_w + (size_t)(rank * DSL) * dm, dm, a.w.out_proj_b + rank * DSL, o, dm, qkv, DSL, DSL, dm, 1, false, warp, lane);
  __syncthreads();
  for (int e = tid; e < DSL; e += CNT) {
    const float v = qkv[e] + x[rank * DSL + e];
    for (int dst = 0; dst < CL; ++dst) cluster.map_shared_rank(y, dst)[rank * DSL + e] = v;
  }
  cluster.sync();
  {  // LayerNorm of 128 values by warp 0 (every CTA holds the full row)
    if (warp == 0) {
      float s = 0.f;
      for (int i = lane; i < dm; i += 32) s += y[i];
      const float mean = warp_sum(s) / (float)dm;
      float q2 = 0.f;
      for (int i = lane; i < dm; i += 32) { const float d = y[i] - mean; q2 = fmaf(d, d, q2); }
      const float rstd = rsqrtf(warp_sum(q2) / (float)dm + a.ln_eps);
      for (int i = lane; i < dm; i += 32) x1[i] = (y[i] - mean) * rstd * __ldg(a.w.norm1_w + i) + __ldg(a.w.norm1_b + i);
    }
    __syncthreads();
  }
  // ---- feed-forward: this CTA's FSL rows of linear1, then the matching FSL columns of linear2 (partial sums)
  rows_tokens(a.w.lin1_w + (size_t)(rank * FSL) * dm, dm, a.w.lin1_b + rank * FSL, x1, dm, h, FSL, FSL, dm, 1, true, warp, lane);
  __syncthreads();
  rows_tokens(a.w.lin2_w + rank * FSL, dff, nullptr, h, FSL, o, dm, dm, FSL, 1, false, warp, lane);  // (o is free again)
  __syncthreads();
  for (int e = tid; e < dm; e += CNT) cluster.map_shared_rank(part, 0)[rank * dm + e] = o[e];
  cluster.sync();
  if (rank == 0) {
    for (int e = tid; e < dm; e += CNT) {
      float s = __ldg(a.w.lin2_b + e);
      for (int r = 0; r < CL; ++r) s += part[r * dm + e];
      y[e] = s + x1[e];
    }
    __syncthreads();
    if (warp == 0) {
      float s = 0.f;
      for (int i = lane; i < dm; i += 32) s += y[i];
      const float mean = warp_sum(s) / (float)dm;
      float q2 = 0.f;
      for (int i = lane; i < dm; i += 32) { const float d = y[i] - mean; q2 = fmaf(d, d, q2); }
      const float rstd = rsqrtf(warp_sum(q2) / (float)dm + a.ln_eps);
      for (int i = lane; i < dm; i += 32)
        a.out[(size_t)img * dm + i] = (y[i] - mean) * rstd * __ldg(a.w.norm2_w + i) + __ldg(a.w.norm2_b + i);
    }
  }
  (void)red;
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
  // up to 64 images the cluster form (8 CTAs per image: 35 us against 70 us at the rollout batches); beyond, one CTA per
  // image keeps more images per wave (B = 256: 137 us against 160 us)
  if (B <= 64 && nhead % pool::CL == 0 && d_model % (8 * pool::CL) == 0 && dff % (32 * pool::CL) == 0 && dff / pool::CL <= 256 && Din <= 256 &&
      (d_model / nhead) * (nhead / pool::CL) == d_model / pool::CL) {
    const int DSL = d_model / pool::CL, FSL = dff / pool::CL;
    const size_t fl = (size_t)S * Din + (size_t)T1 * d_model + (size_t)T1 * 3 * DSL + (size_t)(nhead / pool::CL) * T1 + 3 * (size_t)d_model +
                      FSL + (size_t)pool::CL * d_model + pool::CNW;
    const size_t smem_c = fl * sizeof(float);
    OCRL_CHECK_CUDA(cudaFuncSetAttribute(pool::pool_transformer_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(B * pool::CL));
    cfg.blockDim = dim3(pool::CNT);
    cfg.dynamicSmemBytes = smem_c;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = pool::CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    OCRL_CHECK_CUDA(cudaLaunchKernelEx(&cfg, pool::pool_transformer_cluster_kernel, a));
    ocrl::count_launch();
    return OCRL_OK;
  }
  const size_t floats = (size_t)S * Din + (size_t)T1 * d_model + (size_t)T1 * 2 * d_model + 4 * (size_t)d_model + dff +
                        (size_t)nhead * T1 + pool::NW + (size_t)pool::WC * (2 * d_model + 1);
  const size_t smem = floats * sizeof(float);
  OCRL_CHECK_CUDA(cudaFuncSetAttribute(pool::pool_transformer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pool::pool_transformer_kernel<<<B, pool::NT, smem, (cudaStream_t)stream>>>(a);
  ocrl::count_launch();
  OCRL_CHECK_CUDA(cudaGetLastError());
  return OCRL_OK;
}

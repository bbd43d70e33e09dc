// Per-token stage in front of the slot-attention loop, fp32 FFMA version (parity mode).
//
//   [pos-table add + NCHW -> token-major]          ocrs/common/utils.py:28-33, slate_module.py:132-133
//   [LayerNorm -> Linear+ReLU -> Linear]           SlotAttentionEncoder, slot_attn.py:125-129,151
//   LayerNorm(norm_inputs) -> k = D^-1/2 W_k x^, v = W_v x^     slot_attn.py:54-61
//
// One persistent CTA per SM; the three weight matrices stay in shared memory, a 128-token
// activation tile ping-pongs between two shared buffers and never goes back to HBM between
// the layers.  Traffic per token: read C_in*4 B, write 2*D*sizeof(kv) B (+ C_in*4 B if y is kept).
#include "common.cuh"

namespace ocrl {

constexpr int TS_C = 64;        // token feature width (ocr.cnn.hidden_size)
constexpr int TS_LD = TS_C + 4; // padded row stride (floats): conflict-free LDS.128 across rows
constexpr int TS_TM = 128;      // tokens per tile
constexpr int TS_NT = 256;

struct TokenStageArgs {
  const float* x;          // [B,N,C], or NCHW [B,C,N] when nchw != 0
  const float* pos;        // [C,N] or null
  int nchw;
  ocrl_token_weights w;
  float* y_out;            // [B,N,C] or null
  void* k_out;
  void* v_out;
  int B, N, D;
  long long M;             // B*N
  float ln_eps, kscale;
};

// acc[i][jj] += sum_c A[ty*8+i][c] * W[wrow0 + tx + 16*jj][c]
__device__ __forceinline__ void tile_gemm(const float* __restrict__ As, const float* __restrict__ Ws, int wrow0,
                                          int ty, int tx, float (&acc)[8][4]) {
#pragma unroll 4
  for (int c = 0; c < TS_C; c += 4) {
    float4 av[8], wv[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) av[i] = *reinterpret_cast<const float4*>(As + (ty * 8 + i) * TS_LD + c);
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) wv[jj] = *reinterpret_cast<const float4*>(Ws + (wrow0 + tx + 16 * jj) * TS_LD + c);
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
}

// in-place LayerNorm of the 128 rows of a tile (warp per row, 2 features per lane)
__device__ __forceinline__ void tile_layer_norm(float* As, const float* __restrict__ gw, const float* __restrict__ gb,
                                                float eps, int warp, int lane) {
  const float w0 = __ldg(gw + lane), w1 = __ldg(gw + lane + 32);
  const float b0 = __ldg(gb + lane), b1 = __ldg(gb + lane + 32);
  for (int r = warp; r < TS_TM; r += TS_NT / 32) {
    float x0 = As[r * TS_LD + lane], x1 = As[r * TS_LD + lane + 32];
    const float mean = warp_sum(x0 + x1) * (1.f / TS_C);
    x0 -= mean;
    x1 -= mean;
    const float rstd = rsqrtf(warp_sum(x0 * x0 + x1 * x1) * (1.f / TS_C) + eps);
    As[r * TS_LD + lane] = x0 * rstd * w0 + b0;
    As[r * TS_LD + lane + 32] = x1 * rstd * w1 + b1;
  }
}

template <typename KV>
__global__ void __launch_bounds__(TS_NT, 1) token_stage_kernel(const TokenStageArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* As = reinterpret_cast<float*>(smem_raw);   // [128][68]
  float* Bs = As + TS_TM * TS_LD;                   // [128][68]
  float* W1s = Bs + TS_TM * TS_LD;                  // [64][68]
  float* W2s = W1s + TS_C * TS_LD;                  // [64][68]
  float* Wkv = W2s + TS_C * TS_LD;                  // [2D][68]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ty = tid >> 4, tx = tid & 15;
  const int D = a.D;
  const bool has_mlp = a.w.mlp_w1 != nullptr;

  // stage the weights once (float4 rows of 64)
  for (int e = tid; e < TS_C * (TS_C / 4); e += TS_NT) {
    const int r = e / (TS_C / 4), c4 = e % (TS_C / 4);
    if (has_mlp) {
      *reinterpret_cast<float4*>(W1s + r * TS_LD + 4 * c4) = __ldg(reinterpret_cast<const float4*>(a.w.mlp_w1) + e);
      *reinterpret_cast<float4*>(W2s + r * TS_LD + 4 * c4) = __ldg(reinterpret_cast<const float4*>(a.w.mlp_w2) + e);
    }
  }
  for (int e = tid; e < 2 * D * (TS_C / 4); e += TS_NT) {
    const int r = e / (TS_C / 4), c4 = e % (TS_C / 4);
    const float4 wv = (r < D) ? __ldg(reinterpret_cast<const float4*>(a.w.wk) + e)
                              : __ldg(reinterpret_cast<const float4*>(a.w.wv) + (e - D * (TS_C / 4)));
    *reinterpret_cast<float4*>(Wkv + r * TS_LD + 4 * c4) = wv;
  }
  __syncthreads();

  const long long ntiles = (a.M + TS_TM - 1) / TS_TM;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long m0 = tile * TS_TM;
    // ---- load the input tile ----------------------------------------------------------------
    if (!a.nchw) {
      for (int e = tid; e < TS_TM * (TS_C / 4); e += TS_NT) {
        const int r = e / (TS_C / 4), c4 = e % (TS_C / 4);
        const long long m = m0 + r;
        float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < a.M) {
          xv = __ldg(reinterpret_cast<const float4*>(a.x + m * TS_C) + c4);
          if (a.pos != nullptr) {
            const int n = (int)(m % a.N);
            xv.x += __ldg(a.pos + (long long)(4 * c4 + 0) * a.N + n);
            xv.y += __ldg(a.pos + (long long)(4 * c4 + 1) * a.N + n);
            xv.z += __ldg(a.pos + (long long)(4 * c4 + 2) * a.N + n);
            xv.w += __ldg(a.pos + (long long)(4 * c4 + 3) * a.N + n);
          }
        }
        *reinterpret_cast<float4*>(As + r * TS_LD + 4 * c4) = xv;
      }
    } else {
      // NCHW feature map + position table, transposed to token-major on the way in
      for (int e = tid; e < TS_TM * TS_C; e += TS_NT) {
        const int c = e / TS_TM, r = e % TS_TM;
        const long long m = m0 + r;
        float xv = 0.f;
        if (m < a.M) {
          const long long b = m / a.N;
          const int n = (int)(m - b * a.N);
          xv = __ldg(a.x + (b * TS_C + c) * a.N + n) + (a.pos ? __ldg(a.pos + (long long)c * a.N + n) : 0.f);
        }
        As[r * TS_LD + c] = xv;
      }
    }
    __syncthreads();
    if (has_mlp) {
      tile_layer_norm(As, a.w.enc_ln_w, a.w.enc_ln_b, a.ln_eps, warp, lane);
      __syncthreads();
      {
        float acc[8][4] = {};
        tile_gemm(As, W1s, 0, ty, tx, acc);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const float bb = __ldg(a.w.mlp_b1 + tx + 16 * jj);
#pragma unroll
          for (int i = 0; i < 8; ++i) Bs[(ty * 8 + i) * TS_LD + tx + 16 * jj] = fmaxf(acc[i][jj] + bb, 0.f);
        }
      }
      __syncthreads();
      {
        float acc[8][4] = {};
        tile_gemm(Bs, W2s, 0, ty, tx, acc);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const float bb = __ldg(a.w.mlp_b2 + tx + 16 * jj);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float yv = acc[i][jj] + bb;
            As[(ty * 8 + i) * TS_LD + tx + 16 * jj] = yv;
          }
        }
      }
      __syncthreads();
    }
    if (a.y_out != nullptr) {
      for (int e = tid; e < TS_TM * (TS_C / 4); e += TS_NT) {
        const int r = e / (TS_C / 4), c4 = e % (TS_C / 4);
        const long long m = m0 + r;
        if (m < a.M)
          *(reinterpret_cast<float4*>(a.y_out + m * TS_C) + c4) = *reinterpret_cast<const float4*>(As + r * TS_LD + 4 * c4);
      }
      __syncthreads();
    }
    tile_layer_norm(As, a.w.in_ln_w, a.w.in_ln_b, a.ln_eps, warp, lane);
    __syncthreads();
    // ---- k / v projection, 64 output features per pass ------------------------------------------
    for (int chunk = 0; chunk < 2 * D / 64; ++chunk) {
      float acc[8][4] = {};
      tile_gemm(As, Wkv, chunk * 64, ty, tx, acc);
      const bool is_k = chunk * 64 < D;
      const float sc = is_k ? a.kscale : 1.f;
      KV* outp = reinterpret_cast<KV*>(is_k ? a.k_out : a.v_out);
      const int f0 = (is_k ? chunk * 64 : chunk * 64 - D) + tx;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const long long m = m0 + ty * 8 + i;
        if (m < a.M) {
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const float val = acc[i][jj] * sc;
            if constexpr (sizeof(KV) == 4) outp[m * D + f0 + 16 * jj] = val;
            else outp[m * D + f0 + 16 * jj] = __float2bfloat16_rn(val);
          }
        }
      }
    }
    __syncthreads();
  }
}

int token_stage_launch(const ocrl_sa_dims* d, const void* x, const float* pos, const ocrl_token_weights* w,
                       float* y_out, void* k_out, void* v_out, cudaStream_t stream) {
  if (d->C_in != TS_C) {
    set_error("kv_proj: C_in=%d not supported (64)", d->C_in);
    return OCRL_E_SHAPE;
  }
  if (d->D % 64 != 0 || d->D > 192) {
    set_error("kv_proj: slot_size=%d not supported (64, 128, 192)", d->D);
    return OCRL_E_SHAPE;
  }
  if (d->x_format != OCRL_X_TOKENS_F32 && d->x_format != OCRL_X_NCHW_F32) {
    set_error("kv_proj: bf16 tokens need the tensor-core path (kv_dtype = bf16, math_mode = tensor, N %% 128 == 0)");
    return OCRL_E_SHAPE;
  }
  TokenStageArgs a;
  a.x = reinterpret_cast<const float*>(x); a.pos = pos; a.nchw = (d->x_format == OCRL_X_NCHW_F32); a.w = *w; a.y_out = y_out; a.k_out = k_out; a.v_out = v_out;
  a.B = d->B; a.N = d->N; a.D = d->D; a.M = (long long)d->B * d->N;
  a.ln_eps = d->ln_eps;
  a.kscale = 1.0f / sqrtf((float)d->D);
  const size_t smem = sizeof(float) * ((size_t)2 * TS_TM * TS_LD + 2 * TS_C * TS_LD + (size_t)2 * d->D * TS_LD);
  int dev = 0, sms = 148;
  OCRL_CHECK_CUDA(cudaGetDevice(&dev));
  OCRL_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long ntiles = (a.M + TS_TM - 1) / TS_TM;
  const int grid = (int)(ntiles < sms ? ntiles : sms);
  if (grid <= 0) return OCRL_OK;
  if (d->kv_dtype == OCRL_DT_F32) {
    OCRL_CHECK_CUDA(cudaFuncSetAttribute(token_stage_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    token_stage_kernel<float><<<grid, TS_NT, smem, stream>>>(a);
    ocrl::count_launch();
  } else {
    OCRL_CHECK_CUDA(cudaFuncSetAttribute(token_stage_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    token_stage_kernel<__nv_bfloat16><<<grid, TS_NT, smem, stream>>>(a);
    ocrl::count_launch();
  }
  OCRL_CHECK_CUDA(cudaGetLastError());
  return OCRL_OK;
}

}  // namespace ocrl

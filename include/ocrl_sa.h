/*
 * ocrl_sa.h -- C ABI of the B200-native slot-attention hot path (libocrl_sa.so).
 *
 * The reference (ugadiarov-la-phystech-edu/OCRL) has no native code; every entry point below
 * replaces a stretch of eager PyTorch in the reference and is what a binding for this path
 * would call.  Citations are relative to the reference tree.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch types.  All pointers are DEVICE pointers
 *     owned by the caller (torch allocates inputs, outputs, saved state and workspaces); the
 *     library never allocates or frees device memory and never synchronises the device.
 *   - Work is enqueued on the cudaStream_t passed as `stream` (void* to keep this header
 *     free of CUDA includes).
 *   - Return value: 0 = ok, < 0 = OCRL_E_*; ocrl_last_error() gives a thread-local message.
 *   - Tensors are dense, row-major, in the reference's layouts:
 *       x/inputs [B,N,C_in] fp32, k/v [B,N,D] (fp32 or bf16, see kv_dtype),
 *       slots [B,K,D] fp32, attn_vis [B,N,K] fp32.
 *   - Built for sm_100a only; there is no CPU or other-architecture fallback.
 */
#ifndef OCRL_SA_H_
#define OCRL_SA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OCRL_ABI_VERSION 6

enum {
  OCRL_OK = 0,
  OCRL_E_SHAPE = -1,   /* unsupported or inconsistent dimensions */
  OCRL_E_ALIGN = -2,   /* pointer not 16-byte aligned / null where required */
  OCRL_E_ARCH = -3,    /* device is not sm_100 */
  OCRL_E_LAUNCH = -4   /* CUDA launch / runtime error */
};

enum { OCRL_DT_F32 = 0, OCRL_DT_BF16 = 1 };
/* x_format: layout / type of the token-stage input */
enum { OCRL_X_TOKENS_F32 = 0  /* [B,N,C_in] fp32 */,
       OCRL_X_NCHW_F32 = 1    /* CNN feature map [B,C_in,H*W] fp32 (transposed on the way in) */,
       OCRL_X_TOKENS_BF16 = 2 /* [B,N,C_in] bf16, e.g. a channels-last bf16 feature map (tensor path only) */,
       OCRL_X_PADDED_BF16 = 3 /* bf16 feature map in the padded channels-last layout of ocrl_conv5x5_c64_tc
                                 (tensor path only; dims->frame_w = W, N = H*W) */ };
/* math_mode: how the token contractions are evaluated */
enum { OCRL_MATH_FP32 = 0 /* fp32 FFMA everywhere (parity mode) */,
       OCRL_MATH_TENSOR = 1 /* bf16 operands on tensor cores, fp32 accumulate */ };

/* Problem description, mirrors SlotAttention.__init__ (ocrs/common/slot_attn.py:10-45) and the
 * Hydra keys ocr.slotattr.{num_iterations,num_slots,slot_size,mlp_hidden_size,num_slot_heads}
 * (configs/ocr/slate.yaml:14-20). */
typedef struct ocrl_sa_dims {
  int32_t B;        /* images in the batch */
  int32_t N;        /* tokens per image (H*W) */
  int32_t C_in;     /* token feature width (ocr.cnn.hidden_size, 64) */
  int32_t D;        /* slot_size */
  int32_t H_mlp;    /* mlp_hidden_size */
  int32_t K;        /* num_slots, 1..16 */
  int32_t T;        /* num_iterations */
  int32_t heads;    /* num_slot_heads; only 1 is supported (all shipped configs) */
  float eps;        /* SlotAttention epsilon, 1e-8 (slot_attn.py:18,86) */
  float ln_eps;     /* nn.LayerNorm eps, 1e-5 */
  int32_t kv_dtype; /* OCRL_DT_* storage type of k and v */
  int32_t math_mode;/* OCRL_MATH_* */
  int32_t x_format; /* OCRL_X_* (token stage only) */
  int32_t frame_w;  /* frame width W for OCRL_X_PADDED_BF16 (0 otherwise) */
} ocrl_sa_dims;

/* Parameters of the iteration loop, fp32, in the reference's state_dict layout
 * (slot_attention.{norm_slots,norm_mlp,project_q,gru,mlp}; slot_attn.py:31-45). */
typedef struct ocrl_sa_weights {
  const float* ln_slots_w; const float* ln_slots_b;   /* norm_slots   [D]       */
  const float* ln_mlp_w;   const float* ln_mlp_b;     /* norm_mlp     [D]       */
  const float* wq;                                    /* project_q.weight [D,D] */
  const float* w_ih; const float* w_hh;               /* gru.weight_* [3D,D], rows [r;z;n] */
  const float* b_ih; const float* b_hh;               /* gru.bias_*   [3D]      */
  const float* w1; const float* b1;                   /* mlp.0 [H,D],[H]        */
  const float* w2; const float* b2;                   /* mlp.2 [D,H],[D]        */
} ocrl_sa_weights;

/* Gradient accumulators with the same shapes (written, not accumulated into). */
typedef struct ocrl_sa_weight_grads {
  float* ln_slots_w; float* ln_slots_b;
  float* ln_mlp_w;   float* ln_mlp_b;
  float* wq;
  float* w_ih; float* w_hh;
  float* b_ih; float* b_hh;
  float* w1; float* b1;
  float* w2; float* b2;
} ocrl_sa_weight_grads;

/* Parameters of the per-token stage in front of the loop:
 *   SlotAttentionEncoder.layer_norm + mlp (slot_attn.py:125-129,151) when mlp_w1 != NULL,
 *   then SlotAttention.norm_inputs + project_k / project_v (slot_attn.py:30,36-37,54-61). */
typedef struct ocrl_token_weights {
  const float* enc_ln_w; const float* enc_ln_b;       /* layer_norm [C_in] (NULL: skip the MLP) */
  const float* mlp_w1; const float* mlp_b1;           /* mlp.0 [C_in,C_in],[C_in] */
  const float* mlp_w2; const float* mlp_b2;           /* mlp.2 [C_in,C_in],[C_in] */
  const float* in_ln_w; const float* in_ln_b;         /* norm_inputs [C_in] */
  const float* wk; const float* wv;                   /* project_{k,v}.weight [D,C_in] */
} ocrl_token_weights;

int ocrl_version(void);
/* "sm_100a": the only architecture the library contains code for. */
const char* ocrl_built_arch(void);
const char* ocrl_last_error(void);
/* Number of kernels this library has launched (or captured into a CUDA graph) in this process so far. */
unsigned long long ocrl_launch_count(void);

/* Bytes the caller must provide: `fwd_ws`/`bwd_ws` scratch for the iteration kernels and
 * `saved` for the per-iteration state kept for the backward (slots_in, updates, row sums). */
int ocrl_sa_query_workspace(const ocrl_sa_dims* dims, size_t* fwd_ws, size_t* bwd_ws, size_t* saved);

/* Token stage, forward.  Replaces slot_attn.py:151 (optional) and :54-61.
 *   x        the tokens in dims->x_format: [B,N,C_in] fp32 / bf16, or the CNN feature map
 *            [B,C_in,H*W] (NCHW), transposed to token-major on the way in (slate_module.py:133).
 *   pos_table NULL, or the position-embedding table [C_in,H*W] fp32 that is added to every image's
 *            tokens (ocrs/common/utils.py:28-33).
 *   y_out    [B,N,C_in] fp32 output of the token MLP (NULL if not needed / MLP skipped)
 *   k_out,v_out [B,N,D] in dims->kv_dtype; k already scaled by D^-1/2 (slot_attn.py:61).
 *   workspace: ocrl_kv_proj_fwd_workspace(dims) bytes of scratch (bf16 weight copies for the tensor-core
 *            path); may be NULL, which selects the fp32 FFMA kernel. */
size_t ocrl_kv_proj_fwd_workspace(const ocrl_sa_dims* dims);
int ocrl_kv_proj_fwd(const ocrl_sa_dims* dims, const void* x, const float* pos_table,
                     const ocrl_token_weights* w, float* y_out, void* k_out, void* v_out,
                     void* workspace, void* stream);

/* Token stage, backward of norm_inputs + project_k/v (autograd of slot_attn.py:54-61).
 *   x [B,N,C_in] is the input of norm_inputs; dk,dv [B,N,D] fp32.
 *   Writes dx [B,N,C_in], d in_ln_w/b [C_in], dwk, dwv [D,C_in].
 *   ws: scratch of ocrl_kv_proj_bwd_workspace(dims) bytes. */
size_t ocrl_kv_proj_bwd_workspace(const ocrl_sa_dims* dims);
int ocrl_kv_proj_bwd(const ocrl_sa_dims* dims, const float* x, const ocrl_token_weights* w,
                     const float* dk, const float* dv, float* dx, float* d_ln_w, float* d_ln_b,
                     float* dwk, float* dwv, void* ws, void* stream);

/* Token stage, backward, straight from the rank-(2 T K) coefficients of ocrl_sa_iter_bwd: call ocrl_sa_iter_bwd with
 * dk = dv = NULL and pass its workspace here (plus the forward's `saved`); dk, dv [B,N,D] are never materialised
 * (dxh_n = sum_r coef_nr M_r with M = [s W_k^T q ; W_v^T dU/S], dW from coef^T xh).  Same outputs as ocrl_kv_proj_bwd.
 *   ws: scratch of ocrl_kv_proj_bwd_lowrank_workspace(dims) bytes.  C_in = 64, 2 T K <= 224. */
size_t ocrl_kv_proj_bwd_lowrank_workspace(const ocrl_sa_dims* dims);
int ocrl_kv_proj_bwd_lowrank(const ocrl_sa_dims* dims, const float* x, const ocrl_token_weights* w, const void* saved,
                             const void* iter_bwd_workspace, float* dx, float* d_ln_w, float* d_ln_b,
                             float* dwk, float* dwv, void* ws, void* stream);

/* The fused T-iteration loop, forward.  Replaces slot_attn.py:64-102.
 *   k,v [B,N,D]; slots0 [B,K,D]; slots_out [B,K,D]; attn_vis_out [B,N,K] or NULL;
 *   saved: NULL (inference) or `saved` bytes from ocrl_sa_query_workspace;
 *   workspace: `fwd_ws` bytes (may be NULL when fwd_ws == 0). */
int ocrl_sa_iter_fwd(const ocrl_sa_dims* dims, const void* k, const void* v, const float* slots0,
                     const ocrl_sa_weights* w, float* slots_out, float* attn_vis_out,
                     void* saved, void* workspace, void* stream);

/* Launch options of the forward loop (everything that used to be a process-global setting).  A zeroed struct, or a
 * NULL pointer, is the default: the fastest kernel that covers the shape, its own grid size.
 *   variant       which implementation to run (OCRL_SA_AUTO picks in the order TCGEN05, PIPE, CLUSTER_TC, FFMA for bf16 k/v
 *                 with math_mode TENSOR; fp32 k/v always run FFMA)
 *   max_clusters  > 0: upper bound on the resident clusters of the persistent kernels (TCGEN05, PIPE) -- a latency-bound
 *                 launch on fewer, fuller clusters leaves SMs to concurrent streams; results do not depend on it
 *   lanes         images in flight per cluster of the persistent kernels: 0 = default, 2 or 3
 *   strict        != 0: return OCRL_E_SHAPE when `variant` (or, with OCRL_SA_AUTO, the tcgen05 kernel) does not cover the
 *                 shape instead of running a slower kernel
 *   trace         != 0: the kernel writes clock64() phase stamps of CTA 0 to the last 4 KB of `workspace` (development aid)
 *   prepared      != 0: `workspace` still holds the prepared (bf16, LayerNorm-folded) weight copies an earlier call with
 *                 the same dims, weights and workspace left there; the weight-preparation launch is skipped.  The caller
 *                 owns the invariant (inference with frozen weights: prepare once, then pass prepared = 1)
 */
enum { OCRL_SA_AUTO = 0, OCRL_SA_TCGEN05 = 1, OCRL_SA_PIPE = 2, OCRL_SA_CLUSTER_TC = 3, OCRL_SA_FFMA = 4 };
typedef struct ocrl_sa_launch_opts {
  int32_t variant;
  int32_t max_clusters;
  int32_t lanes;
  int32_t strict;
  int32_t trace;
  int32_t prepared;
} ocrl_sa_launch_opts;
int ocrl_sa_iter_fwd_ex(const ocrl_sa_dims* dims, const void* k, const void* v, const float* slots0,
                        const ocrl_sa_weights* w, float* slots_out, float* attn_vis_out,
                        void* saved, void* workspace, const ocrl_sa_launch_opts* opts, void* stream);
/* Name of the kernel the last ocrl_sa_iter_fwd / _ex call of this thread launched ("" before the first call):
 * "tcgen05", "tcgen05_xhat" (factored form), "pipe", "cluster_tc" or "ffma". */
const char* ocrl_sa_last_kernel(void);

/* FACTORED form of the inference path (no `saved`, no backward).  project_k / project_v have no bias
 * (slot_attn.py:36-37), so k = s W_k x^ and v = W_v x^ are rank-C_in functions of the normalised tokens
 * x^ = norm_inputs(x) [B,N,C_in]:   k_n . q = x^_n . (s W_k^T q)   and   sum_n w_n v_n = W_v (sum_n w_n x^_n).
 * The loop can therefore stream x^ (2 C_in bytes per token in bf16) instead of k and v (4 D bytes per token) and fold
 * the two projections into the slot-update weights (W_q'' = s W_k^T W_q, W_ih'' = W_ih W_v; prepared once per
 * parameter version like the other bf16 weight copies).  Same results as ocrl_kv_proj_fwd + ocrl_sa_iter_fwd within
 * the bf16-mode tolerance (2e-2; measured 3e-4 on the goldens, the k/v form measures 2.4e-4).
 *   ocrl_xhat_fwd: the token stage up to and including norm_inputs; xhat_out [B,N,C_in] bf16.  Arguments as
 *     ocrl_kv_proj_fwd (w->wk / w->wv are not read); workspace of ocrl_kv_proj_fwd_workspace(dims) bytes, required.
 *   ocrl_sa_iter_fwd_xhat: slot_attn.py:64-102 on x^; wk, wv = project_{k,v}.weight [D,C_in] fp32; workspace = `fwd_ws`
 *     of ocrl_sa_query_workspace; opts as in ocrl_sa_iter_fwd_ex (variant must be AUTO or TCGEN05).
 * Covered: C_in = 64, (D, H_mlp) = (192, 192), K <= 16, math_mode TENSOR; anything else returns OCRL_E_SHAPE (the
 * caller then takes the k/v form). */
int ocrl_xhat_fwd(const ocrl_sa_dims* dims, const void* x, const float* pos_table, const ocrl_token_weights* w,
                  float* y_out, void* xhat_out, void* workspace, void* stream);
int ocrl_sa_iter_fwd_xhat(const ocrl_sa_dims* dims, const void* xhat, const float* wk, const float* wv,
                          const float* slots0, const ocrl_sa_weights* w, float* slots_out, float* attn_vis_out,
                          void* workspace, const ocrl_sa_launch_opts* opts, void* stream);

/* The fused backward of the loop; attention logits are recomputed from k and the saved
 * per-iteration slots rather than stored (autograd of slot_attn.py:64-102).
 *   d_slots [B,K,D]; d_attn_vis [B,N,K] or NULL;
 *   writes dk, dv [B,N,D] fp32 (both may be NULL: the per-token coefficients stay in `workspace` for
 *   ocrl_kv_proj_bwd_lowrank), d_slots0 [B,K,D] and every member of dw. */
int ocrl_sa_iter_bwd(const ocrl_sa_dims* dims, const void* k, const void* v, const void* saved,
                     const ocrl_sa_weights* w, const float* d_slots, const float* d_attn_vis,
                     float* dk, float* dv, float* d_slots0, const ocrl_sa_weight_grads* dw,
                     void* workspace, void* stream);

/* Element-wise pieces of the feature stage between the (library) convolutions, bf16 inference path.
 *   ocrl_conv_bias_relu_bf16: y = relu(y + bias[c]) in place on a channels-last bf16 tensor [npixels, channels]
 *     -- the bias + ReLU of Conv2dBlock (ocrs/common/networks.py:38-53); bias fp32 [channels].
 *   ocrl_frames_to_nhwc_bf16: obs [B,C,H,W] fp32 in [0,1] (utils/datasets.py:17) -> out [B,H,W,CP] bf16 with the
 *     channels zero-padded to CP = 8 (input of the first convolution, ocrs/common/models.py:99). */
int ocrl_conv_bias_relu_bf16(void* y, const float* bias, long long npixels, int channels, void* stream);
/* First layer of SlotAttnCNNEncoder (ocrs/common/models.py:99 = Conv2dBlock(obs_channels, 64, 5, 1, 2)), fused:
 *   out[B,H,W,64] (bf16, channels-last) = relu(conv5x5(obs[B,C,H,W] fp32, weight[64,C,5,5]) + bias[64]),
 * operands rounded to bf16, fp32 accumulate (the arithmetic of the module under bf16 autocast).  C = 3, CO = 64,
 * W a multiple of 16. */
int ocrl_conv_first_relu_bf16(const float* obs, const float* weight, const float* bias, void* out, int B, int C,
                              int H, int W, int CO, void* stream);
int ocrl_frames_to_nhwc_bf16(const float* obs, void* out, int B, int C, int H, int W, int CP, void* stream);

/* Layers 2-4 of SlotAttnCNNEncoder (ocrs/common/models.py:100-102: Conv2dBlock(64, 64, 5, 1, 2) x2 and the final
 * conv2d(64, 64, 5, 1, 2)) as a hand-written tcgen05 implicit GEMM, bf16 operands / fp32 accumulate.
 * Activations use a PADDED channels-last layout: a flat bf16 array [(2 + B*(H+2)) rows][W+4][64]; image b, pixel (y, x)
 * sits at row 2 + b*(H+2) + y, column x + 2; every other position is zero (the convolution's zero padding is stored).
 * The caller allocates ocrl_conv_padded_bytes(B, H, W) bytes per array; no initialisation is needed: both producers
 * below write every position of their output (real pixels, or zeros at the padding positions).
 *   ocrl_conv5x5_pack_weights: weight [64,64,5,5] fp32 (OIHW, the reference's state_dict) -> packed bf16 [25][64][64]
 *     (25*64*64*2 bytes, 128-byte aligned); once per parameter version.
 *   ocrl_conv5x5_c64_tc: out = conv5x5(in) (+ bias[64] if not NULL) (ReLU if relu != 0); W in {32, 64, 128}.
 *   ocrl_conv_first_relu_bf16p: ocrl_conv_first_relu_bf16 writing the padded layout. */
size_t ocrl_conv_padded_bytes(int B, int H, int W);
int ocrl_conv5x5_pack_weights(const float* weight, void* packed, int CO, int CI, void* stream);
int ocrl_conv5x5_c64_tc(const void* in_padded, const void* packed_w, const float* bias, void* out_padded, int B, int H,
                        int W, int relu, void* stream);
int ocrl_conv_first_relu_bf16p(const float* obs, const float* weight, const float* bias, void* out_padded, int B, int C,
                               int H, int W, int CO, void* stream);
/* The same layer fed with the frames as the datasets / environments hold them: uint8 HWC [B,H,W,C]
 * (utils/datasets.py:17).  The ingest `obs / 255.0` (fp32) happens while the rows are staged: bit-identical to
 * ocrl_conv_first_relu_bf16p on the converted float CHW tensor, a quarter of the input bytes. */
int ocrl_conv_first_relu_u8p(const unsigned char* frames_hwc, const float* weight, const float* bias, void* out_padded,
                             int B, int C, int H, int W, int CO, void* stream);

/* PPO consumer: the slot pooling of poolings/common/transformer.py:9-33 (configs/pooling/transformer.yaml: d_model 128,
 * 8 heads, ONE post-norm nn.TransformerEncoderLayer with ReLU and dim_feedforward 2048, no positional encoding), forward
 * only, dropout inactive -- what sb3s/ocr_extractor.py:45 runs on the slots during rollouts:
 *   out[B, d_model] = TransformerEncoderLayer(cat([cls, Linear(slots)]))[cls]
 * slots [B,S,Din] fp32 (S <= 16, Din % 32 == 0); weights are the module's own parameters (PyTorch layouts). */
typedef struct ocrl_pool_weights {
  const float* lin_w; const float* lin_b;             /* _linear [d_model, Din], [d_model] */
  const float* cls;                                    /* _cls_token._cls_token [d_model] */
  const float* in_proj_w; const float* in_proj_b;     /* self_attn.in_proj_weight [3 d_model, d_model], bias [3 d_model] */
  const float* out_proj_w; const float* out_proj_b;   /* self_attn.out_proj [d_model, d_model], [d_model] */
  const float* lin1_w; const float* lin1_b;           /* linear1 [dff, d_model], [dff] */
  const float* lin2_w; const float* lin2_b;           /* linear2 [d_model, dff], [d_model] */
  const float* norm1_w; const float* norm1_b;         /* norm1 [d_model] */
  const float* norm2_w; const float* norm2_b;         /* norm2 [d_model] */
} ocrl_pool_weights;
int ocrl_pool_transformer_fwd(const float* slots, const ocrl_pool_weights* w, float* out, int B, int S, int Din,
                              int d_model, int nhead, int dff, float ln_eps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OCRL_SA_H_ */

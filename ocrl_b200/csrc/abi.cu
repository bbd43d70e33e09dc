// extern "C" surface of libocrl_sa.so (see include/ocrl_sa.h).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "slot_math.cuh"

namespace ocrl {

static thread_local char g_err[512] = "";
static unsigned long long g_launches = 0;  // kernels this library launched (or captured into a CUDA graph) so far
void count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int token_stage_launch(const ocrl_sa_dims* d, const void* x, const float* pos, const ocrl_token_weights* w,
                       float* y_out, void* k_out, void* v_out, cudaStream_t stream);
int sa_iter_fwd_launch(const ocrl_sa_dims* d, const void* k, const void* v, const float* slots0,
                       const ocrl_sa_weights* w, float* slots_out, float* attn_out, float* saved,
                       void* workspace, const ocrl_sa_launch_opts* opts, cudaStream_t stream);
const char* sa_iter_last_kernel();
int sa_iter_bwd_launch(const ocrl_sa_dims* d, const void* k, const void* v, const float* saved,
                       const ocrl_sa_weights* w, const float* d_slots, const float* d_attn, float* dk, float* dv,
                       float* d_slots0, const ocrl_sa_weight_grads* dw, void* ws, cudaStream_t stream);
size_t sa_iter_bwd_workspace(const ocrl_sa_dims* d);
int kv_proj_bwd_launch(const ocrl_sa_dims* d, const float* x, const ocrl_token_weights* w, const float* dk,
                       const float* dv, float* dx, float* d_ln_w, float* d_ln_b, float* dwk, float* dwv, void* ws,
                       cudaStream_t stream);
size_t kv_proj_bwd_workspace(const ocrl_sa_dims* d);
int kv_proj_bwd_lowrank_launch(const ocrl_sa_dims* d, const float* x, const ocrl_token_weights* w, const float* saved,
                               const void* iter_bwd_ws, float* dx, float* d_ln_w, float* d_ln_b, float* dwk, float* dwv,
                               void* ws, cudaStream_t stream);
size_t kv_proj_bwd_lowrank_workspace(const ocrl_sa_dims* d);
size_t kv_proj_tc_workspace(const ocrl_sa_dims* d);
size_t sa_iter_tc_workspace(const ocrl_sa_dims* d);
int kv_proj_tc_launch(const ocrl_sa_dims* d, const void* x, const float* pos, const ocrl_token_weights* w, float* y_out,
                      void* k_out, void* v_out, void* workspace, cudaStream_t stream, void* xhat_out = nullptr);
int sa_iter_fwd_xhat_launch(const ocrl_sa_dims* d, const void* xhat, const float* wk, const float* wv, const float* slots0,
                            const ocrl_sa_weights* w, float* slots_out, float* attn_out, void* workspace,
                            const ocrl_sa_launch_opts* opts, cudaStream_t stream);

// Cluster size = CTAs per image.  Needs D % CL == 0 and H % CL == 0 (each CTA owns D/CL slot
// features in the GRU/MLP) and enough tokens per CTA to keep 8 warps busy.
int sa_iter_pick_cluster(const ocrl_sa_dims* d) {
  const int cands[4] = {8, 4, 2, 1};
  for (int i = 0; i < 4; ++i) {
    const int cl = cands[i];
    if (d->D % cl || d->H_mlp % cl) continue;
    // measured on B200 (profiles/r1/sweep_cl.log): ~2048-4096 tokens per CTA is the sweet spot -- larger
    // clusters spend more time in the DSMEM exchanges of the slot update than they save in the token pass
    const int want = d->N >= 8192 ? 4 : (d->N >= 512 ? 2 : 1);
    if (cl > want) continue;
    return cl;
  }
  return 1;
}

// Warps per CTA of the iteration kernel: 4-warp CTAs let two clusters share every SM (the slot
// update of one image overlaps the token pass of another); needs <= ~112 KB of shared memory per CTA.
int sa_iter_pick_warps(int D, int KP, int CL) {
  (void)D; (void)KP; (void)CL;
  return 8;
}

static int check_dims(const ocrl_sa_dims* d) {
  if (!d) { set_error("dims is null"); return OCRL_E_SHAPE; }
  if (d->heads != 1) { set_error("num_slot_heads=%d not supported (1)", d->heads); return OCRL_E_SHAPE; }
  if (d->B < 0 || d->N <= 0 || d->K < 1 || d->K > 16 || d->T < 1) {
    set_error("bad dims B=%d N=%d K=%d T=%d", d->B, d->N, d->K, d->T);
    return OCRL_E_SHAPE;
  }
  if (d->D % 64 || d->H_mlp % 64 || d->D > 192 || d->H_mlp > 512) {
    set_error("slot_size=%d / mlp_hidden_size=%d not supported (multiples of 64, D <= 192)", d->D, d->H_mlp);
    return OCRL_E_SHAPE;
  }
  if (d->kv_dtype != OCRL_DT_F32 && d->kv_dtype != OCRL_DT_BF16) {
    set_error("bad kv_dtype %d", d->kv_dtype);
    return OCRL_E_SHAPE;
  }
  return OCRL_OK;
}

static int check_arch() {
  static int cached = 0;  // 0 unknown, 1 ok, -1 bad
  if (cached == 0) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
      set_error("no CUDA device");
      return OCRL_E_ARCH;
    }
    cached = (major == 10) ? 1 : -1;
  }
  if (cached < 0) {
    set_error("libocrl_sa is built for sm_100a only");
    return OCRL_E_ARCH;
  }
  return OCRL_OK;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace ocrl

using namespace ocrl;

extern "C" {

int ocrl_version(void) { return OCRL_ABI_VERSION; }
const char* ocrl_built_arch(void) { return "sm_100a"; }
const char* ocrl_last_error(void) { return g_err; }
unsigned long long ocrl_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

int ocrl_sa_query_workspace(const ocrl_sa_dims* d, size_t* fwd_ws, size_t* bwd_ws, size_t* saved) {
  int rc = check_dims(d);
  if (rc) return rc;
  if (fwd_ws) *fwd_ws = sa_iter_tc_workspace(d);
  if (bwd_ws) *bwd_ws = sa_iter_bwd_workspace(d);
  if (saved) *saved = sizeof(float) * (size_t)d->B * d->T * (size_t)SavedLayout(d->K, d->D, d->H_mlp).stride();
  return OCRL_OK;
}

size_t ocrl_kv_proj_fwd_workspace(const ocrl_sa_dims* d) {
  if (check_dims(d)) return 0;
  return kv_proj_tc_workspace(d);
}

int ocrl_kv_proj_fwd(const ocrl_sa_dims* d, const void* x, const float* pos_table, const ocrl_token_weights* w,
                     float* y_out, void* k_out, void* v_out, void* workspace, void* stream) {
  int rc = check_dims(d);
  if (rc) return rc;
  if ((rc = check_arch())) return rc;
  if (!x || !w || !k_out || !v_out || !w->in_ln_w || !w->in_ln_b || !w->wk || !w->wv) {
    set_error("kv_proj_fwd: null pointer");
    return OCRL_E_ALIGN;
  }
  if (!aligned16(x) || !aligned16(k_out) || !aligned16(v_out) || !aligned16(w->wk) || !aligned16(w->wv) ||
      (y_out && !aligned16(y_out)) || (w->mlp_w1 && (!aligned16(w->mlp_w1) || !aligned16(w->mlp_w2)))) {
    set_error("kv_proj_fwd: pointers must be 16-byte aligned");
    return OCRL_E_ALIGN;
  }
  if (d->B == 0) return OCRL_OK;
  if (d->kv_dtype == OCRL_DT_BF16 && d->math_mode == OCRL_MATH_TENSOR && workspace != nullptr) {
    rc = kv_proj_tc_launch(d, x, pos_table, w, y_out, k_out, v_out, workspace, (cudaStream_t)stream);
    if (rc != OCRL_E_SHAPE) return rc;  // shapes the tcgen05 kernel does not cover take the FFMA kernel
  }
  if (d->x_format == OCRL_X_PADDED_BF16 || d->x_format == OCRL_X_TOKENS_BF16) {
    set_error("kv_proj_fwd: bf16 token / padded feature-map input is a tensor-path format (bf16 k/v, math_mode TENSOR, workspace)");
    return OCRL_E_SHAPE;
  }
  return token_stage_launch(d, x, pos_table, w, y_out, k_out, v_out, (cudaStream_t)stream);
}

int ocrl_xhat_fwd(const ocrl_sa_dims* d, const void* x, const float* pos_table, const ocrl_token_weights* w, float* y_out,
                  void* xhat_out, void* workspace, void* stream) {
  int rc = check_dims(d);
  if (rc) return rc;
  if ((rc = check_arch())) return rc;
  if (!x || !w || !xhat_out || !workspace || !w->in_ln_w || !w->in_ln_b) {
    set_error("xhat_fwd: null pointer (x, weights, xhat_out and the workspace are required)");
    return OCRL_E_ALIGN;
  }
  if (!aligned16(x) || !aligned16(xhat_out) || (y_out && !aligned16(y_out)) ||
      (w->mlp_w1 && (!aligned16(w->mlp_w1) || !aligned16(w->mlp_w2)))) {
    set_error("xhat_fwd: pointers must be 16-byte aligned");
    return OCRL_E_ALIGN;
  }
  if (d->math_mode != OCRL_MATH_TENSOR) {
    set_error("xhat_fwd: the factored form is a tensor-path feature (math_mode TENSOR)");
    return OCRL_E_SHAPE;
  }
  if (d->B == 0) return OCRL_OK;
  rc = kv_proj_tc_launch(d, x, pos_table, w, y_out, nullptr, nullptr, workspace, (cudaStream_t)stream, xhat_out);
  if (rc == OCRL_E_SHAPE) set_error("xhat_fwd: shape not covered (C_in = 64, D in {64, 128, 192})");
  return rc;
}

int ocrl_sa_iter_fwd_xhat(const ocrl_sa_dims* d, const void* xhat, const float* wk, const float* wv, const float* slots0,
                          const ocrl_sa_weights* w, float* slots_out, float* attn_vis_out, void* workspace,
                          const ocrl_sa_launch_opts* opts, void* stream) {
  int rc = check_dims(d);
  if (rc) return rc;
  if ((rc = check_arch())) return rc;
  if (!xhat || !wk || !wv || !slots0 || !w || !slots_out) {
    set_error("sa_iter_fwd_xhat: null pointer");
    return OCRL_E_ALIGN;
  }
  if (!aligned16(xhat) || !aligned16(wk) || !aligned16(wv) || !aligned16(w->wq) || !aligned16(w->w_ih) ||
      !aligned16(w->w_hh) || !aligned16(w->w1) || !aligned16(w->w2)) {
    set_error("sa_iter_fwd_xhat: x^ and weight matrices must be 16-byte aligned");
    return OCRL_E_ALIGN;
  }
  if (d->B == 0) return OCRL_OK;
  if (opts && (opts->max_clusters < 0 || (opts->lanes != 0 && opts->lanes != 2 && opts->lanes != 3))) {
    set_error("sa_iter_fwd_xhat: bad launch options (max_clusters %d, lanes %d)", opts->max_clusters, opts->lanes);
    return OCRL_E_SHAPE;
  }
  return sa_iter_fwd_xhat_launch(d, xhat, wk, wv, slots0, w, slots_out, attn_vis_out, workspace, opts, (cudaStream_t)stream);
}

size_t ocrl_kv_proj_bwd_workspace(const ocrl_sa_dims* d) {
  if (check_dims(d)) return 0;
  return kv_proj_bwd_workspace(d);
}

int ocrl_kv_proj_bwd(const ocrl_sa_dims* d, const float* x, const ocrl_token_weights* w, const float* dk,
                     const float* dv, float* dx, float* d_ln_w, float* d_ln_b, float* dwk, float* dwv, void* ws,
                     void* stream) {
  int rc = check_dims(d);
  if (rc) return rc;
  if ((rc = check_arch())) return rc;
  if (!x || !w || !dk || !dv || !dx || !d_ln_w || !d_ln_b || !dwk || !dwv) {
    set_error("kv_proj_bwd: null pointer");
    return OCRL_E_ALIGN;
  }
  if (!aligned16(x) || !aligned16(dk) || !aligned16(dv) || !aligned16(dx) || !aligned16(ws)) {
    set_error("kv_proj_bwd: pointers must be 16-byte aligned");
    return OCRL_E_ALIGN;
  }
  return kv_proj_bwd_launch(d, x, w, dk, dv, dx, d_ln_w, d_ln_b, dwk, dwv, ws, (cudaStream_t)stream);
}

size_t ocrl_kv_proj_bwd_lowrank_workspace(const ocrl_sa_dims* d) {
  if (check_dims(d)) return 0;
  return kv_proj_bwd_lowrank_workspace(d);
}

int ocrl_kv_proj_bwd_lowrank(const ocrl_sa_dims* d, const float* x, const ocrl_token_weights* w, const void* saved,
                             const void* iter_bwd_workspace, float* dx, float* d_ln_w, float* d_ln_b, float* dwk,
                             float* dwv, void* ws, void* stream) {
  int rc = check_dims(d);
  if (rc) return rc;
  if ((rc = check_arch())) return rc;
  if (!x || !w || !saved || !iter_bwd_workspace || !dx || !d_ln_w || !d_ln_b || !dwk || !dwv || !ws) {
    set_error("kv_proj_bwd_lowrank: null pointer");
    return OCRL_E_ALIGN;
  }
  if (!aligned16(x) || !aligned16(dx) || !aligned16(ws) || !aligned16(iter_bwd_workspace)) {
    set_error("kv_proj_bwd_lowrank: pointers must be 16-byte aligned");
    return OCRL_E_ALIGN;
  }
  if (d->B == 0) return OCRL_OK;
  return kv_proj_bwd_lowrank_launch(d, x, w, reinterpret_cast<const float*>(saved), iter_bwd_workspace, dx, d_ln_w,
                                    d_ln_b, dwk, dwv, ws, (cudaStream_t)stream);
}

const char* ocrl_sa_last_kernel(void) { return sa_iter_last_kernel(); }

int ocrl_sa_iter_fwd(const ocrl_sa_dims* d, const void* k, const void* v, const float* slots0,
                     const ocrl_sa_weights* w, float* slots_out, float* attn_vis_out, void* saved, void* workspace,
                     void* stream) {
  return ocrl_sa_iter_fwd_ex(d, k, v, slots0, w, slots_out, attn_vis_out, saved, workspace, nullptr, stream);
}

int ocrl_sa_iter_fwd_ex(const ocrl_sa_dims* d, const void* k, const void* v, const float* slots0,
                        const ocrl_sa_weights* w, float* slots_out, float* attn_vis_out, void* saved, void* workspace,
                        const ocrl_sa_launch_opts* opts, void* stream) {
  int rc = check_dims(d);
  if (rc) return rc;
  if ((rc = check_arch())) return rc;
  if (!k || !v || !slots0 || !w || !slots_out) {
    set_error("sa_iter_fwd: null pointer");
    return OCRL_E_ALIGN;
  }
  if (!aligned16(k) || !aligned16(v) || !aligned16(w->wq) || !aligned16(w->w_ih) || !aligned16(w->w_hh) ||
      !aligned16(w->w1) || !aligned16(w->w2)) {
    set_error("sa_iter_fwd: k, v and weight matrices must be 16-byte aligned");
    return OCRL_E_ALIGN;
  }
  if (d->B == 0) return OCRL_OK;
  if (opts && (opts->variant < OCRL_SA_AUTO || opts->variant > OCRL_SA_FFMA || opts->max_clusters < 0 ||
               (opts->lanes != 0 && opts->lanes != 2 && opts->lanes != 3))) {
    set_error("sa_iter_fwd: bad launch options (variant %d, max_clusters %d, lanes %d)", opts->variant, opts->max_clusters,
              opts->lanes);
    return OCRL_E_SHAPE;
  }
  return sa_iter_fwd_launch(d, k, v, slots0, w, slots_out, attn_vis_out, reinterpret_cast<float*>(saved), workspace,
                            opts, (cudaStream_t)stream);
}

int ocrl_sa_iter_bwd(const ocrl_sa_dims* d, const void* k, const void* v, const void* saved, const ocrl_sa_weights* w,
                     const float* d_slots, const float* d_attn_vis, float* dk, float* dv, float* d_slots0,
                     const ocrl_sa_weight_grads* dw, void* workspace, void* stream) {
  int rc = check_dims(d);
  if (rc) return rc;
  if ((rc = check_arch())) return rc;
  if (!k || !v || !saved || !w || !d_slots || ((dk == nullptr) != (dv == nullptr)) || !d_slots0 || !dw || !workspace) {
    set_error("sa_iter_bwd: null pointer (dk and dv may be NULL together)");
    return OCRL_E_ALIGN;
  }
  if (!aligned16(k) || !aligned16(v) || !aligned16(dk) || !aligned16(dv) || !aligned16(workspace)) {
    set_error("sa_iter_bwd: k, v, dk, dv, workspace must be 16-byte aligned");
    return OCRL_E_ALIGN;
  }
  if (d->B == 0) return OCRL_OK;
  return sa_iter_bwd_launch(d, k, v, reinterpret_cast<const float*>(saved), w, d_slots, d_attn_vis, dk, dv, d_slots0,
                            dw, workspace, (cudaStream_t)stream);
}

}  // extern "C"

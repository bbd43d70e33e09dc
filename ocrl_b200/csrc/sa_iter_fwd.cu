// Host entry of the fused iteration forward (kernels are instantiated in sa_iter_fwd_{f32,bf16}.cu,
// sa_iter_fwd_umma.cu, sa_iter_fwd_pipe.cu, sa_iter_fwd_tc.cu).  Which kernel runs is decided here from the
// dims and the caller's ocrl_sa_launch_opts -- no process-global state -- and reported by ocrl_sa_last_kernel().
#include "sa_iter_fwd.cuh"

namespace ocrl {

int sa_iter_pick_cluster(const ocrl_sa_dims* d);
extern template int sa_iter_fwd_dispatch<float>(const IterFwdArgs&, cudaStream_t);
extern template int sa_iter_fwd_dispatch<__nv_bfloat16>(const IterFwdArgs&, cudaStream_t);
int sa_iter_fwd_tc_dispatch(const IterFwdArgs& a, cudaStream_t s);
int sa_iter_fwd_pipe_dispatch(const IterFwdArgs& a, cudaStream_t s);
int sa_iter_fwd_umma_dispatch(const IterFwdArgs& a, cudaStream_t s);
size_t sa_iter_tc_workspace(const ocrl_sa_dims* d);
const __nv_bfloat16* sa_iter_tc_prepare(const ocrl_sa_dims* d, const ocrl_sa_weights* w, void* workspace,
                                        cudaStream_t stream);

static thread_local const char* g_last_kernel = "";
const char* sa_iter_last_kernel() { return g_last_kernel; }

int sa_iter_fwd_launch(const ocrl_sa_dims* d, const void* k, const void* v, const float* slots0,
                       const ocrl_sa_weights* w, float* slots_out, float* attn_out, float* saved,
                       void* workspace, const ocrl_sa_launch_opts* opts, cudaStream_t stream) {
  ocrl_sa_launch_opts o = {OCRL_SA_AUTO, 0, 0, 0, 0, 0};
  if (opts) o = *opts;
  IterFwdArgs a;
  a.k = k; a.v = v; a.slots0 = slots0; a.w = *w; a.slots_out = slots_out; a.attn_out = attn_out; a.saved = saved;
  a.B = d->B; a.N = d->N; a.D = d->D; a.H = d->H_mlp; a.K = d->K; a.T = d->T;
  a.eps = d->eps; a.ln_eps = d->ln_eps;
  a.trace = nullptr;
  a.wb16 = nullptr;
  a.workspace = workspace;
  a.workspace_bytes = workspace ? sa_iter_tc_workspace(d) : 0;
  a.max_clusters = o.max_clusters;
  a.lanes = o.lanes;
  a.prepared = o.prepared;
  if (o.trace && workspace != nullptr)  // last 4 KB of the workspace: phase timestamps
    a.trace = reinterpret_cast<long long*>(reinterpret_cast<unsigned char*>(workspace) + sa_iter_tc_workspace(d) - 4096);
  a.CL = sa_iter_pick_cluster(d);
  g_last_kernel = "";
  const bool tensor_ok = (d->kv_dtype == OCRL_DT_BF16 && d->math_mode == OCRL_MATH_TENSOR);
  if (o.variant != OCRL_SA_AUTO && o.variant != OCRL_SA_FFMA && !tensor_ok) {
    set_error("sa_iter_fwd: variant %d needs bf16 k/v and math_mode TENSOR", o.variant);
    return OCRL_E_SHAPE;
  }
  auto want = [&](int variant) { return tensor_ok && (o.variant == OCRL_SA_AUTO || o.variant == variant); };
  // a kernel that does not cover the shape returns OCRL_E_SHAPE without launching; unless the caller asked for
  // exactly that kernel (or for strict dispatch) the next one in the list takes over
  auto give_up = [&](int variant) { return o.variant == variant || (o.strict && variant == OCRL_SA_TCGEN05); };
  if (want(OCRL_SA_TCGEN05)) {
    const int rc = sa_iter_fwd_umma_dispatch(a, stream);
    if (rc == OCRL_OK) g_last_kernel = "tcgen05";
    if (rc != OCRL_E_SHAPE || give_up(OCRL_SA_TCGEN05)) return rc;
  }
  if (want(OCRL_SA_PIPE)) {
    const int rc = sa_iter_fwd_pipe_dispatch(a, stream);
    if (rc == OCRL_OK) g_last_kernel = "pipe";
    if (rc != OCRL_E_SHAPE || give_up(OCRL_SA_PIPE)) return rc;
  }
  if (want(OCRL_SA_CLUSTER_TC)) {
    a.wb16 = sa_iter_tc_prepare(d, w, workspace, stream);
    const int rc = sa_iter_fwd_tc_dispatch(a, stream);
    a.wb16 = nullptr;
    if (rc == OCRL_OK) g_last_kernel = "cluster_tc";
    if (rc != OCRL_E_SHAPE || give_up(OCRL_SA_CLUSTER_TC)) return rc;  // shapes whose state does not fit next to the tile ring
  }
  int rc;
  if (d->kv_dtype == OCRL_DT_F32) rc = sa_iter_fwd_dispatch<float>(a, stream);
  else if (d->kv_dtype == OCRL_DT_BF16) rc = sa_iter_fwd_dispatch<__nv_bfloat16>(a, stream);
  else {
    set_error("sa_iter_fwd: unknown kv_dtype %d", d->kv_dtype);
    return OCRL_E_SHAPE;
  }
  if (rc == OCRL_OK) g_last_kernel = "ffma";
  return rc;
}


// Factored form (ocrl_sa_iter_fwd_xhat): tcgen05 kernel only, inference only.
int sa_iter_fwd_xhat_launch(const ocrl_sa_dims* d, const void* xhat, const float* wk, const float* wv, const float* slots0,
                            const ocrl_sa_weights* w, float* slots_out, float* attn_out, void* workspace,
                            const ocrl_sa_launch_opts* opts, cudaStream_t stream) {
  ocrl_sa_launch_opts o = {OCRL_SA_AUTO, 0, 0, 0, 0, 0};
  if (opts) o = *opts;
  g_last_kernel = "";
  if (d->math_mode != OCRL_MATH_TENSOR || (o.variant != OCRL_SA_AUTO && o.variant != OCRL_SA_TCGEN05)) {
    set_error("sa_iter_fwd_xhat: needs math_mode TENSOR and variant AUTO / TCGEN05");
    return OCRL_E_SHAPE;
  }
  IterFwdArgs a;
  a.k = nullptr; a.v = nullptr; a.slots0 = slots0; a.w = *w; a.slots_out = slots_out; a.attn_out = attn_out; a.saved = nullptr;
  a.B = d->B; a.N = d->N; a.D = d->D; a.H = d->H_mlp; a.K = d->K; a.T = d->T;
  a.eps = d->eps; a.ln_eps = d->ln_eps;
  a.trace = nullptr;
  a.wb16 = nullptr;
  a.workspace = workspace;
  a.workspace_bytes = workspace ? sa_iter_tc_workspace(d) : 0;
  a.max_clusters = o.max_clusters;
  a.lanes = o.lanes;
  a.prepared = o.prepared;
  a.xhat = xhat; a.wk = wk; a.wv = wv; a.F = d->C_in;
  if (o.trace && workspace != nullptr)
    a.trace = reinterpret_cast<long long*>(reinterpret_cast<unsigned char*>(workspace) + sa_iter_tc_workspace(d) - 4096);
  a.CL = 8;
  const int rc = sa_iter_fwd_umma_dispatch(a, stream);
  if (rc == OCRL_OK) g_last_kernel = "tcgen05_xhat";
  return rc;
}

}  // namespace ocrl

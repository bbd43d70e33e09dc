// Host entry of the fused iteration forward (kernels are instantiated in sa_iter_fwd_{f32,bf16}.cu)
#include <stdlib.h>

#include "sa_iter_fwd.cuh"

namespace ocrl {

int sa_iter_pick_cluster(const ocrl_sa_dims* d);
extern template int sa_iter_fwd_dispatch<float>(const IterFwdArgs&, cudaStream_t);
extern template int sa_iter_fwd_dispatch<__nv_bfloat16>(const IterFwdArgs&, cudaStream_t);
int sa_iter_fwd_tc_dispatch(const IterFwdArgs& a, cudaStream_t s);
int sa_iter_fwd_pc_dispatch(const IterFwdArgs& a, cudaStream_t s);
int sa_iter_fwd_pipe_dispatch(const IterFwdArgs& a, cudaStream_t s);
size_t sa_iter_tc_workspace(const ocrl_sa_dims* d);
const __nv_bfloat16* sa_iter_tc_prepare(const ocrl_sa_dims* d, const ocrl_sa_weights* w, void* workspace,
                                        cudaStream_t stream);

int sa_iter_fwd_launch(const ocrl_sa_dims* d, const void* k, const void* v, const float* slots0,
                       const ocrl_sa_weights* w, float* slots_out, float* attn_out, float* saved,
                       void* workspace, cudaStream_t stream) {
  IterFwdArgs a;
  a.k = k; a.v = v; a.slots0 = slots0; a.w = *w; a.slots_out = slots_out; a.attn_out = attn_out; a.saved = saved;
  a.B = d->B; a.N = d->N; a.D = d->D; a.H = d->H_mlp; a.K = d->K; a.T = d->T;
  a.eps = d->eps; a.ln_eps = d->ln_eps;
  a.trace = nullptr;
  a.wb16 = nullptr;
  if (workspace != nullptr && getenv("OCRL_SA_TRACE") != nullptr)  // last 4 KB of the workspace: phase timestamps
    a.trace = reinterpret_cast<long long*>(reinterpret_cast<unsigned char*>(workspace) + sa_iter_tc_workspace(d) - 4096);
  a.CL = sa_iter_pick_cluster(d);
  if (d->kv_dtype == OCRL_DT_F32) return sa_iter_fwd_dispatch<float>(a, stream);
  if (d->kv_dtype == OCRL_DT_BF16) {
    if (d->math_mode == OCRL_MATH_TENSOR) {
      // persistent clusters with weight-stationary slot update (K <= 8, inference); OCRL_SA_PC=-1 disables
      // two-engine pipeline (pass of one image under the slot update of another); OCRL_SA_PIPE=-1 disables
      const char* ppv = getenv("OCRL_SA_PIPE");
      if (ppv == nullptr || atoi(ppv) >= 0) {
        const int rc = sa_iter_fwd_pipe_dispatch(a, stream);
        if (rc != OCRL_E_SHAPE) return rc;
      }
      const char* pcv = getenv("OCRL_SA_PC");
      if (pcv == nullptr || atoi(pcv) >= 0) {
        const int rc = sa_iter_fwd_pc_dispatch(a, stream);
        if (rc != OCRL_E_SHAPE) return rc;
      }
      if (getenv("OCRL_SA_CHAIN_FP32") == nullptr) a.wb16 = sa_iter_tc_prepare(d, w, workspace, stream);
      const int rc = sa_iter_fwd_tc_dispatch(a, stream);
      a.wb16 = nullptr;
      if (rc != OCRL_E_SHAPE) return rc;  // shapes whose state does not fit next to the tile ring use the FFMA path
    }
    return sa_iter_fwd_dispatch<__nv_bfloat16>(a, stream);
  }
  set_error("sa_iter_fwd: unknown kv_dtype %d", d->kv_dtype);
  return OCRL_E_SHAPE;
}

}  // namespace ocrl

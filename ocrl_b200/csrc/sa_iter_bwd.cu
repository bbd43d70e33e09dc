// Host entry of the fused iteration backward + the small expansion kernel (weight gradients: sa_iter_wgrad.cu)
// (the cluster kernel is instantiated in sa_iter_bwd_{f32,bf16}.cu)
#include "sa_iter_bwd.cuh"

namespace ocrl {

extern template int sa_iter_bwd_dispatch<float>(const IterBwdArgs&, cudaStream_t);
extern template int sa_iter_bwd_dispatch<__nv_bfloat16>(const IterBwdArgs&, cudaStream_t);

// dk[n][:] = sum_t sum_j dl^t_nj q^t_j ,  dv[n][:] = sum_t sum_j w^t_nj gm^t_j   (warp per token)
template <int D>
__global__ void __launch_bounds__(256) expand_coef_kernel(const float* __restrict__ coef, const float* __restrict__ saved,
                                                          const float* __restrict__ gm, float* __restrict__ dk,
                                                          float* __restrict__ dv, int B, int N, int K, int H, int T) {
  constexpr int NC = D / 64;
  const int lane = threadIdx.x & 31;
  const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const SavedLayout SL(K, D, H);
  for (long long tok = warp_global; tok < (long long)B * N; tok += nwarps) {
    const int img = (int)(tok / N);
    float2 ak[NC], av[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) { ak[c] = make_float2(0.f, 0.f); av[c] = make_float2(0.f, 0.f); }
    for (int t = 0; t < T; ++t) {
      const float* cf = coef + (tok * T + t) * 2 * K;
      const float* qt = saved + ((size_t)img * T + t) * SL.stride() + SL.off_q();
      const float* gt = gm + ((size_t)img * T + t) * K * D;
      for (int j = 0; j < K; ++j) {
        const float dl = __ldg(cf + j), wv = __ldg(cf + K + j);
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const float2 qq = __ldg(reinterpret_cast<const float2*>(qt + j * D + 64 * c + 2 * lane));
          const float2 gg = __ldg(reinterpret_cast<const float2*>(gt + j * D + 64 * c + 2 * lane));
          ak[c].x = fmaf(dl, qq.x, ak[c].x); ak[c].y = fmaf(dl, qq.y, ak[c].y);
          av[c].x = fmaf(wv, gg.x, av[c].x); av[c].y = fmaf(wv, gg.y, av[c].y);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      *reinterpret_cast<float2*>(dk + tok * D + 64 * c + 2 * lane) = ak[c];
      *reinterpret_cast<float2*>(dv + tok * D + 64 * c + 2 * lane) = av[c];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
int sa_iter_pick_cluster(const ocrl_sa_dims* d);
size_t sa_iter_wgrad_part_floats(const ocrl_sa_dims* d);
int sa_iter_wgrad_launch(const ocrl_sa_dims* d, const float* flog, const float* saved, float* part,
                         const ocrl_sa_weight_grads* dw, cudaStream_t stream);

static int bwd_num_clusters(const ocrl_sa_dims* d, int CL) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int ncl = sms / CL;
  if (ncl < 1) ncl = 1;
  if (ncl > d->B) ncl = d->B;
  return ncl;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct BwdWorkspace {
  size_t coef_off, gm_off, flog_off, part_off, total;
};
static BwdWorkspace bwd_ws_layout(const ocrl_sa_dims* d) {
  BwdWorkspace w;
  w.coef_off = 0;
  size_t coef = sizeof(float) * (size_t)d->B * d->N * d->T * 2 * d->K;
  w.gm_off = align_up(coef, 256);
  size_t gm = sizeof(float) * (size_t)d->B * d->T * d->K * d->D;
  w.flog_off = align_up(w.gm_off + gm, 256);
  size_t fl = sizeof(float) * (size_t)d->B * d->T * FLog(d->K, d->D, d->H_mlp).stride();
  w.part_off = align_up(w.flog_off + fl, 256);
  w.total = align_up(w.part_off + sizeof(float) * sa_iter_wgrad_part_floats(d), 256);
  return w;
}

size_t sa_iter_bwd_workspace(const ocrl_sa_dims* d) { return bwd_ws_layout(d).total; }

// where the coefficients and dU / S live inside the workspace (read by the low-rank projection backward)
void sa_iter_bwd_ws_offsets(const ocrl_sa_dims* d, size_t* coef_off, size_t* gm_off) {
  const BwdWorkspace L = bwd_ws_layout(d);
  *coef_off = L.coef_off;
  *gm_off = L.gm_off;
}

int sa_iter_bwd_launch(const ocrl_sa_dims* d, const void* k, const void* v, const float* saved,
                       const ocrl_sa_weights* w, const float* d_slots, const float* d_attn, float* dk, float* dv,
                       float* d_slots0, const ocrl_sa_weight_grads* dw, void* ws, cudaStream_t stream) {
  const BwdWorkspace L = bwd_ws_layout(d);
  unsigned char* base = reinterpret_cast<unsigned char*>(ws);
  IterBwdArgs a;
  a.k = k; a.v = v; a.saved = saved; a.w = *w; a.d_slots = d_slots; a.d_attn = d_attn;
  a.coef = reinterpret_cast<float*>(base + L.coef_off);
  a.gm = reinterpret_cast<float*>(base + L.gm_off);
  a.flog = reinterpret_cast<float*>(base + L.flog_off);
  a.d_slots0 = d_slots0;
  a.B = d->B; a.N = d->N; a.D = d->D; a.H = d->H_mlp; a.K = d->K; a.T = d->T;
  a.eps = d->eps; a.ln_eps = d->ln_eps;
  a.CL = sa_iter_pick_cluster(d);
  a.NCL = bwd_num_clusters(d, a.CL);
  int rc = (d->kv_dtype == OCRL_DT_F32) ? sa_iter_bwd_dispatch<float>(a, stream) : sa_iter_bwd_dispatch<__nv_bfloat16>(a, stream);
  if (rc) return rc;
  const long long tokens = (long long)d->B * d->N;
  const int blocks = (int)((tokens + 7) / 8 < 148 * 8 ? (tokens + 7) / 8 : 148 * 8);
  if (dk != nullptr && dv != nullptr)  // (NULL: the caller takes the coefficients themselves, ocrl_kv_proj_bwd_lowrank)
  switch (d->D) {
    case 64: expand_coef_kernel<64><<<blocks, 256, 0, stream>>>(a.coef, saved, a.gm, dk, dv, d->B, d->N, d->K, d->H_mlp, d->T); ocrl::count_launch(); break;
    case 128: expand_coef_kernel<128><<<blocks, 256, 0, stream>>>(a.coef, saved, a.gm, dk, dv, d->B, d->N, d->K, d->H_mlp, d->T); ocrl::count_launch(); break;
    default: expand_coef_kernel<192><<<blocks, 256, 0, stream>>>(a.coef, saved, a.gm, dk, dv, d->B, d->N, d->K, d->H_mlp, d->T); ocrl::count_launch(); break;
  }
  OCRL_CHECK_CUDA(cudaGetLastError());
  return sa_iter_wgrad_launch(d, a.flog, saved, reinterpret_cast<float*>(base + L.part_off), dw, stream);
}

}  // namespace ocrl

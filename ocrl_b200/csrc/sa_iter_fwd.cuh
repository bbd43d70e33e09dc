// Fused T-iteration slot-attention forward (replaces ocrs/common/slot_attn.py:64-102).
//
// One thread-block cluster per image.  The image's tokens are split across the CTAs of the
// cluster; the K slots live in shared memory (replicated per CTA) for the whole kernel.  Per
// iteration every warp streams its k/v token groups through a private ring of 1-D bulk async
// copies (TMA engine, mbarrier completion), computes the K logits per token with fp32 FFMA,
// reduces them across the warp with a transposed shuffle reduction, takes the K-way softmax,
// and accumulates sum_n (a+eps) v_n and sum_n (a+eps) -- the renormalisation over tokens is
// deferred to one divide per slot.  The partial sums are reduced inside the CTA, reduce-scattered
// across the cluster through distributed shared memory, and the GRUCell + residual MLP + next
// query projection run distributed over the cluster (each CTA owns D/CL output features) with
// DSMEM pushes + cluster barriers between the dependent layers.
#pragma once
#include "slot_math.cuh"

namespace ocrl {

template <typename KV, int D, int KP, int NW_>
struct FwdCfg {
  static constexpr int NW = NW_;
  static constexpr int NT = NW * 32;
  static constexpr int G = (KP <= 8) ? 4 : 2;  // tokens per warp group
  static constexpr int NV = G * KP;            // logits per group (<= 32)
  static constexpr int DPL = D / 32;           // features per lane
  static constexpr int NC = D / 64;            // float2 chunks per lane
  static constexpr int STAGES = (NW_ == 8) ? 3 : 2;
  static constexpr int MIN_CTAS = (NW_ == 8) ? 1 : 2;  // 4-warp CTAs: two clusters share an SM
  static constexpr int GROUP_ELEMS = G * D;
  static constexpr int GROUP_BYTES = GROUP_ELEMS * (int)sizeof(KV);
  static constexpr bool Q_IN_REGS = (KP * DPL <= 48);
  static_assert(NV <= 32, "group too large");
  static_assert(D % 64 == 0, "D must be a multiple of 64");
};

template <typename KV, int D, int KP, int NWT>
__global__ void __launch_bounds__(FwdCfg<KV, D, KP, NWT>::NT, FwdCfg<KV, D, KP, NWT>::MIN_CTAS)
sa_iter_fwd_kernel(const IterFwdArgs a) {
  using Cfg = FwdCfg<KV, D, KP, NWT>;
  constexpr int NW = Cfg::NW, NT = Cfg::NT, G = Cfg::G, NV = Cfg::NV, NC = Cfg::NC, DPL = Cfg::DPL;
  constexpr int STAGES = Cfg::STAGES;

  cg::cluster_group cluster = cg::this_cluster();
  const int CL = a.CL;
  const int rank = (int)cluster.block_rank();
  const int img = blockIdx.x / CL;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = a.K, H = a.H, N = a.N;
  const int DS = D / CL, HS = H / CL;
  const int LMAX = D > H ? D : H;

  // ---- shared memory carve-up (must match sa_iter_fwd_smem_bytes) ---------------------------
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* sp = smem_raw;
  const size_t ring_bytes = (size_t)NW * STAGES * 2 * Cfg::GROUP_BYTES;
  const size_t ured_bytes = (size_t)NW * KP * D * sizeof(float);
  KV* ring = reinterpret_cast<KV*>(sp);
  float* ured = reinterpret_cast<float*>(sp);  // aliases the ring between attention passes
  sp += (ring_bytes > ured_bytes ? ring_bytes : ured_bytes);
  float* s_prev = reinterpret_cast<float*>(sp); sp += sizeof(float) * KP * D;
  float* q_s = reinterpret_cast<float*>(sp); sp += sizeof(float) * KP * D;
  float* rs_buf = reinterpret_cast<float*>(sp); sp += sizeof(float) * KP * D;      // [CL][KP][DS]
  float* upd_full = reinterpret_cast<float*>(sp); sp += sizeof(float) * KP * D;
  float* h_full = reinterpret_cast<float*>(sp); sp += sizeof(float) * KP * D;
  float* hid_full = reinterpret_cast<float*>(sp); sp += sizeof(float) * KP * H;
  float* lnb = reinterpret_cast<float*>(sp); sp += sizeof(float) * KP * LMAX;
  float* gates = reinterpret_cast<float*>(sp); sp += sizeof(float) * 6 * KP * DS;    // raw dots
  float* rs_S = reinterpret_cast<float*>(sp); sp += sizeof(float) * 16 * KP;        // [CL<=16][KP]
  float* sred = reinterpret_cast<float*>(sp); sp += sizeof(float) * NW * 32;
  float* scratch = reinterpret_cast<float*>(sp); sp += sizeof(float) * NW * 64;     // per warp: logits[32], w[32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sp); sp += sizeof(uint64_t) * NW * STAGES;

  const KV* kimg = reinterpret_cast<const KV*>(a.k) + (size_t)img * N * D;
  const KV* vimg = reinterpret_cast<const KV*>(a.v) + (size_t)img * N * D;

  // token range of this CTA, in groups of G tokens
  const int groups_total = (N + G - 1) / G;
  const int gpc = (groups_total + CL - 1) / CL;
  const int g_begin = rank * gpc;
  const int g_end = min(groups_total, g_begin + gpc);
  const int my_groups = max(0, g_end - g_begin);
  // groups of this warp: g_begin + warp, + NW, ...
  const int warp_groups = (my_groups > warp) ? (my_groups - warp + NW - 1) / NW : 0;

  if (lane == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&bars[warp * STAGES + s], 1);
  }
  mbar_fence_init();
  for (int e = tid; e < KP * D; e += NT) {
    const int j = e / D;
    s_prev[e] = (j < K) ? a.slots0[(size_t)img * K * D + e] : 0.f;
  }
  for (int e = tid; e < KP * H; e += NT) hid_full[e] = 0.f;
  for (int e = tid; e < KP * D; e += NT) { upd_full[e] = 0.f; h_full[e] = 0.f; q_s[e] = 0.f; }
  __syncthreads();
  cluster.sync();  // every CTA's shared memory is initialised before any remote push

  // q = W_q LN_s(slots): each CTA computes DS output features and pushes them to all peers
  const SavedLayout SL(K, D, H);
  auto saved_at = [&](int t) { return a.saved + ((size_t)img * a.T + t) * SL.stride(); };
  auto compute_q = [&](int tq) {  // tq: the iteration that will consume this q
    ln_rows(s_prev, a.w.ln_slots_w, a.w.ln_slots_b, lnb, K, D, a.ln_eps, warp, lane, NW);
    __syncthreads();
    rows_dot_jobs<KP, G, NC>(1, DS, DS, [&](int, const float*& W, int& row0, const float*& vec, float*& out) {
      W = a.w.wq; row0 = rank * DS; vec = lnb; out = gates;
    }, warp, lane, NW);
    __syncthreads();
    for (int e = tid; e < K * DS; e += NT) {
      const int j = e / DS, o = e % DS;
      const float val = gates[j * DS + o];
      for (int r = 0; r < CL; ++r) cluster.map_shared_rank(q_s, r)[j * D + rank * DS + o] = val;
      if (a.saved) saved_at(tq)[SL.off_q() + j * D + rank * DS + o] = val;
    }
    cluster.sync();
  };
  compute_q(0);

  float* my_scr = scratch + warp * 64;
  uint64_t* my_bar = bars + warp * STAGES;
  KV* my_ring = ring + (size_t)warp * STAGES * 2 * Cfg::GROUP_ELEMS;
  uint32_t it = 0;  // running count of groups consumed by this warp (ring stage / parity)

  auto issue = [&](int local_group, uint32_t seq) {  // lane 0 only
    const int gidx = g_begin + warp + local_group * NW;
    const int tok0 = gidx * G;
    const int nvalid = min(G, N - tok0);
    const uint32_t bytes = (uint32_t)(nvalid * D * sizeof(KV));
    const int st = seq % STAGES;
    KV* dst = my_ring + (size_t)st * 2 * Cfg::GROUP_ELEMS;
    mbar_expect_tx(&my_bar[st], 2 * bytes);
    bulk_g2s(dst, kimg + (size_t)tok0 * D, bytes, &my_bar[st]);
    bulk_g2s(dst + Cfg::GROUP_ELEMS, vimg + (size_t)tok0 * D, bytes, &my_bar[st]);
  };

  for (int t = 0; t < a.T; ++t) {
    const bool last = (t == a.T - 1);
    if (a.saved) {  // slots entering iteration t, for the backward
      float* sv = saved_at(t);
      for (int e = tid; e < K * DS; e += NT) {
        const int j = e / DS, o = e % DS;
        sv[SL.off_h() + j * D + rank * DS + o] = s_prev[j * D + rank * DS + o];
      }
    }

    // ------------------------------ attention pass over this CTA's tokens --------------------
    float U[KP][DPL];
#pragma unroll
    for (int j = 0; j < KP; ++j)
#pragma unroll
      for (int i = 0; i < DPL; ++i) U[j][i] = 0.f;
    float Sl = 0.f;
    {
      float qr[Cfg::Q_IN_REGS ? KP : 1][Cfg::Q_IN_REGS ? DPL : 1];
      if constexpr (Cfg::Q_IN_REGS) {
#pragma unroll
        for (int j = 0; j < KP; ++j)
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            const float2 x = *reinterpret_cast<const float2*>(q_s + j * D + 64 * c + 2 * lane);
            qr[j][2 * c] = x.x;
            qr[j][2 * c + 1] = x.y;
          }
      }
      if (lane == 0) {
        fence_proxy_async();
        for (int p = 0; p < STAGES && p < warp_groups; ++p) issue(p, it + p);
      }
      for (int lg = 0; lg < warp_groups; ++lg, ++it) {
        const int st = it % STAGES;
        const uint32_t parity = (it / STAGES) & 1u;
        const KV* kb = my_ring + (size_t)st * 2 * Cfg::GROUP_ELEMS;
        const KV* vb = kb + Cfg::GROUP_ELEMS;
        const int tok0 = (g_begin + warp + lg * NW) * G;
        const int nvalid = min(G, N - tok0);
        mbar_wait(&my_bar[st], parity);

        float acc[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) acc[i] = 0.f;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          float2 kk[G];
#pragma unroll
          for (int g = 0; g < G; ++g) kk[g] = Elem<KV>::load2(kb + g * D + 64 * c + 2 * lane);
#pragma unroll
          for (int j = 0; j < KP; ++j) {
            float2 qq;
            if constexpr (Cfg::Q_IN_REGS) {
              qq = make_float2(qr[j][2 * c], qr[j][2 * c + 1]);
            } else {
              qq = *reinterpret_cast<const float2*>(q_s + j * D + 64 * c + 2 * lane);
            }
#pragma unroll
            for (int g = 0; g < G; ++g) acc[g * KP + j] = fmaf(kk[g].x, qq.x, fmaf(kk[g].y, qq.y, acc[g * KP + j]));
          }
        }
        int base;
        xreduce<NV>(acc, lane, base);
        if (XReduce<NV, 16>::primary(lane)) {
#pragma unroll
          for (int i = 0; i < XReduce<NV, 16>::kFinal; ++i) my_scr[base + i] = acc[i];
        }
        __syncwarp();
        {
          // K-way softmax over the slot axis: lane -> (token g, slot j)
          float wgt = 0.f;
          if (lane < NV) {
            const int g = lane / KP, j = lane % KP;
            const float* lg_ = my_scr + g * KP;
            float m = lg_[0];
            for (int jj = 1; jj < K; ++jj) m = fmaxf(m, lg_[jj]);
            float sum = 0.f;
            for (int jj = 0; jj < K; ++jj) sum += __expf(lg_[jj] - m);
            const bool valid = (j < K) && (g < nvalid);
            const float av = valid ? __expf(lg_[j] - m) / sum : 0.f;
            if (valid) {
              wgt = av + a.eps;
              if (last && a.attn_out) a.attn_out[((size_t)img * N + tok0 + g) * K + j] = av;
            }
            Sl += wgt;
          }
          my_scr[32 + lane] = wgt;
        }
        __syncwarp();
#pragma unroll
        for (int g = 0; g < G; ++g) {
          if (g < nvalid) {
            float wj[KP];
#pragma unroll
            for (int j = 0; j < KP; ++j) wj[j] = my_scr[32 + g * KP + j];
#pragma unroll
            for (int c = 0; c < NC; ++c) {
              const float2 vv = Elem<KV>::load2(vb + g * D + 64 * c + 2 * lane);
#pragma unroll
              for (int j = 0; j < KP; ++j) {
                U[j][2 * c] = fmaf(wj[j], vv.x, U[j][2 * c]);
                U[j][2 * c + 1] = fmaf(wj[j], vv.y, U[j][2 * c + 1]);
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0 && lg + STAGES < warp_groups) {
          fence_proxy_async();
          issue(lg + STAGES, it + STAGES);
        }
      }
    }

    // ------------------------------ CTA reduction of the partial sums ------------------------
    __syncthreads();  // every warp is done with the ring (ured aliases it)
#pragma unroll
    for (int j = 0; j < KP; ++j)
#pragma unroll
      for (int c = 0; c < NC; ++c)
        *reinterpret_cast<float2*>(ured + ((size_t)warp * KP + j) * D + 64 * c + 2 * lane) =
            make_float2(U[j][2 * c], U[j][2 * c + 1]);
    sred[warp * 32 + lane] = Sl;
    __syncthreads();
    // push this CTA's partial slices to their owners (reduce-scatter), fixed summation order
    for (int e = tid; e < K * D; e += NT) {
      const int j = e / D, d = e % D;
      float s = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < NW; ++w8) s += ured[((size_t)w8 * KP + j) * D + d];
      const int r = d / DS, o = d % DS;
      cluster.map_shared_rank(rs_buf, r)[(rank * KP + j) * DS + o] = s;
    }
    if (tid < K) {
      float s = 0.f;
      for (int w8 = 0; w8 < NW; ++w8)
        for (int g = 0; g < G; ++g) s += sred[w8 * 32 + g * KP + tid];
      for (int r = 0; r < CL; ++r) cluster.map_shared_rank(rs_S, r)[rank * KP + tid] = s;
    }
    cluster.sync();  // #1: all partial slices have landed
    for (int e = tid; e < K * DS; e += NT) {
      const int j = e / DS, o = e % DS;
      float tot = 0.f, st = 0.f;
      for (int r = 0; r < CL; ++r) {
        tot += rs_buf[(r * KP + j) * DS + o];
        st += rs_S[r * KP + j];
      }
      const float upd = tot / st;
      for (int r = 0; r < CL; ++r) cluster.map_shared_rank(upd_full, r)[j * D + rank * DS + o] = upd;
      if (a.saved) {
        float* sv = saved_at(t);
        sv[SL.off_u() + j * D + rank * DS + o] = upd;
        if (rank == 0 && o == 0) sv[SL.off_s() + j] = st;
      }
    }
    cluster.sync();  // #2: updates[K][D] complete everywhere

    // ------------------------------ GRUCell (this CTA's DS features) -------------------------
    rows_dot_jobs<KP, G, NC>(6, DS, DS, [&](int job, const float*& W, int& row0, const float*& vec, float*& out) {
      const int gsel = job % 3;
      W = (job < 3) ? a.w.w_ih : a.w.w_hh;
      vec = (job < 3) ? upd_full : s_prev;
      row0 = gsel * D + rank * DS;
      out = gates + job * KP * DS;
    }, warp, lane, NW);
    __syncthreads();
    for (int e = tid; e < K * DS; e += NT) {
      const int j = e / DS, o = e % DS;
      const int f = rank * DS + o;
      const float gir = gates[(0 * KP + j) * DS + o] + __ldg(a.w.b_ih + f);
      const float giz = gates[(1 * KP + j) * DS + o] + __ldg(a.w.b_ih + D + f);
      const float gin = gates[(2 * KP + j) * DS + o] + __ldg(a.w.b_ih + 2 * D + f);
      const float ghr = gates[(3 * KP + j) * DS + o] + __ldg(a.w.b_hh + f);
      const float ghz = gates[(4 * KP + j) * DS + o] + __ldg(a.w.b_hh + D + f);
      const float ghn = gates[(5 * KP + j) * DS + o] + __ldg(a.w.b_hh + 2 * D + f);
      const float rg = sigmoidf_(gir + ghr);
      const float zg = sigmoidf_(giz + ghz);
      const float ng = tanhf(gin + rg * ghn);
      const float hn = (1.f - zg) * ng + zg * s_prev[j * D + f];
      for (int r = 0; r < CL; ++r) cluster.map_shared_rank(h_full, r)[j * D + f] = hn;
      if (a.saved) {
        float* sv = saved_at(t);
        sv[SL.off_r() + j * D + f] = rg;
        sv[SL.off_z() + j * D + f] = zg;
        sv[SL.off_n() + j * D + f] = ng;
        sv[SL.off_ghn() + j * D + f] = ghn;
        sv[SL.off_hp() + j * D + f] = hn;
      }
    }
    cluster.sync();  // #3: GRU output complete everywhere

    // ------------------------------ residual MLP ---------------------------------------------
    ln_rows(h_full, a.w.ln_mlp_w, a.w.ln_mlp_b, lnb, K, D, a.ln_eps, warp, lane, NW);
    __syncthreads();
    rows_dot_jobs<KP, G, NC>(1, HS, HS, [&](int, const float*& W, int& row0, const float*& vec, float*& out) {
      W = a.w.w1; row0 = rank * HS; vec = lnb; out = gates;
    }, warp, lane, NW);
    __syncthreads();
    for (int e = tid; e < K * HS; e += NT) {
      const int j = e / HS, o = e % HS;
      const float pre = gates[j * HS + o] + __ldg(a.w.b1 + rank * HS + o);
      const float hv = fmaxf(pre, 0.f);
      for (int r = 0; r < CL; ++r) cluster.map_shared_rank(hid_full, r)[j * H + rank * HS + o] = hv;
      if (a.saved) saved_at(t)[SL.off_pre() + j * H + rank * HS + o] = pre;
    }
    cluster.sync();  // #4: hidden layer complete everywhere
    {
      auto w2job = [&](int, const float*& W, int& row0, const float*& vec, float*& out) {
        W = a.w.w2; row0 = rank * DS; vec = hid_full; out = gates;
      };
      switch (H / 64) {
        case 1: rows_dot_jobs<KP, G, 1>(1, DS, DS, w2job, warp, lane, NW); break;
        case 2: rows_dot_jobs<KP, G, 2>(1, DS, DS, w2job, warp, lane, NW); break;
        case 3: rows_dot_jobs<KP, G, 3>(1, DS, DS, w2job, warp, lane, NW); break;
        case 4: rows_dot_jobs<KP, G, 4>(1, DS, DS, w2job, warp, lane, NW); break;
        default: rows_dot<KP, G>(a.w.w2, H, rank * DS, DS, hid_full, gates, DS, 0, warp, lane, NW); break;
      }
    }
    __syncthreads();
    for (int e = tid; e < K * DS; e += NT) {
      const int j = e / DS, o = e % DS;
      const int f = rank * DS + o;
      const float sn = h_full[j * D + f] + gates[j * DS + o] + __ldg(a.w.b2 + f);
      for (int r = 0; r < CL; ++r) cluster.map_shared_rank(s_prev, r)[j * D + f] = sn;
      if (last) a.slots_out[((size_t)img * K + j) * D + f] = sn;
    }
    cluster.sync();  // #5: new slots complete everywhere
    if (!last) compute_q(t + 1);  // #6
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <typename KV, int D, int KP, int NWT>
static size_t fwd_smem_bytes(int H, int CL) {
  using Cfg = FwdCfg<KV, D, KP, NWT>;
  const size_t ring_bytes = (size_t)Cfg::NW * Cfg::STAGES * 2 * Cfg::GROUP_BYTES;
  const size_t ured_bytes = (size_t)Cfg::NW * KP * D * sizeof(float);
  const int LMAX = D > H ? D : H;
  size_t b = ring_bytes > ured_bytes ? ring_bytes : ured_bytes;
  b += sizeof(float) * ((size_t)KP * D * 5 + (size_t)KP * H + (size_t)KP * LMAX + 6 * (size_t)KP * (D / CL));
  b += sizeof(float) * (16 * KP + Cfg::NW * 32 + Cfg::NW * 64);
  b += sizeof(uint64_t) * Cfg::NW * Cfg::STAGES;
  return b + 128;
}

template <typename KV, int D, int KP, int NWT>
static int launch_fwd_nw(const IterFwdArgs& a, cudaStream_t stream) {
  using Cfg = FwdCfg<KV, D, KP, NWT>;
  auto kern = sa_iter_fwd_kernel<KV, D, KP, NWT>;
  const size_t smem = fwd_smem_bytes<KV, D, KP, NWT>(a.H, a.CL);
  if (smem > 227 * 1024) {
    set_error("sa_iter_fwd: shared memory %zu B exceeds 227 KB (D=%d K=%d H=%d)", smem, D, a.K, a.H);
    return OCRL_E_SHAPE;
  }
  OCRL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (a.CL > 8) OCRL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(a.B * a.CL));
  cfg.blockDim = dim3(Cfg::NT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)a.CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  OCRL_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
  ocrl::count_launch();
  return OCRL_OK;
}

int sa_iter_pick_warps(int D, int K, int CL);

template <typename KV, int D, int KP>
static int launch_fwd(const IterFwdArgs& a, cudaStream_t stream) {
  if constexpr (KP <= 8) {
    if (sa_iter_pick_warps(D, KP, a.CL) == 4) return launch_fwd_nw<KV, D, KP, 4>(a, stream);
  }
  return launch_fwd_nw<KV, D, KP, 8>(a, stream);
}

template <typename KV, int D>
static int dispatch_k(const IterFwdArgs& a, cudaStream_t s) {
  const int K = a.K;
  if (K <= 4) return launch_fwd<KV, D, 4>(a, s);
  if (K <= 6) return launch_fwd<KV, D, 6>(a, s);
  if (K <= 8) return launch_fwd<KV, D, 8>(a, s);
  if (K <= 12) return launch_fwd<KV, D, 12>(a, s);
  if (K <= 16) return launch_fwd<KV, D, 16>(a, s);
  set_error("sa_iter_fwd: num_slots=%d not supported (1..16)", K);
  return OCRL_E_SHAPE;
}

template <typename KV>
int sa_iter_fwd_dispatch(const IterFwdArgs& a, cudaStream_t s) {
  switch (a.D) {
    case 64: return dispatch_k<KV, 64>(a, s);
    case 128: return dispatch_k<KV, 128>(a, s);
    case 192: return dispatch_k<KV, 192>(a, s);
    default:
      set_error("sa_iter_fwd: slot_size=%d not supported (64, 128, 192)", a.D);
      return OCRL_E_SHAPE;
  }
}

}  // namespace ocrl

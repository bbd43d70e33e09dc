// Fused backward of the T-iteration slot-attention loop (autograd of ocrs/common/slot_attn.py:64-102).
//
// Same decomposition as the forward: one thread-block cluster per image, tokens split across the
// CTAs, slot-sized state replicated in shared memory, GRU / MLP / projection backward distributed
// over the cluster (each CTA owns D/CL rows of every weight matrix) with DSMEM reduce-scatter /
// all-gather between dependent layers.  The kernel is persistent: cluster c walks images
// c, c+NCL, ... and accumulates weight gradients into its own partial buffer (no atomics).
//
// The attention logits are RECOMPUTED from k and the saved q_t; nothing of size [N,K] is stored by
// the forward.  Per token and iteration the pass emits only 2K coefficients,
//     dl_nk = a_nk (da_nk - sum_j a_nj da_nj)        and        w_nk = a_nk + eps,
// because  dk_n = sum_t sum_k dl^t_nk q^t_k  and  dv_n = sum_t sum_k w^t_nk (dU^t_k / S^t_k)
// are rank-(T*K) products; a second small kernel expands them into dk, dv once.
#pragma once
#include "slot_math.cuh"

namespace ocrl {

struct WGradLayout {  // flat per-cluster partial weight-gradient buffer (floats)
  int D, H;
  __host__ __device__ WGradLayout(int d, int h) : D(d), H(h) {}
  __host__ __device__ int wq() const { return 0; }
  __host__ __device__ int w_ih() const { return D * D; }
  __host__ __device__ int w_hh() const { return 4 * D * D; }
  __host__ __device__ int b_ih() const { return 7 * D * D; }
  __host__ __device__ int b_hh() const { return 7 * D * D + 3 * D; }
  __host__ __device__ int w1() const { return 7 * D * D + 6 * D; }
  __host__ __device__ int b1() const { return w1() + H * D; }
  __host__ __device__ int w2() const { return b1() + H; }
  __host__ __device__ int b2() const { return w2() + D * H; }
  __host__ __device__ int ln_s_w() const { return b2() + D; }
  __host__ __device__ int ln_s_b() const { return ln_s_w() + D; }
  __host__ __device__ int ln_m_w() const { return ln_s_b() + D; }
  __host__ __device__ int ln_m_b() const { return ln_m_w() + D; }
  __host__ __device__ int total() const { return ln_m_b() + D; }
};

// Per (image, iteration) record of the factors of the weight gradients.  Every weight gradient of the loop is a sum of
// K-row outer products per image and iteration (dW2 = dsn^T y, dW1 = dpre^T mhat, dW_ih = dgate_i^T u, dW_hh = dgate_h^T h,
// dWq = dq^T shat), so the cluster kernel only records the row / column factors it has in shared memory anyway and a
// split-K product over all B T K rows forms the gradients once at the end (sa_iter_wgrad.cu) -- instead of a
// read-modify-write of 1.3 MB of partial gradients through the L2 per image and iteration.  u, h and the MLP
// pre-activation come from the forward's saved state.
struct FLog {
  int K, L;  // L = max(D, H): row pitch
  __host__ __device__ FLog(int k, int d, int h) : K(k), L(d > h ? d : h) {}
  enum { DSN = 0, DPRE = 1, DR = 2, DZ = 3, DN = 4, DNR = 5, DQ = 6, MHAT = 7, SHAT = 8, NFACT = 9 };
  enum { LN_M_W = 0, LN_M_B = 1, LN_S_W = 2, LN_S_B = 3, NVEC = 4 };
  __host__ __device__ size_t fact(int f) const { return (size_t)f * K * L; }
  __host__ __device__ size_t vec(int v) const { return (size_t)NFACT * K * L + (size_t)v * L; }
  __host__ __device__ size_t stride() const { return (size_t)NFACT * K * L + (size_t)NVEC * L; }
};

struct IterBwdArgs {
  const void* k;
  const void* v;
  const float* saved;
  ocrl_sa_weights w;
  const float* d_slots;
  const float* d_attn;  // may be null
  float* coef;          // [B][N][T][2K]
  float* gm;            // [B][T][K][D]   dU_t / S_t
  float* d_slots0;
  float* flog;          // [B][T][FLog::stride()] factors of the weight gradients
  int B, N, D, H, K, T, CL, NCL;
  float eps, ln_eps;
};

template <typename KV, int D, int KP>
struct BwdCfg {
  static constexpr int NW = 8;
  static constexpr int NT = NW * 32;
  static constexpr int G = (KP <= 8) ? 4 : 2;
  static constexpr int NV = G * KP;
  static constexpr int DPL = D / 32;
  static constexpr int NC = D / 64;
  static constexpr int STAGES = 2;
  static constexpr int GROUP_ELEMS = G * D;
  static constexpr int GROUP_BYTES = GROUP_ELEMS * (int)sizeof(KV);
  static constexpr bool Q_IN_REGS = (KP * DPL <= 36);
};

// out[j*ldo + c] (+)= sum_{r<nrows} A[j*lda + r] * W[(row0+r)*ldw + c]   for c < ncols (multiple of 64)
template <int KP, int NW, int NT>
__device__ __forceinline__ void cols_dot(const float* __restrict__ W, int ldw, int row0, int nrows, const float* A,
                                         int lda, float* out, int ldo, int ncols, int K, float* red, bool accumulate,
                                         int tid) {
  const int warp = tid >> 5, lane = tid & 31;
  for (int c0 = 0; c0 < ncols; c0 += 64) {
    float2 acc[KP];
#pragma unroll
    for (int j = 0; j < KP; ++j) acc[j] = make_float2(0.f, 0.f);
    for (int r = warp; r < nrows; r += NW) {
      const float2 wv = __ldg(reinterpret_cast<const float2*>(W + (size_t)(row0 + r) * ldw + c0 + 2 * lane));
#pragma unroll
      for (int j = 0; j < KP; ++j) {
        const float av = A[j * lda + r];
        acc[j].x = fmaf(av, wv.x, acc[j].x);
        acc[j].y = fmaf(av, wv.y, acc[j].y);
      }
    }
#pragma unroll
    for (int j = 0; j < KP; ++j) *reinterpret_cast<float2*>(red + (warp * KP + j) * 64 + 2 * lane) = acc[j];
    __syncthreads();
    for (int e = tid; e < K * 64; e += NT) {
      const int j = e / 64, c = e % 64;
      float s = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < NW; ++w8) s += red[(w8 * KP + j) * 64 + c];
      float* o = out + j * ldo + c0 + c;
      *o = accumulate ? (*o + s) : s;
    }
    __syncthreads();
  }
}

// Wg[(row0+r)*ld + c] += sum_k A[k*lda + r] * Bm[k*ldb + c]   r < nrows, c < ncols (multiple of 64)
template <int NW>
__device__ __forceinline__ void outer_acc(float* __restrict__ Wg, int ld, int row0, int nrows, const float* A, int lda,
                                          const float* Bm, int ldb, int ncols, int K, int warp, int lane) {
  for (int r = warp; r < nrows; r += NW) {
    for (int c = 2 * lane; c < ncols; c += 64) {
      float2* p = reinterpret_cast<float2*>(Wg + (size_t)(row0 + r) * ld + c);
      float2 acc = *p;
      for (int kk = 0; kk < K; ++kk) {
        const float av = A[kk * lda + r];
        const float2 bv = *reinterpret_cast<const float2*>(Bm + kk * ldb + c);
        acc.x = fmaf(av, bv.x, acc.x);
        acc.y = fmaf(av, bv.y, acc.y);
      }
      *p = acc;
    }
  }
}

// per-row LayerNorm statistics of `rows` rows of length L in shared memory -> stats[2*j] = mean, [2*j+1] = rstd
__device__ __forceinline__ void ln_stats(const float* src, float* stats, int rows, int L, float eps, int warp, int lane,
                                         int nwarps) {
  for (int j = warp; j < rows; j += nwarps) {
    float s = 0.f;
    for (int d = lane; d < L; d += 32) s += src[j * L + d];
    const float mean = warp_sum(s) / (float)L;
    float q = 0.f;
    for (int d = lane; d < L; d += 32) {
      const float t = src[j * L + d] - mean;
      q = fmaf(t, t, q);
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)L + eps);
    if (lane == 0) {
      stats[2 * j] = mean;
      stats[2 * j + 1] = rstd;
    }
  }
}

// LayerNorm backward over full rows: dx[j][d] = rstd (g - mean(g) - xh mean(g xh)), g = dy*gamma, xh = (x-mean) rstd
__device__ __forceinline__ void ln_bwd_rows(const float* x, const float* dy, const float* stats,
                                            const float* __restrict__ gamma, float* dx, int rows, int L, int warp,
                                            int lane, int nwarps) {
  for (int j = warp; j < rows; j += nwarps) {
    const float mean = stats[2 * j], rstd = stats[2 * j + 1];
    float s1 = 0.f, s2 = 0.f;
    for (int d = lane; d < L; d += 32) {
      const float g = dy[j * L + d] * __ldg(gamma + d);
      const float xh = (x[j * L + d] - mean) * rstd;
      s1 += g;
      s2 = fmaf(g, xh, s2);
    }
    s1 = warp_sum(s1) / (float)L;
    s2 = warp_sum(s2) / (float)L;
    for (int d = lane; d < L; d += 32) {
      const float g = dy[j * L + d] * __ldg(gamma + d);
      const float xh = (x[j * L + d] - mean) * rstd;
      dx[j * L + d] = rstd * (g - s1 - xh * s2);
    }
  }
}

template <typename KV, int D, int KP>
__global__ void __launch_bounds__(BwdCfg<KV, D, KP>::NT, 1) sa_iter_bwd_kernel(const IterBwdArgs a) {
  using Cfg = BwdCfg<KV, D, KP>;
  constexpr int NW = Cfg::NW, NT = Cfg::NT, G = Cfg::G, NV = Cfg::NV, NC = Cfg::NC, DPL = Cfg::DPL;
  constexpr int STAGES = Cfg::STAGES;

  cg::cluster_group cluster = cg::this_cluster();
  const int CL = a.CL;
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / CL;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = a.K, H = a.H, N = a.N, T = a.T;
  const int DS = D / CL, HS = H / CL;
  const int LMAX = D > H ? D : H;
  const SavedLayout SL(K, D, H);
  const FLog FL(K, D, H);

  // ---- shared memory carve-up (must match bwd_smem_bytes) ---------------------------------------
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* sp = smem_raw;
  const size_t ring_bytes = (size_t)NW * STAGES * 2 * Cfg::GROUP_BYTES;
  const size_t red_bytes = (size_t)NW * KP * 64 * sizeof(float);  // cross-warp reductions go 64 columns at a time
  KV* ring = reinterpret_cast<KV*>(sp);
  float* red = reinterpret_cast<float*>(sp);  // aliases the ring outside the token pass
  sp += (ring_bytes > red_bytes ? ring_bytes : red_bytes);
  auto take = [&](size_t nfloats) { float* p = reinterpret_cast<float*>(sp); sp += sizeof(float) * nfloats; return p; };
  // Slot-sized buffers (KP x D floats each; 12 of them -- with 16 slot rows at D = 192 that is 147 KB).  Buffers whose
  // lifetimes do not overlap share storage: the MLP pre-activation (B1 only) with the all-gathered dU (written from
  // E3a on), the queries (loaded right before the token pass) with the second GRU partial (dead after its push in B4),
  // dU / S (token pass only) with the all-gather target of d mhat / d shat (idle between B3 and E5a).
  float* h_s = take(KP * D);
  float* u_s = take(KP * D);
  float* z_s = take(KP * D);
  float* hp_s = take(KP * D);
  float* pre_s = take(KP * LMAX);  // MLP pre-activation [KP][H] ...
  float* du_full = pre_s;          // ... then the all-gathered dU [KP][D]
  float* dsn = take(KP * D);       // d slots_next (input gradient of this iteration), replicated
  float* dhp = take(KP * D);       // d h'
  float* part = take(KP * LMAX);   // local partial sums before an exchange
  float* part2 = take(KP * D);
  float* q_s = part2;              // the iteration's queries during the token pass
  float* rsA = take(KP * LMAX);    // reduce-scatter receive buffers [CL][KP][slice]
  float* rsB = take(KP * D);
  float* fullA = take(KP * D);     // all-gather receive: d mhat, later d shat
  float* gm_s = fullA;             // dU / S during the token pass
  float* dhg_full = take(KP * D);
  float* dgate = take(6 * KP * DS);  // [dgi_r, dgi_z, dgi_n, dgh_r, dgh_z, dgh_n][KP][DS]
  float* dpre = take(KP * HS);
  float* dq_sl = take(KP * DS);
  float* stat_s = take(2 * KP);
  float* stat_m = take(2 * KP);
  float* S_s = take(KP);
  float* c_s = take(KP);
  float* scratch = take(NW * 96);  // per warp: logits[32], p[32], dl[32]
  sp = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(sp) + 7) & ~uintptr_t(7));
  uint64_t* bars = reinterpret_cast<uint64_t*>(sp);

  const int groups_total = (N + G - 1) / G;
  const int gpc = (groups_total + CL - 1) / CL;
  const int g_begin = rank * gpc;
  const int g_end = min(groups_total, g_begin + gpc);
  const int my_groups = max(0, g_end - g_begin);
  const int warp_groups = (my_groups > warp) ? (my_groups - warp + NW - 1) / NW : 0;

  if (lane == 0)
    for (int s = 0; s < STAGES; ++s) mbar_init(&bars[warp * STAGES + s], 1);
  mbar_fence_init();
  __syncthreads();
  cluster.sync();

  float* my_scr = scratch + warp * 96;
  uint64_t* my_bar = bars + warp * STAGES;
  KV* my_ring = ring + (size_t)warp * STAGES * 2 * Cfg::GROUP_ELEMS;
  uint32_t it = 0;

  auto rs_push = [&](const float* P, int L, int LS, float* rsbuf) {
    for (int e = tid; e < K * L; e += NT) {
      const int j = e / L, d = e % L, r = d / LS, o = d % LS;
      cluster.map_shared_rank(rsbuf, r)[(rank * KP + j) * LS + o] = P[j * L + d];
    }
  };
  auto rs_sum = [&](const float* rsbuf, int LS, int j, int o) {
    float s = 0.f;
    for (int r = 0; r < CL; ++r) s += rsbuf[(r * KP + j) * LS + o];
    return s;
  };
  auto ag_push = [&](float* full, int L, int LS, int j, int o, float val) {
    for (int r = 0; r < CL; ++r) cluster.map_shared_rank(full, r)[j * L + rank * LS + o] = val;
  };

  for (int img = cid; img < a.B; img += a.NCL) {
    const KV* kimg = reinterpret_cast<const KV*>(a.k) + (size_t)img * N * D;
    const KV* vimg = reinterpret_cast<const KV*>(a.v) + (size_t)img * N * D;
    auto issue = [&](int local_group, uint32_t seq) {  // lane 0 only
      const int tok0 = (g_begin + warp + local_group * NW) * G;
      const int nvalid = min(G, N - tok0);
      const uint32_t bytes = (uint32_t)(nvalid * D * sizeof(KV));
      const int st = seq % STAGES;
      KV* dst = my_ring + (size_t)st * 2 * Cfg::GROUP_ELEMS;
      mbar_expect_tx(&my_bar[st], 2 * bytes);
      bulk_g2s(dst, kimg + (size_t)tok0 * D, bytes, &my_bar[st]);
      bulk_g2s(dst + Cfg::GROUP_ELEMS, vimg + (size_t)tok0 * D, bytes, &my_bar[st]);
    };

    for (int e = tid; e < KP * D; e += NT) {
      const int j = e / D;
      dsn[e] = (j < K) ? a.d_slots[(size_t)img * K * D + e] : 0.f;
    }

    for (int t = T - 1; t >= 0; --t) {
      const bool last = (t == T - 1);
      const float* sv = a.saved + ((size_t)img * T + t) * SL.stride();
      float* fl = a.flog + ((size_t)img * T + t) * FL.stride();  // this CTA records its feature slice of every factor
      // ---- load the saved forward state of iteration t (replicated) ----------------------------
      for (int e = tid; e < KP * D; e += NT) {
        const bool ok = e < K * D;
        h_s[e] = ok ? sv[SL.off_h() + e] : 0.f;
        u_s[e] = ok ? sv[SL.off_u() + e] : 0.f;
        z_s[e] = ok ? sv[SL.off_z() + e] : 0.f;
        hp_s[e] = ok ? sv[SL.off_hp() + e] : 0.f;
      }
      for (int e = tid; e < KP * H; e += NT) pre_s[e] = (e < K * H) ? sv[SL.off_pre() + e] : 0.f;
      if (tid < KP) S_s[tid] = (tid < K) ? sv[SL.off_s() + tid] : 1.f;
      __syncthreads();
      ln_stats(h_s, stat_s, K, D, a.ln_eps, warp, lane, NW);
      ln_stats(hp_s, stat_m, K, D, a.ln_eps, warp, lane, NW);
      __syncthreads();

      // ---- B1: MLP output layer.  partial dy[k][:] = sum_{d in slice} dsn[k][d] W2[d][:] ----------
      cols_dot<KP, NW, NT>(a.w.w2, H, rank * DS, DS, dsn + rank * DS, D, part, H, H, K, red, false, tid);
      rs_push(part, H, HS, rsA);
      cluster.sync();  // E1
      for (int e = tid; e < K * HS; e += NT) {
        const int j = e / HS, o = e % HS;
        const float dy = rs_sum(rsA, HS, j, o);
        dpre[j * HS + o] = (pre_s[j * H + rank * HS + o] > 0.f) ? dy : 0.f;
      }
      __syncthreads();
      __syncthreads();
      for (int e = tid; e < K * DS; e += NT) {  // dW2 = dsn^T relu(pre), db2 = column sums of dsn
        const int j = e / DS, o = e % DS;
        fl[FL.fact(FLog::DSN) + j * FL.L + rank * DS + o] = dsn[j * D + rank * DS + o];
      }
      // ---- B2: MLP hidden layer.  partial d mhat[k][:] = sum_{j in slice(H)} dpre[k][j] W1[j][:] -
      cols_dot<KP, NW, NT>(a.w.w1, D, rank * HS, HS, dpre, HS, part, D, D, K, red, false, tid);
      rs_push(part, D, DS, rsB);
      for (int e = tid; e < K * HS; e += NT) {  // dW1 = dpre^T mhat, db1 = column sums of dpre
        const int j = e / HS, o = e % HS;
        fl[FL.fact(FLog::DPRE) + j * FL.L + rank * HS + o] = dpre[j * HS + o];
      }
      for (int e = tid; e < K * DS; e += NT) {
        const int j = e / DS, o = e % DS;
        const int d = rank * DS + o;  // LN_m(h'), the input of the MLP
        fl[FL.fact(FLog::MHAT) + j * FL.L + d] =
            (hp_s[j * D + d] - stat_m[2 * j]) * stat_m[2 * j + 1] * __ldg(a.w.ln_mlp_w + d) + __ldg(a.w.ln_mlp_b + d);
      }
      cluster.sync();  // E2a
      for (int e = tid; e < K * DS; e += NT) {
        const int j = e / DS, o = e % DS;
        ag_push(fullA, D, DS, j, o, rs_sum(rsB, DS, j, o));
      }
      cluster.sync();  // E2b: d mhat complete everywhere
      // ---- B3: LayerNorm (norm_mlp) backward; dh' = dsn + dLN ------------------------------------
      for (int o = tid; o < DS; o += NT) {
        const int d = rank * DS + o;
        float gw = 0.f, gb = 0.f;
        for (int j = 0; j < K; ++j) {
          const float dyv = fullA[j * D + d];
          gw = fmaf(dyv, (hp_s[j * D + d] - stat_m[2 * j]) * stat_m[2 * j + 1], gw);
          gb += dyv;
        }
        fl[FL.vec(FLog::LN_M_W) + d] = gw;
        fl[FL.vec(FLog::LN_M_B) + d] = gb;
      }
      ln_bwd_rows(hp_s, fullA, stat_m, a.w.ln_mlp_w, dhp, K, D, warp, lane, NW);
      __syncthreads();
      for (int e = tid; e < K * D; e += NT) dhp[e] += dsn[e];
      __syncthreads();
      // ---- B4: GRUCell backward on this CTA's features --------------------------------------------
      for (int e = tid; e < K * DS; e += NT) {
        const int j = e / DS, o = e % DS;
        const int f = j * D + rank * DS + o;
        const float rg = sv[SL.off_r() + f], ng = sv[SL.off_n() + f], ghn = sv[SL.off_ghn() + f];
        const float zg = z_s[f], dh = dhp[f];
        const float dn_pre = dh * (1.f - zg) * (1.f - ng * ng);
        const float dz_pre = dh * (h_s[f] - ng) * zg * (1.f - zg);
        const float dr_pre = dn_pre * ghn * rg * (1.f - rg);
        dgate[(0 * KP + j) * DS + o] = dr_pre;
        dgate[(1 * KP + j) * DS + o] = dz_pre;
        dgate[(2 * KP + j) * DS + o] = dn_pre;
        dgate[(3 * KP + j) * DS + o] = dr_pre;
        dgate[(4 * KP + j) * DS + o] = dz_pre;
        dgate[(5 * KP + j) * DS + o] = dn_pre * rg;
      }
      __syncthreads();
      for (int gsel = 0; gsel < 3; ++gsel) {
        cols_dot<KP, NW, NT>(a.w.w_ih, D, gsel * D + rank * DS, DS, dgate + (gsel)*KP * DS, DS, part, D, D, K, red,
                             gsel > 0, tid);
        cols_dot<KP, NW, NT>(a.w.w_hh, D, gsel * D + rank * DS, DS, dgate + (3 + gsel) * KP * DS, DS, part2, D, D, K,
                             red, gsel > 0, tid);
      }
      rs_push(part, D, DS, rsA);
      rs_push(part2, D, DS, rsB);
      for (int e = tid; e < K * DS; e += NT) {  // dW_ih = [dr dz dn]^T u, dW_hh = [dr dz dn r]^T h, biases = column sums
        const int j = e / DS, o = e % DS;
        const size_t at = (size_t)j * FL.L + rank * DS + o;
        fl[FL.fact(FLog::DR) + at] = dgate[(0 * KP + j) * DS + o];
        fl[FL.fact(FLog::DZ) + at] = dgate[(1 * KP + j) * DS + o];
        fl[FL.fact(FLog::DN) + at] = dgate[(2 * KP + j) * DS + o];
        fl[FL.fact(FLog::DNR) + at] = dgate[(5 * KP + j) * DS + o];
      }
      cluster.sync();  // E3a
      for (int e = tid; e < K * DS; e += NT) {
        const int j = e / DS, o = e % DS;
        ag_push(du_full, D, DS, j, o, rs_sum(rsA, DS, j, o));
        ag_push(dhg_full, D, DS, j, o, rs_sum(rsB, DS, j, o));
      }
      cluster.sync();  // E3b: dU and W_hh^T dgh complete everywhere
      // ---- B5: token pass -------------------------------------------------------------------------
      for (int e = tid; e < KP * D; e += NT) {
        const int j = e / D;
        gm_s[e] = (j < K) ? du_full[e] / S_s[j] : 0.f;
        q_s[e] = (e < K * D) ? sv[SL.off_q() + e] : 0.f;
      }
      __syncthreads();
      for (int j = warp; j < K; j += NW) {
        float s = 0.f;
        for (int d = lane; d < D; d += 32) s = fmaf(u_s[j * D + d], gm_s[j * D + d], s);
        s = warp_sum(s);
        if (lane == 0) c_s[j] = s;
      }
      for (int e = tid; e < K * DS; e += NT) {
        const int j = e / DS, o = e % DS;
        a.gm[(((size_t)img * T + t) * K + j) * D + rank * DS + o] = gm_s[j * D + rank * DS + o];
      }
      __syncthreads();

      float dq[KP][DPL];
#pragma unroll
      for (int j = 0; j < KP; ++j)
#pragma unroll
        for (int i = 0; i < DPL; ++i) dq[j][i] = 0.f;
      {
        float qr[Cfg::Q_IN_REGS ? KP : 1][Cfg::Q_IN_REGS ? DPL : 1];
        float gr[Cfg::Q_IN_REGS ? KP : 1][Cfg::Q_IN_REGS ? DPL : 1];
        if constexpr (Cfg::Q_IN_REGS) {
#pragma unroll
          for (int j = 0; j < KP; ++j)
#pragma unroll
            for (int c = 0; c < NC; ++c) {
              const float2 x = *reinterpret_cast<const float2*>(q_s + j * D + 64 * c + 2 * lane);
              const float2 y = *reinterpret_cast<const float2*>(gm_s + j * D + 64 * c + 2 * lane);
              qr[j][2 * c] = x.x; qr[j][2 * c + 1] = x.y;
              gr[j][2 * c] = y.x; gr[j][2 * c + 1] = y.y;
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

          float acc1[NV], acc2[NV];
#pragma unroll
          for (int i = 0; i < NV; ++i) { acc1[i] = 0.f; acc2[i] = 0.f; }
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            float2 kk[G], vv[G];
#pragma unroll
            for (int g = 0; g < G; ++g) {
              kk[g] = Elem<KV>::load2(kb + g * D + 64 * c + 2 * lane);
              vv[g] = Elem<KV>::load2(vb + g * D + 64 * c + 2 * lane);
            }
#pragma unroll
            for (int j = 0; j < KP; ++j) {
              float2 qq, gg;
              if constexpr (Cfg::Q_IN_REGS) {
                qq = make_float2(qr[j][2 * c], qr[j][2 * c + 1]);
                gg = make_float2(gr[j][2 * c], gr[j][2 * c + 1]);
              } else {
                qq = *reinterpret_cast<const float2*>(q_s + j * D + 64 * c + 2 * lane);
                gg = *reinterpret_cast<const float2*>(gm_s + j * D + 64 * c + 2 * lane);
              }
#pragma unroll
              for (int g = 0; g < G; ++g) {
                acc1[g * KP + j] = fmaf(kk[g].x, qq.x, fmaf(kk[g].y, qq.y, acc1[g * KP + j]));
                acc2[g * KP + j] = fmaf(vv[g].x, gg.x, fmaf(vv[g].y, gg.y, acc2[g * KP + j]));
              }
            }
          }
          int base;
          xreduce<NV>(acc1, lane, base);
          xreduce<NV>(acc2, lane, base);
          if (XReduce<NV, 16>::primary(lane)) {
#pragma unroll
            for (int i = 0; i < XReduce<NV, 16>::kFinal; ++i) {
              my_scr[base + i] = acc1[i];
              my_scr[32 + base + i] = acc2[i];
            }
          }
          __syncwarp();
          {
            float dl = 0.f;
            if (lane < NV) {
              const int g = lane / KP, j = lane % KP;
              const float* lg_ = my_scr + g * KP;
              const float* pg_ = my_scr + 32 + g * KP;
              float m = lg_[0];
              for (int jj = 1; jj < K; ++jj) m = fmaxf(m, lg_[jj]);
              float sum = 0.f;
              for (int jj = 0; jj < K; ++jj) sum += __expf(lg_[jj] - m);
              const float inv = 1.f / sum;
              const bool valid = (j < K) && (g < nvalid);
              const float* dat = (last && a.d_attn && valid) ? a.d_attn + ((size_t)img * N + tok0 + g) * K : nullptr;
              float dot = 0.f, da_own = 0.f, a_own = 0.f;
              for (int jj = 0; jj < K; ++jj) {
                const float av = __expf(lg_[jj] - m) * inv;
                float da = pg_[jj] - c_s[jj];
                if (dat) da += __ldg(dat + jj);
                dot = fmaf(av, da, dot);
                if (jj == j) { da_own = da; a_own = av; }
              }
              if (valid) {
                dl = a_own * (da_own - dot);
                float* cf = a.coef + (((size_t)img * N + tok0 + g) * T + t) * 2 * K;
                cf[j] = dl;
                cf[K + j] = a_own + a.eps;
              }
            }
            my_scr[64 + lane] = dl;
          }
          __syncwarp();
#pragma unroll
          for (int g = 0; g < G; ++g) {
            if (g < nvalid) {
              float dj[KP];
#pragma unroll
              for (int j = 0; j < KP; ++j) dj[j] = my_scr[64 + g * KP + j];
#pragma unroll
              for (int c = 0; c < NC; ++c) {
                const float2 kk = Elem<KV>::load2(kb + g * D + 64 * c + 2 * lane);
#pragma unroll
                for (int j = 0; j < KP; ++j) {
                  dq[j][2 * c] = fmaf(dj[j], kk.x, dq[j][2 * c]);
                  dq[j][2 * c + 1] = fmaf(dj[j], kk.y, dq[j][2 * c + 1]);
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
      // ---- CTA reduction of dq (64 columns at a time), reduce-scatter across the cluster -------------
      __syncthreads();
#pragma unroll
      for (int c = 0; c < NC; ++c) {
#pragma unroll
        for (int j = 0; j < KP; ++j)
          *reinterpret_cast<float2*>(red + ((size_t)warp * KP + j) * 64 + 2 * lane) = make_float2(dq[j][2 * c], dq[j][2 * c + 1]);
        __syncthreads();
        for (int e = tid; e < K * 64; e += NT) {
          const int j = e / 64, d = 64 * c + e % 64;
          float s = 0.f;
#pragma unroll
          for (int w8 = 0; w8 < NW; ++w8) s += red[((size_t)w8 * KP + j) * 64 + e % 64];
          const int r = d / DS, o = d % DS;
          cluster.map_shared_rank(rsA, r)[(rank * KP + j) * DS + o] = s;
        }
        __syncthreads();
      }
      cluster.sync();  // E4
      for (int e = tid; e < K * DS; e += NT) {
        const int j = e / DS, o = e % DS;
        dq_sl[j * DS + o] = rs_sum(rsA, DS, j, o);
      }
      __syncthreads();
      // ---- B6: query projection backward -----------------------------------------------------------
      cols_dot<KP, NW, NT>(a.w.wq, D, rank * DS, DS, dq_sl, DS, part, D, D, K, red, false, tid);
      rs_push(part, D, DS, rsB);
      for (int e = tid; e < K * DS; e += NT) {  // dWq = dq^T shat
        const int j = e / DS, o = e % DS;
        fl[FL.fact(FLog::DQ) + j * FL.L + rank * DS + o] = dq_sl[j * DS + o];
        const int d = rank * DS + o;  // LN_s(h), the input of the query projection
        fl[FL.fact(FLog::SHAT) + j * FL.L + d] =
            (h_s[j * D + d] - stat_s[2 * j]) * stat_s[2 * j + 1] * __ldg(a.w.ln_slots_w + d) + __ldg(a.w.ln_slots_b + d);
      }
      cluster.sync();  // E5a
      for (int e = tid; e < K * DS; e += NT) {
        const int j = e / DS, o = e % DS;
        ag_push(fullA, D, DS, j, o, rs_sum(rsB, DS, j, o));
      }
      cluster.sync();  // E5b: d shat complete everywhere
      // ---- B7: LayerNorm (norm_slots) backward; gradient of the slots entering this iteration -------
      for (int o = tid; o < DS; o += NT) {
        const int d = rank * DS + o;
        float gw = 0.f, gb = 0.f;
        for (int j = 0; j < K; ++j) {
          const float dyv = fullA[j * D + d];
          gw = fmaf(dyv, (h_s[j * D + d] - stat_s[2 * j]) * stat_s[2 * j + 1], gw);
          gb += dyv;
        }
        fl[FL.vec(FLog::LN_S_W) + d] = gw;
        fl[FL.vec(FLog::LN_S_B) + d] = gb;
      }
      ln_bwd_rows(h_s, fullA, stat_s, a.w.ln_slots_w, part, K, D, warp, lane, NW);
      __syncthreads();
      for (int e = tid; e < K * D; e += NT) dsn[e] = dhp[e] * z_s[e] + dhg_full[e] + part[e];
      __syncthreads();
    }
    for (int e = tid; e < K * DS; e += NT) {
      const int j = e / DS, o = e % DS;
      a.d_slots0[((size_t)img * K + j) * D + rank * DS + o] = dsn[j * D + rank * DS + o];
    }
    __syncthreads();
  }
}

template <typename KV, int D, int KP>
static size_t bwd_smem_bytes(int H, int CL) {
  using Cfg = BwdCfg<KV, D, KP>;
  const size_t ring_bytes = (size_t)Cfg::NW * Cfg::STAGES * 2 * Cfg::GROUP_BYTES;
  const size_t red_bytes = (size_t)Cfg::NW * KP * 64 * sizeof(float);
  const int LMAX = D > H ? D : H;
  const int DS = D / CL, HS = H / CL;
  size_t f = (size_t)KP * D * 10 + 3 * (size_t)KP * LMAX + 6 * (size_t)KP * DS + (size_t)KP * HS +
             (size_t)KP * DS + 4 * (size_t)KP + 2 * (size_t)KP + Cfg::NW * 96;
  return (ring_bytes > red_bytes ? ring_bytes : red_bytes) + sizeof(float) * f + sizeof(uint64_t) * Cfg::NW * Cfg::STAGES +
         128;
}

template <typename KV, int D, int KP>
static int launch_bwd(const IterBwdArgs& a_in, cudaStream_t stream) {
  using Cfg = BwdCfg<KV, D, KP>;
  auto kern = sa_iter_bwd_kernel<KV, D, KP>;
  IterBwdArgs a = a_in;
  size_t smem = bwd_smem_bytes<KV, D, KP>(a.H, a.CL);
  // many slot rows: the per-CTA slices of the gate gradients shrink with the cluster size -- take a larger cluster than
  // the token count alone would ask for until the state fits (K = 16, D = 192 needs clusters of four)
  while (smem > 227 * 1024 && a.CL < 8 && D % (2 * a.CL) == 0 && a.H % (2 * a.CL) == 0) {
    const int total_ctas = a.NCL * a.CL;
    a.CL *= 2;
    a.NCL = max(1, min(a.B, max(total_ctas, 148) / a.CL));
    smem = bwd_smem_bytes<KV, D, KP>(a.H, a.CL);
  }
  if (smem > 227 * 1024) {
    set_error("sa_iter_bwd: shared memory %zu B exceeds 227 KB (D=%d K=%d H=%d); backward: slot-sized state does not fit", smem,
              D, a.K, a.H);
    return OCRL_E_SHAPE;
  }
  OCRL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (a.CL > 8) OCRL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(a.NCL * a.CL));
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

template <typename KV, int D>
static int bwd_dispatch_k(const IterBwdArgs& a, cudaStream_t s) {
  const int K = a.K;
  if (K <= 4) return launch_bwd<KV, D, 4>(a, s);
  if (K <= 6) return launch_bwd<KV, D, 6>(a, s);
  if (K <= 8) return launch_bwd<KV, D, 8>(a, s);
  if (K <= 12) return launch_bwd<KV, D, 12>(a, s);
  return launch_bwd<KV, D, 16>(a, s);
}

template <typename KV>
int sa_iter_bwd_dispatch(const IterBwdArgs& a, cudaStream_t s) {
  switch (a.D) {
    case 64: return bwd_dispatch_k<KV, 64>(a, s);
    case 128: return bwd_dispatch_k<KV, 128>(a, s);
    case 192: return bwd_dispatch_k<KV, 192>(a, s);
    default:
      set_error("sa_iter_bwd: slot_size=%d not supported (64, 128, 192)", a.D);
      return OCRL_E_SHAPE;
  }
}

}  // namespace ocrl

// Fused T-iteration slot-attention forward, tensor-core variant for bf16 k/v
// (replaces ocrs/common/slot_attn.py:64-102; selected when kv_dtype = bf16, math_mode = TENSOR).
//
// Same decomposition as sa_iter_fwd.cuh (one cluster per image, slots resident in shared memory,
// GRU/MLP distributed over the cluster through DSMEM) but the two token contractions run on the
// tensor cores, flash-attention style with the roles of queries and keys swapped:
//   logits^T [slots x 16 tokens] = q [16 x D] . k_tile^T      (A = q fragments, B = k via ldmatrix)
//   softmax over the slot axis = across the 8 lane-rows of the accumulator fragment (3 shuffles)
//   U^T [D x slots] += v_tile^T [D x 16 tokens] . w^T         (A = v via ldmatrix.trans, B = the softmax
//                                                             accumulator re-used in registers as bf16)
// k/v tiles of 16 tokens stream global -> shared with 16-byte cp.async into rows padded by 16 bytes
// (conflict-free ldmatrix), a 3-stage ring private to each warp: no block barrier in the token loop.
// Accumulation is fp32; the token-sum S uses the same bf16-rounded weights as the numerator.
#include <stdlib.h>

#include "slot_math.cuh"

namespace ocrl {

template <int D, int KP, int NPW_, int GT_ = 16>
struct TcCfg {
  static constexpr int NPW = NPW_;              // warps that stream tokens (each owns a tile ring)
  static constexpr int NW = 8;                  // all warps take part in the slot update
  static constexpr int NT = NW * 32;
  static constexpr int GT = GT_;                // tokens per warp group (16, or 8 with twice the warps)
  static constexpr int PITCH = D * 2 + 16;      // bytes per padded bf16 row
  static constexpr int TILE_BYTES = GT * PITCH; // k or v of one group
  static constexpr int STAGE_BYTES = 2 * TILE_BYTES;
  static constexpr int STAGES = (GT_ == 8) ? 2 : ((KP > 8 && D >= 192) ? 2 : 3);
  static constexpr int NKS = D / 16;            // k-steps of the logits / m-tiles of U^T
  static constexpr int NSL = KP / 8;            // slot n-tiles (1 or 2)
  static constexpr int RB = (KP <= 8) ? 4 : 2;  // rows per batch in the slot-update matvecs
  static constexpr int CHUNKS_PER_ROW = D * 2 / 16;
  static_assert(KP == 8 || KP == 16, "slots are padded to 8 or 16");
  static_assert(D % 64 == 0, "D must be a multiple of 64");
};

template <int D, int KP, int NWT, int GTT>
__global__ void __launch_bounds__(TcCfg<D, KP, NWT, GTT>::NT, 1) sa_iter_fwd_tc_kernel(const IterFwdArgs a) {
  using Cfg = TcCfg<D, KP, NWT, GTT>;
  constexpr int NW = Cfg::NW, NPW = Cfg::NPW, NT = Cfg::NT, GT = Cfg::GT, PITCH = Cfg::PITCH, STAGES = Cfg::STAGES;
  constexpr int NKS = Cfg::NKS, NSL = Cfg::NSL, RB = Cfg::RB, NC = D / 64;
  typedef __nv_bfloat16 bf16;

  cg::cluster_group cluster = cg::this_cluster();
  const int CL = a.CL;
  const int rank = (int)cluster.block_rank();
  const int img = blockIdx.x / CL;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g8 = lane >> 2, t4 = lane & 3;
  const int K = a.K, H = a.H, N = a.N;
  const int DS = D / CL, HS = H / CL;
  const int LMAX = D > H ? D : H;

  // ---- shared memory carve-up (must match tc_smem_bytes) ----------------------------------------
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* sp = smem_raw;
  const size_t ring_bytes = (size_t)NPW * STAGES * Cfg::STAGE_BYTES;
  const size_t ured_bytes = (size_t)NPW * KP * D * sizeof(float);
  unsigned char* ring = sp;                    // k/v tiles in the token pass, staged weights in the slot update
  float* wstage = reinterpret_cast<float*>(sp);
  sp += ring_bytes;
  float* ured = reinterpret_cast<float*>(sp);  // per-warp partial U; later the raw dot products ("gates")
  {
    const size_t gates_bytes = sizeof(float) * 6 * (size_t)KP * DS;
    sp += (ured_bytes > gates_bytes ? ured_bytes : gates_bytes);
  }
  unsigned char* qb = sp; sp += 16 * PITCH;    // q as bf16, 16 padded rows (rows >= K are zero)
  auto take = [&](size_t n) { float* p = reinterpret_cast<float*>(sp); sp += sizeof(float) * n; return p; };
  float* s_prev = take(KP * D);
  float* q_s = take(KP * D);
  float* rs_buf = take(KP * D);
  float* upd_full = take(KP * D);
  float* h_full = take(KP * D);
  float* hid_full = take(KP * H);
  float* lnb = rs_buf;  // LayerNorm output; live only while rs_buf is not (see the barriers around them)
  float* gates = ured;
  float* rs_S = take(16 * KP);
  float* sred = take(NPW * KP);
  float* lnp = take(4 * D);            // norm_slots w,b  norm_mlp w,b
  float* bsl = take(6 * DS + HS + DS); // bias slices: b_ih[3][DS], b_hh[3][DS], b1[HS], b2[DS]
  sp = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(sp) + 7) & ~uintptr_t(7));
  uint64_t* wbar = reinterpret_cast<uint64_t*>(sp);  // weights landed: 0 = GRU, 1 = W1, 2 = W2, 3 = Wq
  // the CTA's weight slices fit in the idle ring -> stream them with bulk copies under the exchanges
  const size_t gru_floats = (size_t)DS * D, w1_floats = (size_t)HS * D, w2_floats = (size_t)DS * H;
  const bool w_smem = (a.wb16 == nullptr) && sizeof(float) * (6 * gru_floats + w1_floats + w2_floats) <= ring_bytes;
  float* ws_gru = wstage;                 // [ih_r, ih_z, ih_n, hh_r, hh_z, hh_n][DS][D]
  float* ws_w1 = ws_gru + 6 * gru_floats; // [HS][D]
  float* ws_w2 = ws_w1 + w1_floats;       // [DS][H]
  float* ws_wq = wstage;                  // [DS][D], re-uses the first GRU block once the GRU is done
  // tensor-core slot update: bf16 weights straight from L2, bf16 hi/lo vector pairs staged in the idle ring
  const bool use_mma = (a.wb16 != nullptr);
  const int LP = LMAX + 8;
  bf16* vst0h = reinterpret_cast<bf16*>(ring);
  bf16* vst0l = vst0h + KP * LP;
  bf16* vst1h = vst0l + KP * LP;
  bf16* vst1l = vst1h + KP * LP;
  const bf16* wb_q = a.wb16;
  const bf16* wb_ih = wb_q + (size_t)D * D;
  const bf16* wb_hh = wb_ih + (size_t)3 * D * D;
  const bf16* wb_1 = wb_hh + (size_t)3 * D * D;
  const bf16* wb_2 = wb_1 + (size_t)H * D;

  const bf16* kimg = reinterpret_cast<const bf16*>(a.k) + (size_t)img * N * D;
  const bf16* vimg = reinterpret_cast<const bf16*>(a.v) + (size_t)img * N * D;

  const int groups_total = (N + GT - 1) / GT;
  const int gpc = (groups_total + CL - 1) / CL;
  const int g_begin = rank * gpc;
  const int g_end = min(groups_total, g_begin + gpc);
  const int my_groups = max(0, g_end - g_begin);
  const int warp_groups = (warp < NPW && my_groups > warp) ? (my_groups - warp + NPW - 1) / NPW : 0;

  for (int e = tid; e < KP * D; e += NT) {
    const int j = e / D;
    s_prev[e] = (j < K) ? a.slots0[(size_t)img * K * D + e] : 0.f;
    upd_full[e] = 0.f; h_full[e] = 0.f; q_s[e] = 0.f;
  }
  for (int e = tid; e < KP * H; e += NT) hid_full[e] = 0.f;
  for (int e = tid; e < KP * D; e += NT) lnb[e] = 0.f;
  for (int e = tid; e < 16 * PITCH / 4; e += NT) reinterpret_cast<uint32_t*>(qb)[e] = 0u;
  for (int e = tid; e < D; e += NT) {
    lnp[e] = __ldg(a.w.ln_slots_w + e); lnp[D + e] = __ldg(a.w.ln_slots_b + e);
    lnp[2 * D + e] = __ldg(a.w.ln_mlp_w + e); lnp[3 * D + e] = __ldg(a.w.ln_mlp_b + e);
  }
  for (int e = tid; e < 3 * DS; e += NT) {
    const int g = e / DS, o = e % DS;
    bsl[e] = __ldg(a.w.b_ih + g * D + rank * DS + o);
    bsl[3 * DS + e] = __ldg(a.w.b_hh + g * D + rank * DS + o);
  }
  for (int e = tid; e < HS; e += NT) bsl[6 * DS + e] = __ldg(a.w.b1 + rank * HS + e);
  for (int e = tid; e < DS; e += NT) bsl[6 * DS + HS + e] = __ldg(a.w.b2 + rank * DS + e);
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&wbar[i], 1);
  }
  mbar_fence_init();
  __syncthreads();
  if (w_smem && tid == 0) {  // query weights for the first projection
    fence_proxy_async();
    mbar_expect_tx(&wbar[3], (uint32_t)(sizeof(float) * gru_floats));
    bulk_g2s(ws_wq, a.w.wq + (size_t)rank * DS * D, (uint32_t)(sizeof(float) * gru_floats), &wbar[3]);
  }
  cluster.sync();

  int trace_i = 0;
  auto TRACE = [&]() {
    if (a.trace != nullptr && blockIdx.x == 0 && tid == 0) a.trace[trace_i] = clock64();
    ++trace_i;
  };
  TRACE();
  const SavedLayout SL(K, D, H);
  auto saved_at = [&](int t) { return a.saved + ((size_t)img * a.T + t) * SL.stride(); };

  auto compute_q = [&](int tq) {  // tq-th use of the staged query weights
    ln_rows_fast<D>(s_prev, lnp, lnp + D, lnb, K, a.ln_eps, warp, lane, NW);
    __syncthreads();
    DotDesc d;
    d.vec0 = d.vec1 = lnb; d.out0 = gates;
    d.njobs = 1; d.split = 1; d.split_mod = 1; d.row_stride = 0; d.nrows = DS;
    d.out_stride = 0; d.ldo = DS;
    if (use_mma) {
      stage_vec_hilo(lnb, K, KP, D, vst0h, vst0l, LP, tid, NT);
      __syncthreads();
      rows_mma_jobs<KP, D / 32>(wb_q, rank * DS, 0, 1, DS, vst0h, vst0l, LP, gates, 0, DS, warp, lane, NW);
    } else if (w_smem) {
      mbar_wait(&wbar[3], (uint32_t)(tq & 1));
      d.W0 = d.W1 = ws_wq; d.row_base = 0;
      rows_dot_desc<KP, RB, NC, 2, false>(d, warp, lane, NW);
    } else {
      d.W0 = d.W1 = a.w.wq; d.row_base = rank * DS;
      rows_dot_desc<KP, RB, NC, 2>(d, warp, lane, NW);
    }
    __syncthreads();
    for (int e = tid; e < K * DS; e += NT) {
      const int j = e / DS, o = e % DS;
      const float val = gates[j * DS + o];
      for (int r = 0; r < CL; ++r) cluster.map_shared_rank(q_s, r)[j * D + rank * DS + o] = val;
      if (a.saved) saved_at(tq)[SL.off_q() + j * D + rank * DS + o] = val;
    }
    cluster.sync();
    // bf16 copy of the complete q for the tensor-core pass (rows >= K stay zero)
    for (int e = tid; e < K * (D / 2); e += NT) {
      const int j = e / (D / 2), c2 = e % (D / 2);
      const float2 x = *reinterpret_cast<const float2*>(q_s + j * D + 2 * c2);
      *reinterpret_cast<uint32_t*>(qb + j * PITCH + 4 * c2) = pack_bf16x2(x.x, x.y);
    }
    __syncthreads();
  };
  compute_q(0);

  unsigned char* my_ring = ring + (size_t)warp * STAGES * Cfg::STAGE_BYTES;
  // softmax scratch of a pass warp lives in its own (still unused) slice of the partial-U buffer
  float* my_lg = ured + (size_t)(warp < NPW ? warp : 0) * KP * D;  // logits [16 tokens][KP] fp32
  bf16* my_wb = reinterpret_cast<bf16*>(my_lg + GT * KP);            // weights [KP][16 tokens] bf16

  // all 32 lanes copy one 16-token group of k and v into ring stage `st` (zero-fill past N).
  // Token rows are contiguous in global memory, so chunk c of the tile sits at byte 16*c; in shared
  // memory every row is padded by 16 bytes (conflict-free ldmatrix).
  int dst_off[GT * Cfg::CHUNKS_PER_ROW / 32];
#pragma unroll
  for (int i = 0; i < GT * Cfg::CHUNKS_PER_ROW / 32; ++i) {
    const int c = lane + 32 * i;
    dst_off[i] = c * 16 + (c / Cfg::CHUNKS_PER_ROW) * 16;
  }
  auto issue = [&](int local_group, int st) {
    const int tok0 = (g_begin + warp + local_group * NPW) * GT;
    unsigned char* kd = my_ring + (size_t)st * Cfg::STAGE_BYTES;
    unsigned char* vd = kd + Cfg::TILE_BYTES;
    const unsigned char* ks = reinterpret_cast<const unsigned char*>(kimg + (size_t)tok0 * D) + lane * 16;
    const unsigned char* vs = reinterpret_cast<const unsigned char*>(vimg + (size_t)tok0 * D) + lane * 16;
    if (tok0 + GT <= N) {
#pragma unroll
      for (int i = 0; i < GT * Cfg::CHUNKS_PER_ROW / 32; ++i) {
        cp_async16(kd + dst_off[i], ks + 512 * i, 16);
        cp_async16(vd + dst_off[i], vs + 512 * i, 16);
      }
    } else {  // ragged tail: rows past N are zero-filled
#pragma unroll
      for (int i = 0; i < GT * Cfg::CHUNKS_PER_ROW / 32; ++i) {
        const int row = (lane + 32 * i) / Cfg::CHUNKS_PER_ROW;
        const bool ok = (tok0 + row) < N;
        cp_async16(kd + dst_off[i], ok ? ks + 512 * i : reinterpret_cast<const unsigned char*>(kimg), ok ? 16 : 0);
        cp_async16(vd + dst_off[i], ok ? vs + 512 * i : reinterpret_cast<const unsigned char*>(vimg), ok ? 16 : 0);
      }
    }
  };

  for (int t = 0; t < a.T; ++t) {
    const bool last = (t == a.T - 1);
    if (a.saved) {
      float* sv = saved_at(t);
      for (int e = tid; e < K * DS; e += NT) {
        const int j = e / DS, o = e % DS;
        sv[SL.off_h() + j * D + rank * DS + o] = s_prev[j * D + rank * DS + o];
      }
    }

    TRACE();  // [1 + 9t] start of token pass
    // ------------------------------ token pass on the tensor cores ---------------------------------
    float UT[NSL][NKS][4];  // U^T fragments: rows d = 16*mt + g8 (+8), cols slots 8*sl + 2*t4 + {0,1}
#pragma unroll
    for (int sl = 0; sl < NSL; ++sl)
#pragma unroll
      for (int mt = 0; mt < NKS; ++mt)
#pragma unroll
        for (int i = 0; i < 4; ++i) UT[sl][mt][i] = 0.f;
    float Sl[KP] = {};  // lanes 0-15: sum over this lane's tokens of the (bf16-rounded) weights of every slot
    {
      uint32_t qa[NKS][4];  // A fragments of q (16 slot rows x D)
      {
        const int row = (lane & 7) + ((lane >> 3) & 1) * 8;
        const int colb = (lane >> 4) * 16;
#pragma unroll
        for (int ks = 0; ks < NKS; ++ks) ldmatrix_x4(qa[ks], qb + row * PITCH + ks * 32 + colb);
      }
      // per-lane byte offset inside a k / v tile for the x4 loads (same for both operands)
      const int frag_off = ((lane & 7) + (lane >> 4) * 8) * PITCH + ((lane >> 3) & 1) * 16;
      const int frag_off8 = (lane & 7) * PITCH + ((lane >> 3) & 1) * 16;  // x2 loads of an 8-token tile
      (void)frag_off; (void)frag_off8;

      for (int p = 0; p < STAGES - 1; ++p) {
        if (p < warp_groups) issue(p, p % STAGES);
        cp_async_commit();
      }
      for (int lg = 0; lg < warp_groups; ++lg) {
        {
          const int nxt = lg + STAGES - 1;
          if (nxt < warp_groups) issue(nxt, nxt % STAGES);
          cp_async_commit();
        }
        const bool ptrace = (a.trace != nullptr && blockIdx.x == 0 && tid == 0 && t == 0 && lg < 6);
        if (ptrace) a.trace[128 + lg * 5 + 0] = clock64();
        cp_async_wait<STAGES - 1>();
        __syncwarp();
        if (ptrace) a.trace[128 + lg * 5 + 1] = clock64();
        const unsigned char* kb = my_ring + (size_t)(lg % STAGES) * Cfg::STAGE_BYTES;
        const unsigned char* vb = kb + Cfg::TILE_BYTES;
        const int tok0 = (g_begin + warp + lg * NPW) * GT;

        // logits^T: 8-token blocks (two per 16-token group), four interleaved accumulator chains for ILP
        constexpr int NBLK = GT / 8;
        float lgc[4][NBLK][4] = {};
#pragma unroll
        for (int ks = 0; ks < NKS; ++ks) {
          if constexpr (GT == 16) {
            uint32_t kf[4];
            ldmatrix_x4(kf, kb + frag_off + ks * 32);
            mma_bf16_16816(lgc[ks & 3][0], qa[ks], kf[0], kf[1]);
            mma_bf16_16816(lgc[ks & 3][1], qa[ks], kf[2], kf[3]);
          } else {
            uint32_t kf[2];
            ldmatrix_x2(kf, kb + frag_off8 + ks * 32);
            mma_bf16_16816(lgc[ks & 3][0], qa[ks], kf[0], kf[1]);
          }
        }
        if (ptrace) a.trace[128 + lg * 5 + 2] = clock64();
        // K-way softmax over the slot axis.  The accumulator fragments are transposed through a small
        // per-warp scratch so that one lane owns one token: max / exp / sum need no shuffles.
#pragma unroll
        for (int nb = 0; nb < NBLK; ++nb) {
          float x[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) x[i] = (lgc[0][nb][i] + lgc[1][nb][i]) + (lgc[2][nb][i] + lgc[3][nb][i]);
          const int tk = nb * 8 + 2 * t4;
          my_lg[tk * KP + g8] = x[0];
          my_lg[(tk + 1) * KP + g8] = x[1];
          if constexpr (NSL > 1) {
            my_lg[tk * KP + g8 + 8] = x[2];
            my_lg[(tk + 1) * KP + g8 + 8] = x[3];
          }
        }
        __syncwarp();
        if (lane < GT) {
          float l[KP];
#pragma unroll
          for (int j4 = 0; j4 < KP / 4; ++j4) {
            const float4 v4 = *reinterpret_cast<const float4*>(my_lg + lane * KP + 4 * j4);
            l[4 * j4] = v4.x; l[4 * j4 + 1] = v4.y; l[4 * j4 + 2] = v4.z; l[4 * j4 + 3] = v4.w;
          }
          float m = l[0];
#pragma unroll
          for (int j = 1; j < KP; ++j) m = (j < K) ? fmaxf(m, l[j]) : m;
          float sum = 0.f;
#pragma unroll
          for (int j = 0; j < KP; ++j) {
            l[j] = (j < K) ? __expf(l[j] - m) : 0.f;
            sum += l[j];
          }
          const float inv = __fdividef(1.f, sum);
          const bool tok_ok = (tok0 + lane) < N;
          float* ao = (last && a.attn_out && tok_ok) ? a.attn_out + ((size_t)img * N + tok0 + lane) * K : nullptr;
#pragma unroll
          for (int j = 0; j < KP; ++j) {
            const float av = l[j] * inv;
            if (ao != nullptr && j < K) ao[j] = av;
            // weight a + eps, rounded to bf16 once and used for both the numerator and the token sum
            const bf16 wq = __float2bfloat16_rn((tok_ok && j < K) ? av + a.eps : 0.f);
            my_wb[j * GT + lane] = wq;
            Sl[j] += __bfloat162float(wq);
          }
        }
        __syncwarp();
        uint32_t wb[NSL][2];  // B fragments of w^T for the U^T product: [slot tile][token half]
#pragma unroll
        for (int sl = 0; sl < NSL; ++sl) {
          wb[sl][0] = *reinterpret_cast<const uint32_t*>(my_wb + (8 * sl + g8) * GT + 2 * t4);
          wb[sl][1] = (GT == 16) ? *reinterpret_cast<const uint32_t*>(my_wb + (8 * sl + g8) * GT + (GT - 8) + 2 * t4) : 0u;
        }
        if (ptrace) a.trace[128 + lg * 5 + 3] = clock64();
        // U^T += v_tile^T . w^T
#pragma unroll
        for (int mt = 0; mt < NKS; ++mt) {
          if constexpr (GT == 16) {
            uint32_t vf[4];
            ldmatrix_x4_trans(vf, vb + frag_off + mt * 32);
            // ldmatrix order: [tok 0-7, d 0-7], [tok 0-7, d 8-15], [tok 8-15, d 0-7], [tok 8-15, d 8-15]
            // mma A order:    a0 = (d 0-7, tok 0-7), a1 = (d 8-15, tok 0-7), a2 = (d 0-7, tok 8-15), a3 = (d 8-15, tok 8-15)
            const uint32_t af[4] = {vf[0], vf[1], vf[2], vf[3]};
#pragma unroll
            for (int sl = 0; sl < NSL; ++sl) mma_bf16_16816(UT[sl][mt], af, wb[sl][0], wb[sl][1]);
          } else {
            uint32_t vf[2];  // [tok 0-7, d 0-7], [tok 0-7, d 8-15]: the k = 8 form of the MMA
            ldmatrix_x2_trans(vf, vb + frag_off8 + mt * 32);
#pragma unroll
            for (int sl = 0; sl < NSL; ++sl) mma_bf16_1688(UT[sl][mt], vf[0], vf[1], wb[sl][0]);
          }
        }
        __syncwarp();  // every lane is done with this stage before it is refilled
        if (ptrace) a.trace[128 + lg * 5 + 4] = clock64();
      }
      cp_async_wait<0>();
    }

    // ------------------------------ CTA reduction of the partial sums -------------------------------
    TRACE();  // token pass done (this warp)
    TRACE();
    if (warp < NPW) {
#pragma unroll
    for (int sl = 0; sl < NSL; ++sl)
#pragma unroll
      for (int mt = 0; mt < NKS; ++mt) {
        const int s0 = sl * 8 + 2 * t4, d0 = mt * 16 + g8;
        float* u = ured + (size_t)warp * KP * D;
        u[s0 * D + d0] = UT[sl][mt][0];
        u[(s0 + 1) * D + d0] = UT[sl][mt][1];
        u[s0 * D + d0 + 8] = UT[sl][mt][2];
        u[(s0 + 1) * D + d0 + 8] = UT[sl][mt][3];
      }
#pragma unroll
    for (int j = 0; j < KP; ++j) {
      float s = Sl[j];
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
      if (lane == 0) sred[warp * KP + j] = s;
    }
    }
    __syncthreads();  // partial sums visible; every warp is done with the ring
    if (w_smem && tid == 0) {  // stream this CTA's GRU / MLP weight slices into the idle ring
      fence_proxy_async();
      const uint32_t gb = (uint32_t)(sizeof(float) * gru_floats);
      mbar_expect_tx(&wbar[0], 6 * gb);
      for (int gsel = 0; gsel < 3; ++gsel) {
        bulk_g2s(ws_gru + gsel * gru_floats, a.w.w_ih + ((size_t)gsel * D + rank * DS) * D, gb, &wbar[0]);
        bulk_g2s(ws_gru + (3 + gsel) * gru_floats, a.w.w_hh + ((size_t)gsel * D + rank * DS) * D, gb, &wbar[0]);
      }
      mbar_expect_tx(&wbar[1], (uint32_t)(sizeof(float) * w1_floats));
      bulk_g2s(ws_w1, a.w.w1 + (size_t)rank * HS * D, (uint32_t)(sizeof(float) * w1_floats), &wbar[1]);
      mbar_expect_tx(&wbar[2], (uint32_t)(sizeof(float) * w2_floats));
      bulk_g2s(ws_w2, a.w.w2 + (size_t)rank * DS * H, (uint32_t)(sizeof(float) * w2_floats), &wbar[2]);
    }
    for (int e = tid; e < K * D; e += NT) {
      const int j = e / D, d = e % D;
      float s = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < NPW; ++w8) s += ured[((size_t)w8 * KP + j) * D + d];
      const int r = d / DS, o = d % DS;
      cluster.map_shared_rank(rs_buf, r)[(rank * KP + j) * DS + o] = s;
    }
    if (tid < K) {
      float s = 0.f;
      for (int w8 = 0; w8 < NPW; ++w8) s += sred[w8 * KP + tid];
      for (int r = 0; r < CL; ++r) cluster.map_shared_rank(rs_S, r)[rank * KP + tid] = s;
    }
    TRACE();
    cluster.sync();  // #1
    TRACE();
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
    cluster.sync();  // #2
    TRACE();

    // ------------------------------ GRUCell (this CTA's DS features) ---------------------------------
    {
      DotDesc d;
      d.vec0 = upd_full; d.vec1 = s_prev; d.out0 = gates;
      d.njobs = 6; d.split = 3; d.split_mod = 3; d.nrows = DS; d.out_stride = KP * DS; d.ldo = DS;
      if (use_mma) {
        stage_vec_hilo(upd_full, K, KP, D, vst0h, vst0l, LP, tid, NT);
        stage_vec_hilo(s_prev, K, KP, D, vst1h, vst1l, LP, tid, NT);
        __syncthreads();
        rows_mma_jobs<KP, D / 32>(wb_ih, rank * DS, D, 3, DS, vst0h, vst0l, LP, gates, KP * DS, DS, warp, lane, NW);
        rows_mma_jobs<KP, D / 32>(wb_hh, rank * DS, D, 3, DS, vst1h, vst1l, LP, gates + 3 * KP * DS, KP * DS, DS,
                                  (warp + 4) % NW, lane, NW);
      } else if (w_smem) {
        mbar_wait(&wbar[0], (uint32_t)(t & 1));
        d.W0 = ws_gru; d.W1 = ws_gru + 3 * gru_floats; d.row_base = 0; d.row_stride = DS;
        rows_dot_desc<KP, RB, NC, 2, false>(d, warp, lane, NW);
      } else {
        d.W0 = a.w.w_ih; d.W1 = a.w.w_hh; d.row_base = rank * DS; d.row_stride = D;
        rows_dot_desc<KP, RB, NC, 2>(d, warp, lane, NW);
      }
    }
    __syncthreads();
    if (w_smem && !last && tid == 0) {  // the first GRU block is free: fetch the query weights behind it
      fence_proxy_async();
      mbar_expect_tx(&wbar[3], (uint32_t)(sizeof(float) * gru_floats));
      bulk_g2s(ws_wq, a.w.wq + (size_t)rank * DS * D, (uint32_t)(sizeof(float) * gru_floats), &wbar[3]);
    }
    for (int e = tid; e < K * DS; e += NT) {
      const int j = e / DS, o = e % DS;
      const int f = rank * DS + o;
      const float gir = gates[(0 * KP + j) * DS + o] + bsl[o];
      const float giz = gates[(1 * KP + j) * DS + o] + bsl[DS + o];
      const float gin = gates[(2 * KP + j) * DS + o] + bsl[2 * DS + o];
      const float ghr = gates[(3 * KP + j) * DS + o] + bsl[3 * DS + o];
      const float ghz = gates[(4 * KP + j) * DS + o] + bsl[4 * DS + o];
      const float ghn = gates[(5 * KP + j) * DS + o] + bsl[5 * DS + o];
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
    cluster.sync();  // #3
    TRACE();

    // ------------------------------ residual MLP ------------------------------------------------------
    ln_rows_fast<D>(h_full, lnp + 2 * D, lnp + 3 * D, lnb, K, a.ln_eps, warp, lane, NW);
    __syncthreads();
    {
      DotDesc d;
      d.vec0 = d.vec1 = lnb; d.out0 = gates;
      d.njobs = 1; d.split = 1; d.split_mod = 1; d.row_stride = 0; d.nrows = HS; d.out_stride = 0; d.ldo = HS;
      if (use_mma) {
        stage_vec_hilo(lnb, K, KP, D, vst0h, vst0l, LP, tid, NT);
        __syncthreads();
        rows_mma_jobs<KP, D / 32>(wb_1, rank * HS, 0, 1, HS, vst0h, vst0l, LP, gates, 0, HS, warp, lane, NW);
      } else if (w_smem) {
        mbar_wait(&wbar[1], (uint32_t)(t & 1));
        d.W0 = d.W1 = ws_w1; d.row_base = 0;
        rows_dot_desc<KP, RB, NC, 2, false>(d, warp, lane, NW);
      } else {
        d.W0 = d.W1 = a.w.w1; d.row_base = rank * HS;
        rows_dot_desc<KP, RB, NC, 2>(d, warp, lane, NW);
      }
    }
    __syncthreads();
    for (int e = tid; e < K * HS; e += NT) {
      const int j = e / HS, o = e % HS;
      const float pre = gates[j * HS + o] + bsl[6 * DS + o];
      const float hv = fmaxf(pre, 0.f);
      for (int r = 0; r < CL; ++r) cluster.map_shared_rank(hid_full, r)[j * H + rank * HS + o] = hv;
      if (a.saved) saved_at(t)[SL.off_pre() + j * H + rank * HS + o] = pre;
    }
    cluster.sync();  // #4
    TRACE();
    {
      DotDesc d;
      d.vec0 = d.vec1 = hid_full; d.out0 = gates;
      d.njobs = 1; d.split = 1; d.split_mod = 1; d.row_stride = 0; d.nrows = DS; d.out_stride = 0; d.ldo = DS;
      if (use_mma) {
        stage_vec_hilo(hid_full, K, KP, H, vst0h, vst0l, LP, tid, NT);
        __syncthreads();
        rows_mma_jobs_len<KP>(H, wb_2, rank * DS, 0, 1, DS, vst0h, vst0l, LP, gates, 0, DS, warp, lane, NW);
      } else if (w_smem) {
        mbar_wait(&wbar[2], (uint32_t)(t & 1));
        d.W0 = d.W1 = ws_w2; d.row_base = 0;
        rows_dot_desc_len<KP, RB, 2, false>(d, H, warp, lane, NW);
      } else {
        d.W0 = d.W1 = a.w.w2; d.row_base = rank * DS;
        rows_dot_desc_len<KP, RB, 2>(d, H, warp, lane, NW);
      }
    }
    __syncthreads();
    for (int e = tid; e < K * DS; e += NT) {
      const int j = e / DS, o = e % DS;
      const int f = rank * DS + o;
      const float sn = h_full[j * D + f] + gates[j * DS + o] + bsl[6 * DS + HS + o];
      for (int r = 0; r < CL; ++r) cluster.map_shared_rank(s_prev, r)[j * D + f] = sn;
      if (last) a.slots_out[((size_t)img * K + j) * D + f] = sn;
    }
    cluster.sync();  // #5
    TRACE();
    if (!last) compute_q(t + 1);  // #6
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <int D, int KP, int NWT, int GTT>
static size_t tc_smem_bytes(int H, int CL) {
  using Cfg = TcCfg<D, KP, NWT, GTT>;
  const size_t ring_bytes = (size_t)Cfg::NPW * Cfg::STAGES * Cfg::STAGE_BYTES;
  const size_t ured_bytes = (size_t)Cfg::NPW * KP * D * sizeof(float);
  const int LMAX = D > H ? D : H;
  const size_t gates_bytes = sizeof(float) * 6 * (size_t)KP * (D / CL);
  size_t b = ring_bytes + (ured_bytes > gates_bytes ? ured_bytes : gates_bytes);
  b += 16 * Cfg::PITCH;
  b += sizeof(float) * ((size_t)KP * D * 5 + (size_t)KP * H + 16 * KP + Cfg::NPW * KP + 4 * D +
                        6 * (D / CL) + H / CL + D / CL);
  b += 8 + 4 * sizeof(uint64_t);
  return b + 128;
}

template <int D, int KP, int NWT, int GTT>
static int launch_tc(const IterFwdArgs& a, cudaStream_t stream) {
  using Cfg = TcCfg<D, KP, NWT, GTT>;
  auto kern = sa_iter_fwd_tc_kernel<D, KP, NWT, GTT>;
  const size_t smem = tc_smem_bytes<D, KP, NWT, GTT>(a.H, a.CL);
  if (smem > 227 * 1024) {
    set_error("sa_iter_fwd(tensor): shared memory %zu B exceeds 227 KB (D=%d K=%d H=%d)", smem, D, a.K, a.H);
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

template <int D>
static int tc_dispatch_k(const IterFwdArgs& a, cudaStream_t s) {
  // token-pass warps x tokens per group: ring = NPW * stages * 2 * GT * (2D+16) bytes.  Eight warps on
  // 8-token groups hide the in-order latencies of the softmax / MMA chain better than four on 16.
  // measured on B200: 0.27 ms (8-token groups, 8 warps) vs 0.51 ms (16-token groups) at B=64, N=4096
  if (D >= 192) {
    if (a.K <= 8) return launch_tc<D, 8, 8, 8>(a, s);
    return launch_tc<D, 16, 4, 16>(a, s);
  } else {
    if (a.K <= 8) return launch_tc<D, 8, 8, 16>(a, s);
    return launch_tc<D, 16, 8, 16>(a, s);
  }
}

// fp32 -> bf16 copies of the slot-update weights, laid out [wq | w_ih | w_hh | w1 | w2]
__global__ void iter_tc_prep_kernel(ocrl_sa_weights w, __nv_bfloat16* out, int D, int H) {
  const int n_q = D * D, n_g = 3 * D * D, n_1 = H * D;
  const int total = n_q + 2 * n_g + 2 * n_1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    float v;
    if (i < n_q) v = w.wq[i];
    else if (i < n_q + n_g) v = w.w_ih[i - n_q];
    else if (i < n_q + 2 * n_g) v = w.w_hh[i - n_q - n_g];
    else if (i < n_q + 2 * n_g + n_1) v = w.w1[i - n_q - 2 * n_g];
    else v = w.w2[i - n_q - 2 * n_g - n_1];
    out[i] = __float2bfloat16_rn(v);
  }
}

size_t sa_iter_tc_workspace(const ocrl_sa_dims* d) {
  // bf16 weight copies of the per-image cluster kernel, or the tensor-memory blocks of the tcgen05 kernel
  // ([<= 8 ranks][2][max(D,H)/2 words][128 rows] + folded-LayerNorm constants), whichever is larger; + alignment, trace slots
  const size_t tc = sizeof(__nv_bfloat16) * ((size_t)7 * d->D * d->D + (size_t)2 * d->H_mlp * d->D);
  const int lx = d->D > d->H_mlp ? d->D : d->H_mlp;
  const size_t um = (size_t)8 * 2 * (lx / 2) * 128 * 4 + (size_t)(2 * d->H_mlp + 2 * d->D) * 4;
  return (tc > um ? tc : um) + 256 + 4096;
}

// converts the weights into `workspace`; returns the bf16 pointer (256-byte aligned) or null
const __nv_bfloat16* sa_iter_tc_prepare(const ocrl_sa_dims* d, const ocrl_sa_weights* w, void* workspace,
                                        cudaStream_t stream) {
  if (workspace == nullptr) return nullptr;
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  iter_tc_prep_kernel<<<148, 256, 0, stream>>>(*w, out, d->D, d->H_mlp);
  ocrl::count_launch();
  return out;
}

int sa_iter_fwd_tc_dispatch(const IterFwdArgs& a, cudaStream_t s) {
  switch (a.D) {
    case 64: return tc_dispatch_k<64>(a, s);
    case 128: return tc_dispatch_k<128>(a, s);
    case 192: return tc_dispatch_k<192>(a, s);
    default:
      set_error("sa_iter_fwd(tensor): slot_size=%d not supported (64, 128, 192)", a.D);
      return OCRL_E_SHAPE;
  }
}

}  // namespace ocrl

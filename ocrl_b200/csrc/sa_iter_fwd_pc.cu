// Persistent-cluster form of the fused T-iteration loop (slot_attn.py:64-102 of the reference) for bf16 k/v.
//
// Work split
//   * One thread-block cluster (CL CTAs) owns one image at a time; every CTA runs NG independent
//     "groups" of 8 warps, each group working on its own image, so that the latency-bound slot update of
//     one image overlaps the bandwidth-bound token pass of another on the same SM.
//   * Clusters are persistent: group g of cluster c walks images (newest first, they are the ones the
//     projection kernel left in L2).  Only NG * (#clusters) images are in flight, so the k/v working set
//     of passes 2..T stays inside the 126 MB L2 and is read from HBM once.
//   * The CTA's slice of every slot-update weight (W_q, W_ih, W_hh, W_1, W_2: D/CL output features each)
//     is converted to bf16 once and stays in shared memory for the whole kernel (weight-stationary).
//
// Token pass (per group: 4 "logit" warps + 4 "U" warps, no block barrier inside the pass)
//   k/v tiles of TOK tokens travel global -> shared with ONE tensor-map TMA copy per operand (the [rows][D]
//   matrix is described as [rows][D/64][64] so that a whole tile is a single box; the unit applies the
//   128-byte swizzle, which makes the dense rows conflict-free for ldmatrix).  Per-row bulk copies were
//   measured to cost ~90 cycles of TMA issue each and bound the first version of this kernel.  The ring of
//   S stages keeps running across iteration and image boundaries, so the first tiles of the next pass land
//   while the slot update runs.  Tile j is handled by logit warp j%4 and U warp j%4:
//   logit warp:  logits[16 tok x 8 slots] = k_tile . q^T  (mma.sync m16n8k16, q fragments in registers),
//                softmax over the slot axis with two xor-shuffles, attn_vis store in the last iteration,
//                w = a + eps rounded to bf16, transposed with movmatrix into the B-fragment layout and
//                handed to the U warp through 256 bytes of shared memory (one mbarrier arrive).
//   U warp:      U^T[d x slot] += v_tile^T . w over all D features for its tiles, then refills the stage it
//                just drained (it is the stage's only reader at that point: no empty barriers).  The four
//                partial sums are combined through two staging buffers at the end of the pass.
//
// Slot update (per group, all 8 warps; six exchange rounds per iteration)
//   The cluster exchanges data with st.async (DSMEM store + complete_tx on the receiver's mbarrier); the
//   receive buffers and barriers are double-buffered by round parity, a peer can be at most one round
//   ahead.  R1 reduce-scatters the partial sum_n w v / sum_n w; R2..R6 all-gather the CTA's slice of
//   updates, h', the MLP hidden layer, the new slots and q.  The matrix-vector products run on tensor
//   cores (weights = A operand from shared memory, activations as a bf16 hi/lo pair = B operand).
#include "pc_common.cuh"

namespace ocrl {

namespace pc {

template <int D_, int H_, int CL_, int NG_, int S_, int SUB_>
struct Cfg {
  static constexpr int D = D_, H = H_, CL = CL_, NG = NG_, S = S_, SUB = SUB_;
  static constexpr int KP = 8;                   // slots padded to one MMA n-tile
  static constexpr int GW = 8, GT = 256;         // warps / threads per group
  static constexpr int NT = NG * GT;
  static constexpr int TOK = 16 * SUB;           // tokens per tile (one TMA box per operand), SUB MMA row blocks
  static constexpr int PITCH = D * 2 + 16;       // bytes per padded bf16 row of length D
  static constexpr int PITCHH = H * 2 + 16;      // ... of length H
  static constexpr int LX = D > H ? D : H;
  static constexpr int PITCHX = LX * 2 + 16;
  static constexpr int TILE_BYTES = TOK * D * 2;           // dense rows, 128-byte swizzle applied by the TMA unit
  static constexpr int STAGE_BYTES = 2 * TILE_BYTES;       // k tile, v tile (multiple of 1024)
  static constexpr int WT_BYTES = SUB * 256;               // softmax weights of a tile in B-fragment order
  static constexpr int DS = D / CL, HS = H / CL; // output features of the GRU/q/MLP2 and of MLP1 owned by a CTA
  static constexpr int NMU = D / 16;             // 16-feature m-tiles of U^T (every U warp covers all of D)
  static constexpr int NMG = (3 * DS + 15) / 16; // m-tiles of one GRU weight slice
  static constexpr int NM1 = (HS + 15) / 16, NM2 = (DS + 15) / 16;
  static constexpr int NKC = 4;                  // k-chunks of the MLP / q products (8 warps share 1-2 m-tiles)
  static constexpr int UP = D + 4;               // float pitch of the U staging rows (conflict-free fragment stores)
  static_assert(D % 64 == 0 && H % 64 == 0, "D, H multiples of 64");
  static_assert((TILE_BYTES % 1024) == 0, "swizzle atoms are 1024 bytes");
  static_assert(DS % 4 == 0 && HS % 4 == 0, "slices must be float4 multiples");

  // ---- CTA-level shared memory (bytes)
  static constexpr int OFF_WIH = 0;
  static constexpr int OFF_WHH = OFF_WIH + 3 * DS * PITCH;
  static constexpr int OFF_W1 = OFF_WHH + 3 * DS * PITCH;
  static constexpr int OFF_W2 = OFF_W1 + HS * PITCH;
  static constexpr int OFF_WQ = OFF_W2 + DS * PITCHH;
  static constexpr int OFF_ZROW = OFF_WQ + DS * PITCH;
  static constexpr int OFF_LNP = OFF_ZROW + PITCHX;                // ln_slots w,b, ln_mlp w,b  [4][D] fp32
  static constexpr int OFF_BIAS = OFF_LNP + 4 * D * 4;             // b_ih[3DS] b_hh[3DS] b1[HS] b2[DS] fp32
  static constexpr int CTA_BYTES = (OFF_BIAS + (7 * DS + HS) * 4 + 1023) & ~1023;
  // ---- per-group shared memory (bytes)
  static constexpr int XBUF_BYTES = KP * LX * 4 + 32 * CL;         // receive buffer of one exchange round
  static constexpr int G_RING = 0;
  static constexpr int G_WT = G_RING + S * STAGE_BYTES;
  static constexpr int G_XBUF = G_WT + S * WT_BYTES;
  static constexpr int G_SLH = G_XBUF + 2 * XBUF_BYTES;            // slots entering the iteration, bf16 hi / lo
  static constexpr int G_ACT = G_SLH + 2 * KP * PITCH;             // activation staging hi / lo (aliases U staging)
  static constexpr int ACT_BYTES = (2 * KP * PITCHX > KP * UP * 4) ? 2 * KP * PITCHX : KP * UP * 4;
  static constexpr int G_P = G_ACT + ACT_BYTES;                    // MMA partial outputs [rows][8] fp32
  static constexpr int P_ROWS = (2 * NMG > NKC * (NM1 > NM2 ? NM1 : NM2)) ? 2 * NMG * 16 : NKC * (NM1 > NM2 ? NM1 : NM2) * 16;
  static constexpr int P_BYTES = (P_ROWS * 32 > KP * UP * 4) ? P_ROWS * 32 : KP * UP * 4;  // also the 2nd U staging buffer
  static constexpr int G_OWN = G_P + P_BYTES;                      // own slice of h / slots [KP][DS] fp32
  static constexpr int G_SRED = G_OWN + KP * DS * 4;               // [4][8] fp32
  static constexpr int G_BAR = G_SRED + 128;                       // 2S + 2 mbarriers
  static constexpr int GROUP_BYTES = (G_BAR + (2 * S + 2) * 8 + 1023) & ~1023;
  static_assert(S % 4 == 0, "a warp must stay on its own stages (mbarrier parity waits alias two phases ahead)");
  static constexpr int SMEM_BYTES = CTA_BYTES + NG * GROUP_BYTES + 1024;
};

template <int D, int H, int CL, int NG, int S, int SUB>
__global__ void __launch_bounds__(NG * 256, 1)
sa_iter_fwd_pc_kernel(const IterFwdArgs a, const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v) {
  using C = Cfg<D, H, CL, NG, S, SUB>;
  constexpr int KP = C::KP, PITCH = C::PITCH, PITCHH = C::PITCHH, PITCHX = C::PITCHX, DS = C::DS, HS = C::HS;
  constexpr int NMU = C::NMU, NMG = C::NMG, NM1 = C::NM1, NM2 = C::NM2, NKC = C::NKC, UP = C::UP, TOK = C::TOK;
  constexpr float LOG2E = 1.4426950408889634f;

  extern __shared__ unsigned char smem_raw[];
  // align by pointer arithmetic on the __shared__ pointer (an integer round trip would turn every access generic)
  unsigned char* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / CL, ncl = gridDim.x / CL;
  const int tid = threadIdx.x, grp = tid >> 8, gtid = tid & 255, warp = gtid >> 5, lane = tid & 31;
  const int g8 = lane >> 2, t4 = lane & 3;
  const int K = a.K, N = a.N, T = a.T, B = a.B;
  const bool tracer = (a.trace != nullptr && blockIdx.x == 0 && tid == 0);
#define PC_TRACE(i) do { if (tracer) a.trace[(i)] = clock64(); } while (0)
  PC_TRACE(0);

  unsigned char* s_wih = sm + C::OFF_WIH;
  unsigned char* s_whh = sm + C::OFF_WHH;
  unsigned char* s_w1 = sm + C::OFF_W1;
  unsigned char* s_w2 = sm + C::OFF_W2;
  unsigned char* s_wq = sm + C::OFF_WQ;
  unsigned char* s_zrow = sm + C::OFF_ZROW;
  float* s_lnp = reinterpret_cast<float*>(sm + C::OFF_LNP);
  float* s_bias = reinterpret_cast<float*>(sm + C::OFF_BIAS);
  const float* s_bih = s_bias;
  const float* s_bhh = s_bias + 3 * DS;
  const float* s_b1 = s_bias + 6 * DS;
  const float* s_b2 = s_bias + 6 * DS + HS;

  unsigned char* gb = sm + C::CTA_BYTES + grp * C::GROUP_BYTES;
  unsigned char* ring = gb + C::G_RING;
  unsigned char* wtiles = gb + C::G_WT;
  unsigned char* xbuf = gb + C::G_XBUF;
  unsigned char* slh_hi = gb + C::G_SLH;
  unsigned char* slh_lo = slh_hi + KP * PITCH;
  unsigned char* act_hi = gb + C::G_ACT;
  unsigned char* act_lo = act_hi + KP * PITCHX;
  float* ustage_a = reinterpret_cast<float*>(gb + C::G_ACT);  // the two U staging buffers alias the activation
  float* ustage_b = reinterpret_cast<float*>(gb + C::G_P);    // staging and the MMA partials (idle during a pass)
  float* P = reinterpret_cast<float*>(gb + C::G_P);
  float* own = reinterpret_cast<float*>(gb + C::G_OWN);
  float* sred = reinterpret_cast<float*>(gb + C::G_SRED);
  uint64_t* bars = reinterpret_cast<uint64_t*>(gb + C::G_BAR);
  uint64_t* full = bars;           // k and v of the stage have landed (tx bytes)
  uint64_t* w_ready = bars + S;    // the logit warp has written the tile's softmax weights (and is done with k)
  uint64_t* xbar = bars + 2 * S;   // exchange rounds, by parity

  // ------------------------------------------------------------------ one-time setup
  {
    // the CTA's weight slices, fp32 global -> bf16 shared; four independent 16-byte loads in flight per thread
    auto load_rows = [&](unsigned char* dst, int pitch, const float* src, int L, int nrows) {
      const int total = nrows * (L / 4);
      for (int i0 = tid; i0 < total; i0 += 4 * C::NT) {
        float4 x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * C::NT;
          if (i < total) x[u] = __ldg(reinterpret_cast<const float4*>(src) + i);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * C::NT;
          if (i < total) {
            const int r = i / (L / 4), c4 = i % (L / 4);
            *reinterpret_cast<uint2*>(dst + r * pitch + 8 * c4) =
                make_uint2(pack_bf16x2(x[u].x, x[u].y), pack_bf16x2(x[u].z, x[u].w));
          }
        }
      }
    };
    for (int gate = 0; gate < 3; ++gate) {
      load_rows(s_wih + gate * DS * PITCH, PITCH, a.w.w_ih + ((size_t)gate * D + rank * DS) * D, D, DS);
      load_rows(s_whh + gate * DS * PITCH, PITCH, a.w.w_hh + ((size_t)gate * D + rank * DS) * D, D, DS);
    }
    load_rows(s_w1, PITCH, a.w.w1 + (size_t)rank * HS * D, D, HS);
    load_rows(s_w2, PITCHH, a.w.w2 + (size_t)rank * DS * H, H, DS);
    load_rows(s_wq, PITCH, a.w.wq + (size_t)rank * DS * D, D, DS);
    for (int i = tid; i < PITCHX / 4; i += C::NT) reinterpret_cast<uint32_t*>(s_zrow)[i] = 0u;
    for (int i = tid; i < D; i += C::NT) {
      s_lnp[i] = a.w.ln_slots_w[i];
      s_lnp[D + i] = a.w.ln_slots_b[i];
      s_lnp[2 * D + i] = a.w.ln_mlp_w[i];
      s_lnp[3 * D + i] = a.w.ln_mlp_b[i];
    }
    for (int i = tid; i < 3 * DS; i += C::NT) {
      const int gate = i / DS, dl = i % DS;
      s_bias[i] = a.w.b_ih[gate * D + rank * DS + dl];
      s_bias[3 * DS + i] = a.w.b_hh[gate * D + rank * DS + dl];
    }
    for (int i = tid; i < HS; i += C::NT) s_bias[6 * DS + i] = a.w.b1[rank * HS + i];
    for (int i = tid; i < DS; i += C::NT) s_bias[6 * DS + HS + i] = a.w.b2[rank * DS + i];
    if (gtid == 0) {
      for (int s = 0; s < S; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&w_ready[s], 1);
      }
      mbar_init(&xbar[0], 1);
      mbar_init(&xbar[1], 1);
      mbar_fence_init();
    }
    // staging rows of the padded slots are never written by the conversions; keep them finite
    for (int i = gtid; i < (2 * KP * PITCH + C::ACT_BYTES) / 4; i += 256) reinterpret_cast<uint32_t*>(slh_hi)[i] = 0u;
  }
  __syncthreads();
  cluster.sync();  // every CTA's barriers are initialised before any peer signals them
  PC_TRACE(1);

  // ------------------------------------------------------------------ work assignment
  const int G = ncl * NG, gi = cid * NG + grp;
  const int n_img = (B > gi) ? (B - gi + G - 1) / G : 0;
  const int ntiles = (N + TOK - 1) / TOK;
  const int TPC = (ntiles + CL - 1) / CL;
  const int tile0 = rank * TPC;
  const int TP = max(0, min(TPC, ntiles - tile0));  // tiles of this CTA per pass
  const long long total_seq = (long long)n_img * T * TP;
  auto image_of = [&](int m) { return B - 1 - (gi + m * G); };  // newest first: still in L2 after the projection

  // one elected lane: stage (j % S) <- tile j of the group's tile sequence.  Rows past the image's N tokens are
  // the next image's (finite) or out-of-bounds zeros; their softmax weight is forced to 0.
  auto issue_tile = [&](long long j) {
    const int s = (int)(j % S);
    const int m = (int)(j / ((long long)T * TP));
    const int tile = (int)(j % TP);
    const int row0 = image_of(m) * N + (tile0 + tile) * TOK;
    unsigned char* kd = ring + (size_t)s * C::STAGE_BYTES;
    mbar_expect_tx(&full[s], (uint32_t)C::STAGE_BYTES);
    tma_load_3d(kd, &tm_k, 0, 0, row0, &full[s]);
    tma_load_3d(kd + C::TILE_BYTES, &tm_v, 0, 0, row0, &full[s]);
  };

  // ---- exchange helpers (group scope).  Round n uses receive buffer / barrier n & 1.
  uint32_t round = 0;
  auto xb = [&](uint32_t n) { return xbuf + (n & 1) * C::XBUF_BYTES; };
  auto arm = [&](uint32_t n, uint32_t bytes) {
    if (gtid == 0) mbar_expect_tx(&xbar[n & 1], bytes);
  };
  auto xwait = [&](uint32_t n) { mbar_wait_cluster(&xbar[n & 1], (n >> 1) & 1); };
  // All-gather push without staging: thread i < K*SL holds element (slot = i / SL, l = i % SL) of the CTA's
  // slice.  The four lanes of a quad assemble the float4 of features 4*(l/4).. and each sends it to CL/4 CTAs.
  // Must be called by whole warps (i may exceed K*SL; those lanes only take part in the shuffles).
  auto quad_push = [&](uint32_t n, float val, int i, int SL, int pitchf) {
    const int qb = lane & ~3;
    float4 v4;
    v4.x = __shfl_sync(FULL, val, qb);
    v4.y = __shfl_sync(FULL, val, qb + 1);
    v4.z = __shfl_sync(FULL, val, qb + 2);
    v4.w = __shfl_sync(FULL, val, qb + 3);
    if (i < K * SL) {
      const int slot = i / SL, l4 = (i % SL) & ~3;
      const uint32_t lbuf = smem_u32(xb(n)) + (uint32_t)((slot * pitchf + rank * SL + l4) * 4);
      const uint32_t lbar = smem_u32(&xbar[n & 1]);
#pragma unroll
      for (int r = 0; r < CL / 4; ++r) {
        const int dest = (lane & 3) + 4 * r;
        st_async_v4(mapa_u32(lbuf, dest), v4, mapa_u32(lbar, dest));
      }
    }
  };
  // fp32 [K][L] (pitch L floats, shared) -> bf16 hi / lo staging (pitch `ap` bytes)
  auto to_hilo = [&](const float* src, int L, unsigned char* hi, unsigned char* lo, int ap) {
    for (int i = gtid; i < K * (L / 2); i += 256) {
      const int slot = i / (L / 2), c2 = i % (L / 2);
      const float2 x = *reinterpret_cast<const float2*>(src + slot * L + 2 * c2);
      uint32_t h, l;
      split_hilo(x.x, x.y, h, l);
      *reinterpret_cast<uint32_t*>(hi + slot * ap + 4 * c2) = h;
      *reinterpret_cast<uint32_t*>(lo + slot * ap + 4 * c2) = l;
    }
  };
  // LayerNorm of rows [K][D] (generic pointer, pitch D) -> hi/lo staging; optionally also the raw rows -> raw hi/lo
  auto ln_to_hilo = [&](const float* src, const float* gw, const float* gbias, unsigned char* hi, unsigned char* lo,
                        unsigned char* raw_hi, unsigned char* raw_lo) {
    constexpr int NCH = D / 64;
    if (warp < K) {
      float2 x[NCH];
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        x[c] = *reinterpret_cast<const float2*>(src + warp * D + 64 * c + 2 * lane);
        s += x[c].x + x[c].y;
        if (raw_hi != nullptr) {
          uint32_t h, l;
          split_hilo(x[c].x, x[c].y, h, l);
          *reinterpret_cast<uint32_t*>(raw_hi + warp * PITCH + (64 * c + 2 * lane) * 2) = h;
          *reinterpret_cast<uint32_t*>(raw_lo + warp * PITCH + (64 * c + 2 * lane) * 2) = l;
        }
      }
      const float mean = warp_sum(s) * (1.f / D);
      float q = 0.f;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        x[c].x -= mean;
        x[c].y -= mean;
        q = fmaf(x[c].x, x[c].x, fmaf(x[c].y, x[c].y, q));
      }
      const float rstd = rsqrtf(warp_sum(q) * (1.f / D) + a.ln_eps);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const float2 g = *reinterpret_cast<const float2*>(gw + 64 * c + 2 * lane);
        const float2 b = *reinterpret_cast<const float2*>(gbias + 64 * c + 2 * lane);
        uint32_t h, l;
        split_hilo(x[c].x * rstd * g.x + b.x, x[c].y * rstd * g.y + b.y, h, l);
        *reinterpret_cast<uint32_t*>(hi + warp * PITCHX + (64 * c + 2 * lane) * 2) = h;
        *reinterpret_cast<uint32_t*>(lo + warp * PITCHX + (64 * c + 2 * lane) * 2) = l;
      }
    }
  };

  // q = W_q . LN(slots) for the CTA's slice, all-gathered as bf16 (pre-multiplied by log2 e): round `round`
  auto q_phase = [&](const float* slots_full /* [K][D], shared or global */) {
    arm(round, (uint32_t)(K * D * 2));
    ln_to_hilo(slots_full, s_lnp, s_lnp + D, act_hi, act_lo, slh_hi, slh_lo);
    group_sync(grp);
    for (int job = warp; job < NM2 * NKC; job += C::GW) {
      const int mt = job / NKC, kc = job % NKC;
      mma_job(s_wq, PITCH, DS, s_zrow, mt, kc * (D / 16) / NKC, (kc + 1) * (D / 16) / NKC, act_hi, act_lo, PITCHX,
              P + kc * NM2 * 128, lane);
    }
    group_sync(grp);
    {  // bf16 q slice: a quad assembles 4 features (8 bytes) and each of its lanes serves CL/4 destinations
      const int nel = K * DS;
      for (int i0 = warp * 32; i0 < nel; i0 += 256) {
        const int i = i0 + lane;
        float val = 0.f;
        if (i < nel) {
          const int slot = i / DS, dl = i % DS;
#pragma unroll
          for (int kc = 0; kc < NKC; ++kc) val += P[(kc * NM2 * 16 + dl) * 8 + slot];
          val *= LOG2E;
        }
        const int qb = lane & ~3;
        const float v0 = __shfl_sync(FULL, val, qb), v1 = __shfl_sync(FULL, val, qb + 1);
        const float v2 = __shfl_sync(FULL, val, qb + 2), v3 = __shfl_sync(FULL, val, qb + 3);
        if (i < nel) {
          const int slot = i / DS, l4 = (i % DS) & ~3;
          const uint32_t lbuf = smem_u32(xb(round)) + (uint32_t)(slot * PITCH + (rank * DS + l4) * 2);
          const uint32_t lbar = smem_u32(&xbar[round & 1]);
          const uint32_t lo = pack_bf16x2(v0, v1), hi = pack_bf16x2(v2, v3);
#pragma unroll
          for (int r = 0; r < CL / 4; ++r) {
            const int dest = (lane & 3) + 4 * r;
            st_async_v2(mapa_u32(lbuf, dest), lo, hi, mapa_u32(lbar, dest));
          }
        }
      }
    }
    xwait(round);
    ++round;
  };

  // ------------------------------------------------------------------ main loop over the group's images
  if (n_img > 0) {
    if (warp >= 4 && lane == 0)  // U warp u owns the stages s = u (mod 4)
      for (int p = warp - 4; p < S; p += 4)
        if (p < total_seq) issue_tile(p);

    for (int m = 0; m < n_img; ++m) {
      const int img = image_of(m);
      // own slice of the initial slots (fp32 state of the GRU blend / residual)
      for (int i = gtid; i < K * DS; i += 256) {
        const int slot = i / DS, dl = i % DS;
        own[slot * DS + dl] = a.slots0[((size_t)img * K + slot) * D + rank * DS + dl];
      }
      q_phase(a.slots0 + (size_t)img * K * D);

      for (int t = 0; t < T; ++t) {
        const bool last = (t == T - 1);
        const int tb = 2 + (m * T + t) * 12;  // trace slots of this image-iteration (first images only)
        const bool tr_on = tracer && tb + 12 < 500;
#define PC_T(i) do { if (tr_on) a.trace[tb + (i)] = clock64(); } while (0)
        PC_T(0);
        const long long jbase = ((long long)m * T + t) * TP;
        const unsigned char* qsrc = xb(round - 1);  // bf16 q [8][PITCH] of the round that just completed
        // ============================================================ token pass
        if (warp < 4) {
          // ---- logit warp
          uint32_t qb[D / 16][2];
#pragma unroll
          for (int ks = 0; ks < D / 16; ++ks) {
            qb[ks][0] = *reinterpret_cast<const uint32_t*>(qsrc + g8 * PITCH + ks * 32 + 4 * t4);
            qb[ks][1] = *reinterpret_cast<const uint32_t*>(qsrc + g8 * PITCH + ks * 32 + 16 + 4 * t4);
          }
          const int c0 = 2 * t4, c1 = 2 * t4 + 1;
          const bool ok0 = c0 < K, ok1 = c1 < K;
          float Sl0 = 0.f, Sl1 = 0.f;
          const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lhalf = (lane >> 4) * 16;
          for (int tile = warp; tile < TP; tile += 4) {
            const long long j = jbase + tile;
            const int s = (int)(j % S);
            const uint32_t ph = (uint32_t)((j / S) & 1);
            const unsigned char* kt = ring + (size_t)s * C::STAGE_BYTES;
            uint32_t* wt = reinterpret_cast<uint32_t*>(wtiles + s * C::WT_BYTES);
            mbar_wait(&full[s], ph);
#pragma unroll
            for (int sub = 0; sub < SUB; ++sub) {
              float ca[4] = {0.f, 0.f, 0.f, 0.f}, cb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
              for (int ks = 0; ks < D / 16; ++ks) {
                uint32_t kf[4];
                ldmatrix_x4(kf, kt + swz_off<D>(16 * sub + lrow, ks * 32 + lhalf));
                if (ks & 1) mma_bf16_16816(cb, kf, qb[ks][0], qb[ks][1]);
                else mma_bf16_16816(ca, kf, qb[ks][0], qb[ks][1]);
              }
              // softmax over the slot axis: a token's 8 logits live in the 4 lanes of a quad (2 each)
              const int tok = (tile0 + tile) * TOK + 16 * sub + g8;
              float w[4];
#pragma unroll
              for (int hrow = 0; hrow < 2; ++hrow) {
                float x0 = ok0 ? ca[2 * hrow] + cb[2 * hrow] : -INFINITY;
                float x1 = ok1 ? ca[2 * hrow + 1] + cb[2 * hrow + 1] : -INFINITY;
                float mx = fmaxf(x0, x1);
                mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, 1));
                mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, 2));
                const float e0 = ex2f(x0 - mx), e1 = ex2f(x1 - mx);
                float sum = e0 + e1;
                sum += __shfl_xor_sync(FULL, sum, 1);
                sum += __shfl_xor_sync(FULL, sum, 2);
                const float inv = __fdividef(1.f, sum);
                const float a0 = e0 * inv, a1 = e1 * inv;
                const int tk = tok + 8 * hrow;
                const bool tok_ok = tk < N;
                if (last && a.attn_out != nullptr && tok_ok) {
                  float* ao = a.attn_out + ((size_t)img * N + tk) * K;
                  if ((K & 1) == 0) {
                    if (ok0) *reinterpret_cast<float2*>(ao + c0) = make_float2(a0, a1);
                  } else {
                    if (ok0) ao[c0] = a0;
                    if (ok1) ao[c1] = a1;
                  }
                }
                w[2 * hrow] = (tok_ok && ok0) ? a0 + a.eps : 0.f;
                w[2 * hrow + 1] = (tok_ok && ok1) ? a1 + a.eps : 0.f;
              }
              // weights rounded to bf16 once; the same rounded values feed the numerator and the token sum
              const __nv_bfloat162 p0 = __floats2bfloat162_rn(w[0], w[1]);
              const __nv_bfloat162 p1 = __floats2bfloat162_rn(w[2], w[3]);
              Sl0 += __low2float(p0) + __low2float(p1);
              Sl1 += __high2float(p0) + __high2float(p1);
              wt[64 * sub + lane] = movmatrix_trans(*reinterpret_cast<const uint32_t*>(&p0));
              wt[64 * sub + 32 + lane] = movmatrix_trans(*reinterpret_cast<const uint32_t*>(&p1));
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&w_ready[s]);
          }
#pragma unroll
          for (int o = 4; o < 32; o <<= 1) {
            Sl0 += __shfl_xor_sync(FULL, Sl0, o);
            Sl1 += __shfl_xor_sync(FULL, Sl1, o);
          }
          if (g8 == 0) {
            sred[warp * 8 + c0] = Sl0;
            sred[warp * 8 + c1] = Sl1;
          }
          PC_T(1);
        } else {
          // ---- U warp: all D features of its tiles
          const int uw = warp - 4;
          float acc[NMU][4];
#pragma unroll
          for (int i = 0; i < NMU; ++i)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[i][e] = 0.f;
          const int urow = (lane & 7) + (lane >> 4) * 8, uhalf = ((lane >> 3) & 1) * 16;
          for (int tile = uw; tile < TP; tile += 4) {
            const long long j = jbase + tile;
            const int s = (int)(j % S);
            const uint32_t ph = (uint32_t)((j / S) & 1);
            const unsigned char* vt = ring + (size_t)s * C::STAGE_BYTES + C::TILE_BYTES;
            const uint32_t* wt = reinterpret_cast<const uint32_t*>(wtiles + s * C::WT_BYTES);
            mbar_wait(&w_ready[s], ph);  // implies full[s]: the logit warp waited for k and v together
#pragma unroll
            for (int sub = 0; sub < SUB; ++sub) {
              const uint32_t b0 = wt[64 * sub + lane], b1 = wt[64 * sub + 32 + lane];
#pragma unroll
              for (int i = 0; i < NMU; ++i) {
                uint32_t vf[4];
                // [tok 0-7, d 0-7], [tok 0-7, d 8-15], [tok 8-15, d 0-7], [tok 8-15, d 8-15] transposed
                //  = A fragments a0 (d 0-7, tok 0-7), a1 (d 8-15, tok 0-7), a2 (d 0-7, tok 8-15), a3 (d 8-15, tok 8-15)
                ldmatrix_x4_trans(vf, vt + swz_off<D>(16 * sub + urow, uhalf + i * 32));
                mma_bf16_16816(acc[i], vf, b0, b1);
              }
            }
            __syncwarp();  // every lane is done with the stage (this warp is its only remaining reader)
            if (lane == 0 && j + S < total_seq) issue_tile(j + S);
          }
          // combine the four partial sums: warps 4, 5 write the two staging buffers, warps 6, 7 add into them
          float* ust = (uw & 1) ? ustage_b : ustage_a;
          if (uw >= 2) asm volatile("bar.sync %0, 128;" ::"r"(3 + grp) : "memory");
#pragma unroll
          for (int i = 0; i < NMU; ++i) {
            const int d0 = 16 * i + g8;
            float* u0 = ust + (2 * t4) * UP + d0;
            float* u1 = ust + (2 * t4 + 1) * UP + d0;
            if (uw >= 2) {
              u0[0] += acc[i][0]; u1[0] += acc[i][1]; u0[8] += acc[i][2]; u1[8] += acc[i][3];
            } else {
              u0[0] = acc[i][0]; u1[0] = acc[i][1]; u0[8] = acc[i][2]; u1[8] = acc[i][3];
            }
          }
          if (uw < 2) asm volatile("bar.sync %0, 128;" ::"r"(3 + grp) : "memory");
        }
        // ============================================================ R1: reduce-scatter of sum w v, all-reduce of sum w
        arm(round, (uint32_t)(K * D * 4 + 32 * CL));
        group_sync(grp);
        PC_T(2);
        {
          const uint32_t lbuf = smem_u32(xb(round)), lbar = smem_u32(&xbar[round & 1]);
          constexpr int QPR = D / 4;  // float4 chunks per slot row
          for (int i = gtid; i < K * QPR; i += 256) {
            const int slot = i / QPR, d = 4 * (i % QPR);
            const int dest = d / DS, dl = d % DS;
            const float4 xa = *reinterpret_cast<const float4*>(ustage_a + slot * UP + d);
            const float4 xb4 = *reinterpret_cast<const float4*>(ustage_b + slot * UP + d);
            const float4 val = make_float4(xa.x + xb4.x, xa.y + xb4.y, xa.z + xb4.z, xa.w + xb4.w);
            const uint32_t off = (uint32_t)(((rank * KP + slot) * DS + dl) * 4);
            st_async_v4(mapa_u32(lbuf + off, dest), val, mapa_u32(lbar, dest));
          }
          if (gtid < 8 * CL) {
            const int slot = gtid & 7, dest = gtid >> 3;
            const float s = sred[slot] + sred[8 + slot] + sred[16 + slot] + sred[24 + slot];
            const uint32_t off = (uint32_t)(KP * D * 4 + (rank * 8 + slot) * 4);
            st_async_b32(mapa_u32(lbuf + off, dest), __float_as_uint(s), mapa_u32(lbar, dest));
          }
        }
        PC_T(3);
        xwait(round);
        PC_T(4);
        // ============================================================ R2: all-gather updates = sum over CTAs / token sum
        arm(round + 1, (uint32_t)(K * D * 4));
        {
          const float* rs = reinterpret_cast<const float*>(xb(round));
          const float* ss = rs + KP * D;
          for (int i0 = warp * 32; i0 < K * DS; i0 += 256) {
            const int i = i0 + lane;
            float val = 0.f;
            if (i < K * DS) {
              const int slot = i / DS, dl = i % DS;
              float u = 0.f, s = 0.f;
#pragma unroll
              for (int src = 0; src < CL; ++src) {
                u += rs[(src * KP + slot) * DS + dl];
                s += ss[src * 8 + slot];
              }
              val = u / s;
            }
            quad_push(round + 1, val, i, DS, D);
          }
        }
        ++round;
        xwait(round);
        PC_T(5);
        to_hilo(reinterpret_cast<const float*>(xb(round)), D, act_hi, act_lo, PITCHX);
        ++round;
        group_sync(grp);
        // ---- GRU: gi = W_ih u, gh = W_hh h
        for (int job = warp; job < 2 * NMG; job += C::GW) {
          const bool hh = job >= NMG;
          const int mt = hh ? job - NMG : job;
          mma_job(hh ? s_whh : s_wih, PITCH, 3 * DS, s_zrow, mt, 0, D / 16, hh ? slh_hi : act_hi, hh ? slh_lo : act_lo,
                  hh ? PITCH : PITCHX, P + (hh ? NMG * 128 : 0), lane);
        }
        arm(round, (uint32_t)(K * D * 4));
        group_sync(grp);
        PC_T(6);
        // ============================================================ R3: all-gather h'
        for (int i0 = warp * 32; i0 < K * DS; i0 += 256) {
          const int i = i0 + lane;
          float hp = 0.f;
          if (i < K * DS) {
            const int slot = i / DS, dl = i % DS;
            const float* Pi = P;
            const float* Ph = P + NMG * 128;
            const float gir = Pi[(dl) * 8 + slot] + s_bih[dl], ghr = Ph[(dl) * 8 + slot] + s_bhh[dl];
            const float giz = Pi[(DS + dl) * 8 + slot] + s_bih[DS + dl], ghz = Ph[(DS + dl) * 8 + slot] + s_bhh[DS + dl];
            const float gin = Pi[(2 * DS + dl) * 8 + slot] + s_bih[2 * DS + dl];
            const float ghn = Ph[(2 * DS + dl) * 8 + slot] + s_bhh[2 * DS + dl];
            const float r = sigmoidf_(gir + ghr), z = sigmoidf_(giz + ghz);
            const float n = tanhf(gin + r * ghn);
            hp = (1.f - z) * n + z * own[i];
            own[i] = hp;
          }
          quad_push(round, hp, i, DS, D);
        }
        xwait(round);
        PC_T(7);
        ln_to_hilo(reinterpret_cast<const float*>(xb(round)), s_lnp + 2 * D, s_lnp + 3 * D, act_hi, act_lo, nullptr, nullptr);
        ++round;
        group_sync(grp);
        for (int job = warp; job < NM1 * NKC; job += C::GW) {
          const int mt = job / NKC, kc = job % NKC;
          mma_job(s_w1, PITCH, HS, s_zrow, mt, kc * (D / 16) / NKC, (kc + 1) * (D / 16) / NKC, act_hi, act_lo, PITCHX,
                  P + kc * NM1 * 128, lane);
        }
        arm(round, (uint32_t)(K * H * 4));
        group_sync(grp);
        PC_T(8);
        // ============================================================ R4: all-gather the MLP hidden layer
        for (int i0 = warp * 32; i0 < K * HS; i0 += 256) {
          const int i = i0 + lane;
          float hid = 0.f;
          if (i < K * HS) {
            const int slot = i / HS, hl = i % HS;
            float sum = s_b1[hl];
#pragma unroll
            for (int kc = 0; kc < NKC; ++kc) sum += P[(kc * NM1 * 16 + hl) * 8 + slot];
            hid = fmaxf(sum, 0.f);
          }
          quad_push(round, hid, i, HS, H);
        }
        xwait(round);
        PC_T(9);
        to_hilo(reinterpret_cast<const float*>(xb(round)), H, act_hi, act_lo, PITCHX);
        ++round;
        group_sync(grp);
        for (int job = warp; job < NM2 * NKC; job += C::GW) {
          const int mt = job / NKC, kc = job % NKC;
          mma_job(s_w2, PITCHH, DS, s_zrow, mt, kc * (H / 16) / NKC, (kc + 1) * (H / 16) / NKC, act_hi, act_lo, PITCHX,
                  P + kc * NM2 * 128, lane);
        }
        if (!last) arm(round, (uint32_t)(K * D * 4));
        group_sync(grp);
        // ============================================================ R5: all-gather the new slots (or write them out)
        for (int i0 = warp * 32; i0 < K * DS; i0 += 256) {
          const int i = i0 + lane;
          float sn = 0.f;
          if (i < K * DS) {
            const int slot = i / DS, dl = i % DS;
            float sum = s_b2[dl];
#pragma unroll
            for (int kc = 0; kc < NKC; ++kc) sum += P[(kc * NM2 * 16 + dl) * 8 + slot];
            sn = own[i] + sum;
            own[i] = sn;
            if (last) a.slots_out[((size_t)img * K + slot) * D + rank * DS + dl] = sn;
          }
          if (!last) quad_push(round, sn, i, DS, D);
        }
        PC_T(10);
        if (!last) {
          xwait(round);
          const float* sfull = reinterpret_cast<const float*>(xb(round));
          ++round;
          PC_T(11);
          q_phase(sfull);  // R6
        } else {
          group_sync(grp);  // `own`, `P` are rewritten by the next image
        }
      }
    }
  }
  __syncthreads();
  cluster.sync();  // no CTA leaves while a peer may still address its shared memory
}

template <int D, int H, int CL, int NG, int S, int SUB>
static int launch_pc(const IterFwdArgs& a, cudaStream_t stream) {
  using C = Cfg<D, H, CL, NG, S, SUB>;
  auto kern = sa_iter_fwd_pc_kernel<D, H, CL, NG, S, SUB>;
  CUtensorMap tm_k, tm_v;
  if (!make_kv_map(&tm_k, a.k, (long long)a.B * a.N, D, C::TOK) || !make_kv_map(&tm_v, a.v, (long long)a.B * a.N, D, C::TOK)) {
    set_error("sa_iter_fwd(persistent): cuTensorMapEncodeTiled failed");
    return OCRL_E_LAUNCH;
  }
  static_assert(C::SMEM_BYTES <= 227 * 1024, "shared memory budget");
  OCRL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  if (CL > 8) OCRL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(C::NT);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // resident clusters: the hardware packs clusters into GPCs, ask it
  static int max_clusters = -1;  // per (template instance, process); all devices of a box are identical
  if (max_clusters < 0) {
    cfg.gridDim = dim3(CL * 148);
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
    if (e != cudaSuccess || n <= 0) {
      (void)cudaGetLastError();
      set_error("sa_iter_fwd(persistent): cluster size %d with %d B of shared memory cannot be scheduled", CL, C::SMEM_BYTES);
      return OCRL_E_SHAPE;
    }
    max_clusters = n;
    if (getenv("OCRL_SA_PC_VERBOSE")) fprintf(stderr, "[ocrl] pc kernel CL=%d NG=%d S=%d SUB=%d smem=%d B: max active clusters %d\n", CL, NG, S, SUB, C::SMEM_BYTES, n);
  }
  int ncl = max_clusters;
  if (const char* e = getenv("OCRL_SA_PC_CLUSTERS")) ncl = max(1, min(ncl, atoi(e)));
  const int want = (a.B + NG - 1) / NG;
  if (ncl > want) ncl = want;
  // equalise the images per group: with r rounds, use the fewest clusters that still need r rounds
  const int rounds = (want + ncl - 1) / ncl;
  ncl = (want + rounds - 1) / rounds;
  cfg.gridDim = dim3((unsigned)(ncl * CL));
  OCRL_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, a, tm_k, tm_v));
  return OCRL_OK;
}

}  // namespace pc

// Returns OCRL_E_SHAPE (without launching) for shapes this kernel does not cover; the caller falls back.
int sa_iter_fwd_pc_dispatch(const IterFwdArgs& a, cudaStream_t s) {
  if (a.K > 8 || a.saved != nullptr) {
    set_error("sa_iter_fwd(persistent): K <= 8, inference only");
    return OCRL_E_SHAPE;
  }
  int variant = 0;
  if (const char* e = getenv("OCRL_SA_PC")) variant = atoi(e);
  if (a.D == 192 && a.H == 192) {
    switch (variant) {
      case 1: return pc::launch_pc<192, 192, 16, 1, 8, 1>(a, s);
      case 2: return pc::launch_pc<192, 192, 16, 2, 4, 1>(a, s);
      case 3: return pc::launch_pc<192, 192, 16, 1, 4, 2>(a, s);
      case 4: return pc::launch_pc<192, 192, 8, 1, 4, 2>(a, s);
      default: return pc::launch_pc<192, 192, 8, 1, 8, 1>(a, s);
    }
  }
  set_error("sa_iter_fwd(persistent): D=%d H=%d not instantiated", a.D, a.H);
  return OCRL_E_SHAPE;
}

}  // namespace ocrl

// Two-engine persistent-cluster form of the fused T-iteration loop (slot_attn.py:64-102) for bf16 k/v.
//
// A cluster of CL CTAs works on TWO images at a time ("lanes").  Every CTA runs
//   * a PASS engine  (warps 0-7: 4 logit warps + 4 U warps) that streams k/v tiles of its token share through a
//     TMA ring and produces the partial sum_n w v and sum_n w of one (image, iteration), and
//   * an UPDATE engine (warps 8-15) that runs the slot update of the other lane: reduce-scatter over the
//     cluster, GRU, residual MLP, next q -- six DSMEM exchange rounds with tensor-core matrix-vector products
//     against the CTA's weight slice, which stays in shared memory for the whole kernel.
// The engines alternate lanes (pass A0 | pass B0 + update A0 | pass A1 + update B0 | ...), so the
// latency-bound update of one image hides behind the bandwidth-bound pass of the other and the ring never
// drains: tiles of the next pass are already in flight when a pass ends.  Hand-offs are shared-memory
// mbarriers (u_ready / u_free / q_ready); the cluster exchange is st.async + complete_tx, double-buffered
// by round parity (see pc_common.cuh / sa_iter_fwd_pc.cu for the single-engine form and the protocol).
// Only 2 * (#clusters) images are in flight, so passes 2..T read k/v from L2.
#include "pc_common.cuh"

namespace ocrl {
namespace pipe {

using namespace pc;

#ifndef TRACE_OP
#define TRACE_OP 2
#endif

template <int D_, int H_, int CL_, int S_, int NL_, int KB_>
struct Cfg {
  // NL image lanes per cluster; KB rows in the slot-indexed buffers (6 when K <= 6: room for one more ring stage)
  static constexpr int D = D_, H = H_, CL = CL_, S = S_, NL = NL_, KB = KB_;
  static constexpr int KP = 8, NT = 512, TOK = 16;
  static constexpr int PITCH = D * 2 + 16, PITCHH = H * 2 + 16;
  static constexpr int LX = D > H ? D : H;
  static constexpr int PITCHX = LX * 2 + 16;
  static constexpr int TILE_BYTES = TOK * D * 2, STAGE_BYTES = 2 * TILE_BYTES, WT_BYTES = 256;
  static constexpr int DS = D / CL, HS = H / CL;
  static constexpr int NMU = D / 16, NMG = (3 * DS + 15) / 16, NM1 = (HS + 15) / 16, NM2 = (DS + 15) / 16, NKC = 4;
  static constexpr int UP = D + 4;
  static_assert(D % 64 == 0 && H % 64 == 0 && DS % 4 == 0 && HS % 4 == 0, "shape");
  static_assert(TILE_BYTES % 1024 == 0, "swizzle atoms are 1024 bytes");

  static constexpr int OFF_RING = 0;
  static constexpr int OFF_WT = OFF_RING + S * STAGE_BYTES;
  static constexpr int OFF_WIH = OFF_WT + S * WT_BYTES;
  static constexpr int OFF_WHH = OFF_WIH + 3 * DS * PITCH;
  static constexpr int OFF_W1 = OFF_WHH + 3 * DS * PITCH;
  static constexpr int OFF_W2 = OFF_W1 + HS * PITCH;
  static constexpr int OFF_WQ = OFF_W2 + DS * PITCHH;
  static constexpr int OFF_ZROW = OFF_WQ + DS * PITCH;
  static constexpr int OFF_BIAS = (OFF_ZROW + PITCHX + 15) & ~15;  // (the LayerNorm affine parameters are folded into W1', Wq')
  // fp32 constants: b_ih[3DS] b_hh[3DS] b1'[HS] b2[DS] c1[HS] cq[DS] bq'[DS] + LayerNorm stats mean[8] rstd[8]
  static constexpr int NCONST = 9 * DS + 2 * HS + 16;
  static constexpr int XBUF_BYTES = KB * D * 4 + 32 * CL;          // R1 receive buffer (fp32 partial sums), single
  static constexpr int OFF_XBUF = (OFF_BIAS + NCONST * 4 + 15) & ~15;
  static constexpr int OFF_ACT = OFF_XBUF + XBUF_BYTES;            // bf16 all-gather targets, by round parity
  static constexpr int P_ROWS = (3 * NMG > NKC * (NM1 > NM2 ? NM1 : NM2)) ? 3 * NMG * 16 : NKC * (NM1 > NM2 ? NM1 : NM2) * 16;
  static constexpr int OFF_UST = OFF_ACT + 2 * KB * PITCHX;        // U staging a, b (pass -> update hand-off)
  static constexpr int OFF_LANE = OFF_UST + 2 * KB * UP * 4;       // per lane: slots (bf16), q (bf16), own slice (fp32)
  static constexpr int LANE_BYTES = KB * PITCH + KB * PITCH + KB * DS * 4;
  // rows KB..7 of the slot-indexed buffers are read as (ignored) padding columns of the MMAs, so plain data follows
  // them: the MMA partial outputs and the token sums close the list
  static constexpr int OFF_P = (OFF_LANE + NL * LANE_BYTES + 15) & ~15;
  static constexpr int OFF_SRED = OFF_P + P_ROWS * 32;             // [4][8] token sums of the logit warps
  static constexpr int OFF_BAR = OFF_SRED + 128;  // full[S] w_ready[S] xbar[2] u_ready u_free q_ready[NL]
  static constexpr int OFF_ISSUED = OFF_BAR + (2 * S + 4 + NL) * 8;
  static constexpr int SMEM_BYTES = OFF_ISSUED + S * 4;
};

__device__ __forceinline__ void upd_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

template <int D, int H, int CL, int S, int NL, int KB>
__global__ void __launch_bounds__(512, 1)
sa_iter_fwd_pipe_kernel(const IterFwdArgs a, const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v) {
  using C = Cfg<D, H, CL, S, NL, KB>;
  constexpr int PITCH = C::PITCH, PITCHH = C::PITCHH, PITCHX = C::PITCHX, DS = C::DS, HS = C::HS;
  constexpr int NMU = C::NMU, NMG = C::NMG, NM1 = C::NM1, NM2 = C::NM2, NKC = C::NKC, UP = C::UP, TOK = C::TOK;
  constexpr float LOG2E = 1.4426950408889634f;

  extern __shared__ __align__(1024) unsigned char sm[];  // swizzled TMA tiles need 1024-byte alignment
  if ((smem_u32(sm) & 1023u) != 0) __trap();
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / CL, ncl = gridDim.x / CL;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g8 = lane >> 2, t4 = lane & 3;
  const int K = a.K, N = a.N, T = a.T, B = a.B;
  const bool tracer = (a.trace != nullptr && blockIdx.x == 0);
#define PP_TRACE(i) do { if (tracer) a.trace[(i)] = clock64(); } while (0)

  unsigned char* ring = sm + C::OFF_RING;
  unsigned char* wtiles = sm + C::OFF_WT;
  unsigned char* s_wih = sm + C::OFF_WIH;
  unsigned char* s_whh = sm + C::OFF_WHH;
  unsigned char* s_w1 = sm + C::OFF_W1;
  unsigned char* s_w2 = sm + C::OFF_W2;
  unsigned char* s_wq = sm + C::OFF_WQ;
  unsigned char* s_zrow = sm + C::OFF_ZROW;
  float* s_bias = reinterpret_cast<float*>(sm + C::OFF_BIAS);
  const float* s_bih = s_bias;
  const float* s_bhh = s_bias + 3 * DS;
  float* s_b1f = s_bias + 6 * DS;            // b1 + W1 . beta_mlp
  const float* s_b2 = s_bias + 6 * DS + HS;
  float* s_c1 = s_bias + 7 * DS + HS;        // row sums of the folded bf16 W1'
  float* s_cq = s_bias + 7 * DS + 2 * HS;    // row sums of the folded bf16 Wq'
  float* s_bqf = s_bias + 8 * DS + 2 * HS;   // Wq . beta_slots
  float* s_mean = s_bias + 9 * DS + 2 * HS;  // LayerNorm statistics of the round in flight
  float* s_rstd = s_mean + 8;
  unsigned char* xbuf = sm + C::OFF_XBUF;
  auto act = [&](uint32_t r) { return sm + C::OFF_ACT + (r & 1) * (KB * PITCHX); };  // bf16 [KB][PITCHX]
  float* P = reinterpret_cast<float*>(sm + C::OFF_P);
  float* ustage_a = reinterpret_cast<float*>(sm + C::OFF_UST);
  float* ustage_b = ustage_a + KB * UP;
  float* sred = reinterpret_cast<float*>(sm + C::OFF_SRED);
  auto slh_hi = [&](int l) { return sm + C::OFF_LANE + l * C::LANE_BYTES; };
  auto qbuf = [&](int l) { return sm + C::OFF_LANE + l * C::LANE_BYTES + KB * PITCH; };
  auto own_of = [&](int l) { return reinterpret_cast<float*>(sm + C::OFF_LANE + l * C::LANE_BYTES + 2 * KB * PITCH); };
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + C::OFF_BAR);
  uint64_t* full = bars;
  uint64_t* w_ready = bars + S;
  uint64_t* xbar = bars + 2 * S;
  uint64_t* u_ready = bars + 2 * S + 2;
  uint64_t* u_free = bars + 2 * S + 3;
  uint64_t* q_ready = bars + 2 * S + 4;
  int* issued = reinterpret_cast<int*>(sm + C::OFF_ISSUED);

  // ------------------------------------------------------------------ work assignment (two lanes per cluster)
  // Images are dealt to clusters round-robin (newest first: the projection kernel left them in L2) and a cluster
  // alternates its images between its two lanes, so every cluster gets floor or ceil of B / #clusters images.
  const int ncimg = (B > cid) ? (B - cid + ncl - 1) / ncl : 0;
  // Lane l takes the cluster's images l, l + NL, ...; ops (one pass + one update of one lane) are streamed row by row:
  // row c holds op c of every lane that still has one, lanes in order (lane counts are non-increasing in l and
  // differ by at most one image).  With NL = 3 an update has two pass slots to finish in, so the stream advances at
  // the pace of the pass engine alone.
  int nops[NL], row_start[NL + 1], row_width[NL + 1], seg_base[NL + 1];
  auto nimg_of = [&](int l) { return (ncimg > l) ? (ncimg - l + NL - 1) / NL : 0; };
#pragma unroll
  for (int l = 0; l < NL; ++l) nops[l] = nimg_of(l) * T;
  // segments of rows with constant width: width NL while c < nops[NL-1], NL-1 while c < nops[NL-2], ...
  int total_ops = 0;
  {
    int prev = 0;
#pragma unroll
    for (int g = 0; g < NL; ++g) {
      const int w = NL - g;               // active lanes in this segment
      const int upto = nops[w - 1];       // rows [prev, upto)
      row_start[g] = prev;
      row_width[g] = w;
      seg_base[g] = total_ops;
      total_ops += (upto - prev) * w;
      prev = upto;
    }
    row_start[NL] = prev; row_width[NL] = 1; seg_base[NL] = total_ops;
  }
  auto op_of = [&](int n, int& l, int& c) {
    l = 0; c = 0;
#pragma unroll
    for (int g = 0; g < NL; ++g) {
      if (n >= seg_base[g] && n < seg_base[g + 1]) {
        const int r = n - seg_base[g];
        c = row_start[g] + r / row_width[g];
        l = r % row_width[g];
      }
    }
  };
  auto image_of = [&](int l, int m) { return B - 1 - (cid + (NL * m + l) * ncl); };
  const int ntiles = (N + TOK - 1) / TOK;
  const int TPC = (ntiles + CL - 1) / CL;
  const int tile0 = rank * TPC;
  const int TP = max(0, min(TPC, ntiles - tile0));
  // tile j is handled by chain j % 4 and lives in stage j % S: with S % 4 == 0 and TP % 4 == 0 a warp only ever sees its own stages
  const bool guard = (S % 4 != 0) || (TP % 4 != 0);

  // one elected lane: stage (j % S) <- tile j = n * TP + tile of the CTA's tile sequence (pass ops in stream order).
  // L2 policy: tiles of iterations 0..T-2 are read again by the next pass (evict_last), the final pass's are dead
  // afterwards (evict_first).  For the first pass of an image the tile `PFD` tiles further on is prefetched into L2
  // at the same rate as tiles are consumed (a burst would queue in front of the demand loads).
  const uint64_t pol_keep = l2_policy_evict_last(), pol_drop = l2_policy_evict_first();
  constexpr int PFD = 24;
  auto issue_tile = [&](int n, int tile) {
    const int j = n * TP + tile;
    const int s = j % S;
    int l, c;
    op_of(n, l, c);
    const int t = c % T;
    const int row0 = image_of(l, c / T) * N + (tile0 + tile) * TOK;
    unsigned char* kd = ring + (size_t)s * C::STAGE_BYTES;
    const uint64_t pol = (t == T - 1) ? pol_drop : pol_keep;
    mbar_expect_tx(&full[s], (uint32_t)C::STAGE_BYTES);
    tma_load_3d_hint(kd, &tm_k, 0, 0, row0, &full[s], pol);
    tma_load_3d_hint(kd + C::TILE_BYTES, &tm_v, 0, 0, row0, &full[s], pol);
    sts_volatile(&issued[s], j);  // after the expect_tx in program order: the barrier is in this tile's phase
    // paced L2 prefetch, only where the data is cold (first pass of an image)
    int pn = n, pt = tile + PFD;
    while (pt >= TP && pn < total_ops) { pt -= TP; ++pn; }
    if (pn < total_ops) {
      int pl, pc;
      op_of(pn, pl, pc);
      if (pc % T == 0) {
        const size_t off = ((size_t)image_of(pl, pc / T) * N + (size_t)(tile0 + pt) * TOK) * D * 2;
        if (off + C::TILE_BYTES <= (size_t)B * N * D * 2) {
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const char*>(a.k) + off), "r"(C::TILE_BYTES) : "memory");
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const char*>(a.v) + off), "r"(C::TILE_BYTES) : "memory");
        }
      }
    }
  };
  // Per-op constants of the refill path (row base of the CTA's share, L2 policy, cold = first pass of an image),
  // computed once per op so that the per-tile refill is a handful of instructions.
  struct OpInfo { int row_base; int cold; int valid; uint64_t pol; };
  auto op_info = [&](int n) {
    OpInfo o;
    o.valid = (n < total_ops);
    o.row_base = 0; o.cold = 0; o.pol = pol_keep;
    if (o.valid) {
      int l, c;
      op_of(n, l, c);
      const int t = c % T;
      o.row_base = image_of(l, c / T) * N + tile0 * TOK;
      o.cold = (t == 0);
      o.pol = (t == T - 1) ? pol_drop : pol_keep;
    }
    return o;
  };
  // refill of stage j % S with tile (op, tile) and paced prefetch; `cur` / `nxt` describe ops n and n + 1
  auto refill_fast = [&](int n, int tile, const OpInfo& cur, const OpInfo& nxt) {
    int tn = tile + S;
    const OpInfo& o = (tn < TP) ? cur : nxt;
    const int nn = (tn < TP) ? n : n + 1;
    if (tn >= TP) tn -= TP;
    if (o.valid) {
      const int j = nn * TP + tn;
      const int s = j % S;
      const int row0 = o.row_base + tn * TOK;
      unsigned char* kd = ring + (size_t)s * C::STAGE_BYTES;
      mbar_expect_tx(&full[s], (uint32_t)C::STAGE_BYTES);
      tma_load_3d_hint(kd, &tm_k, 0, 0, row0, &full[s], o.pol);
      tma_load_3d_hint(kd + C::TILE_BYTES, &tm_v, 0, 0, row0, &full[s], o.pol);
      sts_volatile(&issued[s], j);
    }
    int pt = tile + PFD;
    const OpInfo& po = (pt < TP) ? cur : nxt;
    if (pt >= TP) pt -= TP;
    if (po.valid && po.cold) {
      const size_t off = (size_t)(po.row_base + pt * TOK) * (D * 2);
      if (off + C::TILE_BYTES <= (size_t)B * N * D * 2) {
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const char*>(a.k) + off), "r"(C::TILE_BYTES) : "memory");
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const char*>(a.v) + off), "r"(C::TILE_BYTES) : "memory");
      }
    }
  };
  // tile (n, tile) + S in stream order, or n = -1 past the end
  auto advance = [&](int& n, int& tile) {
    tile += S;
    while (tile >= TP) { tile -= TP; ++n; }
    if (n >= total_ops) n = -1;
  };

  // ------------------------------------------------------------------ one-time setup
  // The slots of lane 0's first image are fetched now and written to shared memory at the end of the setup: their global
  // round trip hides under the weight load instead of standing in front of the first query
  constexpr int PRE2 = (8 * (D / 2) + 255) / 256, PRE1 = (8 * DS + 255) / 256;
  float2 pre2[PRE2];
  float pre1[PRE1];
  if (tid >= 256 && nops[0] > 0) {
    const float* src = a.slots0 + (size_t)image_of(0, 0) * K * D;
#pragma unroll
    for (int u = 0; u < PRE2; ++u) {
      const int i = tid - 256 + 256 * u;
      if (i < K * (D / 2)) pre2[u] = __ldg(reinterpret_cast<const float2*>(src + (i / (D / 2)) * D + 2 * (i % (D / 2))));
    }
#pragma unroll
    for (int u = 0; u < PRE1; ++u) {
      const int i = tid - 256 + 256 * u;
      if (i < K * DS) pre1[u] = __ldg(src + (i / DS) * D + rank * DS + i % DS);
    }
  }
  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&w_ready[s], 1);
      sts_volatile(&issued[s], -1);
    }
    mbar_init(&xbar[0], 1);
    mbar_init(&xbar[1], 1);
    mbar_init(u_ready, 8);
    mbar_init(u_free, 8);
    for (int l = 0; l < NL; ++l) mbar_init(&q_ready[l], 1);
    mbar_fence_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_v) : "memory");
  }
  // first tiles: their HBM latency overlaps the weight load below.  A tile costs its issuing lane several hundred
  // cycles, so the pass warps issue one (or two) each, behind a barrier of their own that publishes the mbarrier init
  if (warp < 8) {
    asm volatile("bar.sync 3, 256;" ::: "memory");
    if (lane == 0 && TP > 0)
      for (int p = warp; p < S; p += 8)
        if (p / TP < total_ops) issue_tile(p / TP, p % TP);
  }
  {
    // the CTA's weight slices, fp32 global -> bf16 shared; eight independent 16-byte loads in flight per thread
    auto load_rows = [&](unsigned char* dst, int pitch, const float* src, int L, int nrows) {
      const int total = nrows * (L / 4);
      for (int i0 = tid; i0 < total; i0 += 8 * C::NT) {
        float4 x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = i0 + u * C::NT;
          if (i < total) x[u] = __ldg(reinterpret_cast<const float4*>(src) + i);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
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
    load_rows(s_w2, PITCHH, a.w.w2 + (size_t)rank * DS * H, H, DS);
    // LayerNorm folded into the product that follows it:  W LN(x) = rstd (W' x - mean c) + W beta,  W' = W diag(gamma),
    // c = row sums of the (bf16-rounded) W'.  One warp per row.
    {
      constexpr int NCH = D / 64;
      for (int r = warp; r < HS + DS; r += C::NT / 32) {
        const bool is1 = r < HS;
        const int rr = is1 ? r : r - HS;
        const float* wrow = is1 ? a.w.w1 + ((size_t)rank * HS + rr) * D : a.w.wq + ((size_t)rank * DS + rr) * D;
        const float* gam = is1 ? a.w.ln_mlp_w : a.w.ln_slots_w;
        const float* bet = is1 ? a.w.ln_mlp_b : a.w.ln_slots_b;
        unsigned char* drow = (is1 ? s_w1 : s_wq) + rr * PITCH;
        float csum = 0.f, bsum = 0.f;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const int d = 64 * c + 2 * lane;
          const float2 wv = __ldg(reinterpret_cast<const float2*>(wrow + d));
          const float2 g = __ldg(reinterpret_cast<const float2*>(gam + d));
          const float2 b = __ldg(reinterpret_cast<const float2*>(bet + d));
          const __nv_bfloat162 wf = __floats2bfloat162_rn(wv.x * g.x, wv.y * g.y);
          *reinterpret_cast<__nv_bfloat162*>(drow + d * 2) = wf;
          csum += __low2float(wf) + __high2float(wf);
          bsum = fmaf(wv.x, b.x, fmaf(wv.y, b.y, bsum));
        }
        csum = warp_sum(csum);
        bsum = warp_sum(bsum);
        if (lane == 0) {
          if (is1) { s_c1[rr] = csum; s_b1f[rr] = bsum + a.w.b1[rank * HS + rr]; }
          else { s_cq[rr] = csum; s_bqf[rr] = bsum; }
        }
      }
    }
    for (int i = tid; i < PITCHX / 4; i += C::NT) reinterpret_cast<uint32_t*>(s_zrow)[i] = 0u;
    for (int i = tid; i < 3 * DS; i += C::NT) {
      const int gate = i / DS, dl = i % DS;
      s_bias[i] = a.w.b_ih[gate * D + rank * DS + dl];
      s_bias[3 * DS + i] = a.w.b_hh[gate * D + rank * DS + dl];
    }
    for (int i = tid; i < DS; i += C::NT) s_bias[6 * DS + HS + i] = a.w.b2[rank * DS + i];
    // staging rows of the padded slots are never written by the conversions; keep them finite
    for (int i = tid; i < (2 * KB * PITCHX) / 4; i += C::NT) reinterpret_cast<uint32_t*>(sm + C::OFF_ACT)[i] = 0u;
    for (int i = tid; i < (NL * C::LANE_BYTES) / 4; i += C::NT) reinterpret_cast<uint32_t*>(sm + C::OFF_LANE)[i] = 0u;
  }
  __syncthreads();
  if (tid >= 256 && nops[0] > 0) {  // after the zero fill of the lane buffers; the cluster barrier below publishes it
    unsigned char* dst = slh_hi(0);
    float* own = own_of(0);
#pragma unroll
    for (int u = 0; u < PRE2; ++u) {
      const int i = tid - 256 + 256 * u;
      if (i < K * (D / 2)) *reinterpret_cast<uint32_t*>(dst + (i / (D / 2)) * PITCH + 4 * (i % (D / 2))) = pack_bf16x2(pre2[u].x, pre2[u].y);
    }
#pragma unroll
    for (int u = 0; u < PRE1; ++u) {
      const int i = tid - 256 + 256 * u;
      if (i < K * DS) own[i] = pre1[u];
    }
  }
  cluster.sync();  // every CTA's barriers are initialised before any peer signals them
  if (tid == 0) PP_TRACE(0);

  if (warp < 8) {
    // ==================================================================================== PASS ENGINE
    for (int n = 0; n < total_ops; ++n) {
      int l, c;
      op_of(n, l, c);
      const int t = c % T, img = image_of(l, c / T);
      const bool last = (t == T - 1);
      const int jbase = n * TP;
      if (warp < 4) {
        // ---- logit warp
        mbar_wait(&q_ready[l], (uint32_t)(c & 1));
        if (tid == 0 && n < 40) PP_TRACE(8 + n * 8);
        const unsigned char* qsrc = qbuf(l);
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
          const int j = jbase + tile;
          const int s = j % S;
          const uint32_t ph = (uint32_t)((j / S) & 1);
          const unsigned char* kt = ring + (size_t)s * C::STAGE_BYTES;
          uint32_t* wt = reinterpret_cast<uint32_t*>(wtiles + s * C::WT_BYTES);
          const bool tt = tracer && tid == 0 && n == TRACE_OP && tile < 32;  // tile stamps of chain 0 in a steady-state op
          if (tt) a.trace[340 + (tile >> 2) * 8 + 0] = clock64();
          // mbarrier parity waits alias two phases ahead: unless a warp always returns to the same stages, make sure the
          // barrier has entered this tile's phase first
          if (guard) while (lds_volatile(&issued[s]) < j) __nanosleep(20);
          mbar_wait(&full[s], ph);
          if (tt) a.trace[340 + (tile >> 2) * 8 + 1] = clock64();
          float ca[4] = {0.f, 0.f, 0.f, 0.f}, cb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int ks = 0; ks < D / 16; ++ks) {
            uint32_t kf[4];
            ldmatrix_x4(kf, kt + swz_off<D>(lrow, ks * 32 + lhalf));
            if (ks & 1) mma_bf16_16816(cb, kf, qb[ks][0], qb[ks][1]);
            else mma_bf16_16816(ca, kf, qb[ks][0], qb[ks][1]);
          }
          if (tt) a.trace[340 + (tile >> 2) * 8 + 2] = clock64();
          // softmax over the slot axis: a token's 8 logits live in the 4 lanes of a quad (2 each).  Both token rows
          // of the fragment go through the reductions together (straight-line code, the two chains interleave);
          // the divergent attn_vis stores come afterwards.
          const int tok = (tile0 + tile) * TOK + g8;
          float x0[2], x1[2], mx[2], e0[2], e1[2], sum[2];
#pragma unroll
          for (int hrow = 0; hrow < 2; ++hrow) {
            x0[hrow] = ok0 ? ca[2 * hrow] + cb[2 * hrow] : -INFINITY;
            x1[hrow] = ok1 ? ca[2 * hrow + 1] + cb[2 * hrow + 1] : -INFINITY;
            mx[hrow] = fmaxf(x0[hrow], x1[hrow]);
          }
#pragma unroll
          for (int o = 1; o <= 2; o <<= 1) {
            const float m0 = __shfl_xor_sync(FULL, mx[0], o), m1 = __shfl_xor_sync(FULL, mx[1], o);
            mx[0] = fmaxf(mx[0], m0);
            mx[1] = fmaxf(mx[1], m1);
          }
#pragma unroll
          for (int hrow = 0; hrow < 2; ++hrow) {
            e0[hrow] = ex2f(x0[hrow] - mx[hrow]);
            e1[hrow] = ex2f(x1[hrow] - mx[hrow]);
            sum[hrow] = e0[hrow] + e1[hrow];
          }
#pragma unroll
          for (int o = 1; o <= 2; o <<= 1) {
            const float s0 = __shfl_xor_sync(FULL, sum[0], o), s1 = __shfl_xor_sync(FULL, sum[1], o);
            sum[0] += s0;
            sum[1] += s1;
          }
          float w[4];
          float av[4];
#pragma unroll
          for (int hrow = 0; hrow < 2; ++hrow) {
            const float inv = __fdividef(1.f, sum[hrow]);
            av[2 * hrow] = e0[hrow] * inv;
            av[2 * hrow + 1] = e1[hrow] * inv;
            const bool tok_ok = (tok + 8 * hrow) < N;
            w[2 * hrow] = (tok_ok && ok0) ? av[2 * hrow] + a.eps : 0.f;
            w[2 * hrow + 1] = (tok_ok && ok1) ? av[2 * hrow + 1] + a.eps : 0.f;
          }
          if (last && a.attn_out != nullptr) {
#pragma unroll
            for (int hrow = 0; hrow < 2; ++hrow) {
              const int tk = tok + 8 * hrow;
              if (tk < N) {
                float* ao = a.attn_out + ((size_t)img * N + tk) * K;
                if ((K & 1) == 0) {
                  if (ok0) __stcs(reinterpret_cast<float2*>(ao + c0), make_float2(av[2 * hrow], av[2 * hrow + 1]));
                } else {
                  if (ok0) ao[c0] = av[2 * hrow];
                  if (ok1) ao[c1] = av[2 * hrow + 1];
                }
              }
            }
          }
          if (tt) a.trace[420 + (tile >> 2) * 4 + 0] = clock64();
          // weights rounded to bf16 once; the same rounded values feed the numerator and the token sum
          const __nv_bfloat162 p0 = __floats2bfloat162_rn(w[0], w[1]);
          const __nv_bfloat162 p1 = __floats2bfloat162_rn(w[2], w[3]);
          Sl0 += __low2float(p0) + __low2float(p1);
          Sl1 += __high2float(p0) + __high2float(p1);
          wt[lane] = movmatrix_trans(*reinterpret_cast<const uint32_t*>(&p0));
          wt[32 + lane] = movmatrix_trans(*reinterpret_cast<const uint32_t*>(&p1));
          if (tt) a.trace[420 + (tile >> 2) * 4 + 1] = clock64();
          __syncwarp();
          if (tt) a.trace[420 + (tile >> 2) * 4 + 2] = clock64();
          if (lane == 0) mbar_arrive(&w_ready[s]);
          if (tt) a.trace[340 + (tile >> 2) * 8 + 3] = clock64();
        }
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          Sl0 += __shfl_xor_sync(FULL, Sl0, o);
          Sl1 += __shfl_xor_sync(FULL, Sl1, o);
        }
        if (n > 0) mbar_wait(u_free, (uint32_t)((n - 1) & 1));  // the previous update has read sred / the U staging
        if (g8 == 0) {
          sred[warp * 8 + c0] = Sl0;
          sred[warp * 8 + c1] = Sl1;
        }
        if (tid == 0 && n < 40) PP_TRACE(8 + n * 8 + 1);
      } else {
        // ---- U warp: all D features of its tiles
        const int uw = warp - 4;
        float acc[NMU][4];
#pragma unroll
        for (int i = 0; i < NMU; ++i)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[i][e] = 0.f;
        const int urow = (lane & 7) + (lane >> 4) * 8, uhalf = ((lane >> 3) & 1) * 16;
        const bool fast = (S <= TP) && (PFD <= TP);
        const OpInfo oi_cur = op_info(n), oi_nxt = op_info(n + 1);
        for (int tile = uw; tile < TP; tile += 4) {
          const int j = jbase + tile;
          const int s = j % S;
          const uint32_t ph = (uint32_t)((j / S) & 1);
          const unsigned char* vt = ring + (size_t)s * C::STAGE_BYTES + C::TILE_BYTES;
          const uint32_t* wt = reinterpret_cast<const uint32_t*>(wtiles + s * C::WT_BYTES);
          const bool tt = tracer && tid == 128 && n == TRACE_OP && tile < 32;
          if (tt) a.trace[340 + (tile >> 2) * 8 + 4] = clock64();
          if (guard) while (lds_volatile(&issued[s]) < j) __nanosleep(20);
          mbar_wait(&w_ready[s], ph);  // implies full[s]: the logit warp waited for k and v together
          if (tt) a.trace[340 + (tile >> 2) * 8 + 5] = clock64();
          const uint32_t b0 = wt[lane], b1 = wt[32 + lane];
#pragma unroll
          for (int i = 0; i < NMU; ++i) {
            uint32_t vf[4];
            ldmatrix_x4_trans(vf, vt + swz_off<D>(urow, uhalf + i * 32));
            mma_bf16_16816(acc[i], vf, b0, b1);
          }
          __syncwarp();  // every lane is done with the stage (this warp is its only remaining reader)
          if (tt) a.trace[340 + (tile >> 2) * 8 + 6] = clock64();
          if (lane == 0) {
            if (fast) {
              refill_fast(n, tile, oi_cur, oi_nxt);
            } else {
              int nn = n, tn = tile;
              advance(nn, tn);
              if (nn >= 0) issue_tile(nn, tn);
            }
          }
          if (tt) a.trace[340 + (tile >> 2) * 8 + 7] = clock64();
        }
        if (n > 0) mbar_wait(u_free, (uint32_t)((n - 1) & 1));
        // combine the four partial sums: warps 4, 5 write the two staging buffers, warps 6, 7 add into them
        float* ust = (uw & 1) ? ustage_b : ustage_a;
        if (uw >= 2) asm volatile("bar.sync 2, 128;" ::: "memory");
#pragma unroll
        for (int i = 0; i < NMU; ++i) {
          const int d0 = 16 * i + g8;
          float* u0 = ust + (2 * t4) * UP + d0;
          float* u1 = ust + (2 * t4 + 1) * UP + d0;
          if (2 * t4 < KB) {  // rows KB..7 are padding slots
            if (uw >= 2) {
              u0[0] += acc[i][0]; u1[0] += acc[i][1]; u0[8] += acc[i][2]; u1[8] += acc[i][3];
            } else {
              u0[0] = acc[i][0]; u1[0] = acc[i][1]; u0[8] = acc[i][2]; u1[8] = acc[i][3];
            }
          }
        }
        if (uw < 2) asm volatile("bar.sync 2, 128;" ::: "memory");
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(u_ready);
    }
  } else {
    // ==================================================================================== UPDATE ENGINE
    // Every exchange round is  push -> wait -> tensor-core product -> one barrier -> epilogue (= next push):
    // all-gather payloads travel as bf16 straight into the receivers' MMA staging buffers (act[parity], slh[lane],
    // qbuf[lane]); the LayerNorms are folded into the products that follow them (statistics from the received rows).
    const int utid = tid - 256, uwarp = warp - 8;
    uint32_t round = 0;
    // training: per-(image, iteration) state for the fused backward (SavedLayout, slot_math.cuh); every CTA writes
    // its own feature slice
    const SavedLayout SL(K, D, H);
    auto saved_at = [&](int img, int t) { return a.saved + ((size_t)img * T + t) * SL.stride(); };
    unsigned char* xb = xbuf;
    auto arm = [&](uint32_t r, uint32_t bytes) {
      if (utid == 0) mbar_expect_tx(&xbar[r & 1], bytes);
    };
    auto xwait = [&](uint32_t r) { mbar_wait_cluster(&xbar[r & 1], (r >> 1) & 1); };
    // All-gather push: thread i < K*SL holds element (slot = i / SL, f = i % SL) of the CTA's slice.  The four lanes
    // of a quad assemble 4 features (8 bytes of bf16) and each sends them to CL/4 CTAs, into `dst` ([8][pitch] bf16).
    auto quad_push = [&](uint32_t r, float val, int i, int SL, unsigned char* dst, int pitch) {
      const int qb = lane & ~3;
      const float v0 = __shfl_sync(FULL, val, qb), v1 = __shfl_sync(FULL, val, qb + 1);
      const float v2 = __shfl_sync(FULL, val, qb + 2), v3 = __shfl_sync(FULL, val, qb + 3);
      if (i < K * SL) {
        const int slot = i / SL, f4 = (i % SL) & ~3;
        const uint32_t lbuf = smem_u32(dst) + (uint32_t)(slot * pitch + (rank * SL + f4) * 2);
        const uint32_t lbar = smem_u32(&xbar[r & 1]);
        const uint32_t lo = pack_bf16x2(v0, v1), hi = pack_bf16x2(v2, v3);
#pragma unroll
        for (int q = 0; q < CL / 4; ++q) {
          const int dest = (lane & 3) + 4 * q;
          st_async_v2(mapa_u32(lbuf, dest), lo, hi, mapa_u32(lbar, dest));
        }
      }
    };
    // mean / rstd of the K rows (bf16, length D) that just arrived: one warp per row
    auto row_stats = [&](const unsigned char* rows, int pitch) {
      constexpr int NCH = D / 64;
      if (uwarp < K) {
        float2 x[NCH];
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(rows + uwarp * pitch + (64 * c + 2 * lane) * 2);
          x[c] = make_float2(__low2float(v), __high2float(v));
          s += x[c].x + x[c].y;
        }
        const float mean = warp_sum(s) * (1.f / D);
        float q = 0.f;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const float dx = x[c].x - mean, dy = x[c].y - mean;
          q = fmaf(dx, dx, fmaf(dy, dy, q));
        }
        const float rstd = rsqrtf(warp_sum(q) * (1.f / D) + a.ln_eps);
        if (lane == 0) { s_mean[uwarp] = mean; s_rstd[uwarp] = rstd; }
      }
    };
    // slots of image `img` (fp32, global) -> this CTA's bf16 copy slh[l] and the fp32 own slice
    auto load_slots0 = [&](int l, int img) {
      const float* src = a.slots0 + (size_t)img * K * D;
      unsigned char* dst = slh_hi(l);
      for (int i = utid; i < K * (D / 2); i += 256) {
        const int slot = i / (D / 2), c2 = i % (D / 2);
        const float2 x = __ldg(reinterpret_cast<const float2*>(src + slot * D + 2 * c2));
        *reinterpret_cast<uint32_t*>(dst + slot * PITCH + 4 * c2) = pack_bf16x2(x.x, x.y);
      }
      float* own = own_of(l);
      for (int i = utid; i < K * DS; i += 256) own[i] = __ldg(src + (i / DS) * D + rank * DS + i % DS);
      upd_sync();
    };
    // q = W_q LN(slots) of lane l for the CTA's slice (slots = slh[l], bf16), all-gathered (times log2 e) into qbuf[l]
    auto q_phase = [&](int l, int q_img, int q_t) {
      for (int job = uwarp; job < NM2 * NKC; job += 8) {
        const int mt = job / NKC, kc = job % NKC;
        mma_job<false>(s_wq, PITCH, DS, s_zrow, mt, kc * (D / 16) / NKC, (kc + 1) * (D / 16) / NKC, slh_hi(l), nullptr, PITCH,
                       P + kc * NM2 * 128, lane);
      }
      row_stats(slh_hi(l), PITCH);
      arm(round, (uint32_t)(K * D * 2));
      upd_sync();
      for (int i0 = uwarp * 32; i0 < K * DS; i0 += 256) {
        const int i = i0 + lane;
        float val = 0.f;
        if (i < K * DS) {
          const int slot = i / DS, dl = i % DS;
          float acc = 0.f;
#pragma unroll
          for (int kc = 0; kc < NKC; ++kc) acc += P[(kc * NM2 * 16 + dl) * 8 + slot];
          const float qv = s_rstd[slot] * (acc - s_mean[slot] * s_cq[dl]) + s_bqf[dl];
          if (a.saved != nullptr) saved_at(q_img, q_t)[SL.off_q() + slot * D + rank * DS + dl] = qv;
          val = qv * LOG2E;
        }
        quad_push(round, val, i, DS, qbuf(l), PITCH);
      }
      xwait(round);
      ++round;
      if (utid == 0) mbar_arrive(&q_ready[l]);  // the logit warps of lane l may load their q fragments
    };

    // initial queries of both lanes
    for (int l = 0; l < NL; ++l)
      if (nops[l] > 0) {
        if (l > 0) load_slots0(l, image_of(l, 0));  // lane 0's were staged during the setup
        q_phase(l, image_of(l, 0), 0);
      }

    for (int n = 0; n < total_ops; ++n) {
      int l, c;
      op_of(n, l, c);
      const int t = c % T, m = c / T, img = image_of(l, m);
      const bool last = (t == T - 1);
      float* own = own_of(l);
      const bool tr_on = tracer && utid == 0 && n < 40;
#define PP_T(i) do { if (tr_on) a.trace[8 + n * 8 + (i)] = clock64(); } while (0)
      // ============================================================ R1: reduce-scatter of sum w v, all-reduce of sum w
      arm(round, (uint32_t)(K * D * 4 + 32 * CL));
      mbar_wait(u_ready, (uint32_t)(n & 1));
      PP_T(2);
      {
        const uint32_t lbuf = smem_u32(xb), lbar = smem_u32(&xbar[round & 1]);
        constexpr int QPR = D / 4;
        for (int i = utid; i < K * QPR; i += 256) {
          const int slot = i / QPR, d = 4 * (i % QPR);
          const int dest = d / DS, dl = d % DS;
          const float4 xa = *reinterpret_cast<const float4*>(ustage_a + slot * UP + d);
          const float4 xb4 = *reinterpret_cast<const float4*>(ustage_b + slot * UP + d);
          const float4 val = make_float4(xa.x + xb4.x, xa.y + xb4.y, xa.z + xb4.z, xa.w + xb4.w);
          const uint32_t off = (uint32_t)(((rank * KB + slot) * DS + dl) * 4);
          st_async_v4(mapa_u32(lbuf + off, dest), val, mapa_u32(lbar, dest));
        }
        if (utid < 8 * CL) {
          const int slot = utid & 7, dest = utid >> 3;
          const float s = sred[slot] + sred[8 + slot] + sred[16 + slot] + sred[24 + slot];
          const uint32_t off = (uint32_t)(KB * D * 4 + (rank * 8 + slot) * 4);
          st_async_b32(mapa_u32(lbuf + off, dest), __float_as_uint(s), mapa_u32(lbar, dest));
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(u_free);  // the pass engine may overwrite the staging buffers
      // gh = W_hh h only needs the slots that entered the iteration: computed under the R1 round trip
      for (int job = uwarp; job < NMG; job += 8)
        mma_job<false>(s_whh, PITCH, 3 * DS, s_zrow, job, 0, D / 16, slh_hi(l), nullptr, PITCH, P + NMG * 128, lane);
      xwait(round);
      ++round;
      PP_T(3);
      // ============================================================ R2: all-gather updates = sum over CTAs / token sum
      arm(round, (uint32_t)(K * D * 2));
      {
        const float* rs = reinterpret_cast<const float*>(xb);
        const float* ss = rs + KB * D;
        for (int i0 = uwarp * 32; i0 < K * DS; i0 += 256) {
          const int i = i0 + lane;
          float val = 0.f;
          if (i < K * DS) {
            const int slot = i / DS, dl = i % DS;
            float u = 0.f, sw = 0.f;
#pragma unroll
            for (int src = 0; src < CL; ++src) {
              u += rs[(src * KB + slot) * DS + dl];
              sw += ss[src * 8 + slot];
            }
            val = u / sw;
            if (a.saved != nullptr) {
              float* sv = saved_at(img, t);
              sv[SL.off_h() + slot * D + rank * DS + dl] = own[i];  // slots entering the iteration
              sv[SL.off_u() + slot * D + rank * DS + dl] = val;
              if (rank == 0 && dl == 0) sv[SL.off_s() + slot] = sw;
            }
          }
          quad_push(round, val, i, DS, act(round), PITCHX);
        }
      }
      xwait(round);
      // ---- GRU: gi = W_ih u (k split in two halves: 2 * NMG jobs keep all 8 warps busy; halves summed in the epilogue)
      for (int job = uwarp; job < 2 * NMG; job += 8) {
        const int mt = job >> 1, half = job & 1;
        mma_job<false>(s_wih, PITCH, 3 * DS, s_zrow, mt, half * (D / 32), (half + 1) * (D / 32), act(round), nullptr, PITCHX,
                       P + (half ? 2 * NMG * 128 : 0), lane);
      }
      ++round;
      arm(round, (uint32_t)(K * D * 2));
      upd_sync();
      PP_T(4);
      // ============================================================ R3: all-gather h'
      for (int i0 = uwarp * 32; i0 < K * DS; i0 += 256) {
        const int i = i0 + lane;
        float hp = 0.f;
        if (i < K * DS) {
          const int slot = i / DS, dl = i % DS;
          const float* Pi = P;
          const float* Pj = P + 2 * NMG * 128;  // second k half of W_ih u
          const float* Ph = P + NMG * 128;
          const float gir = Pi[(dl) * 8 + slot] + Pj[(dl) * 8 + slot] + s_bih[dl], ghr = Ph[(dl) * 8 + slot] + s_bhh[dl];
          const float giz = Pi[(DS + dl) * 8 + slot] + Pj[(DS + dl) * 8 + slot] + s_bih[DS + dl];
          const float ghz = Ph[(DS + dl) * 8 + slot] + s_bhh[DS + dl];
          const float gin = Pi[(2 * DS + dl) * 8 + slot] + Pj[(2 * DS + dl) * 8 + slot] + s_bih[2 * DS + dl];
          const float ghn = Ph[(2 * DS + dl) * 8 + slot] + s_bhh[2 * DS + dl];
          const float r = sigmoidf_(gir + ghr), z = sigmoidf_(giz + ghz);
          const float nn = tanhf(gin + r * ghn);
          hp = (1.f - z) * nn + z * own[i];
          own[i] = hp;
          if (a.saved != nullptr) {
            float* sv = saved_at(img, t);
            const int f = slot * D + rank * DS + dl;
            sv[SL.off_r() + f] = r;
            sv[SL.off_z() + f] = z;
            sv[SL.off_n() + f] = nn;
            sv[SL.off_ghn() + f] = ghn;
            sv[SL.off_hp() + f] = hp;
          }
        }
        quad_push(round, hp, i, DS, act(round), PITCHX);
      }
      xwait(round);
      // ---- MLP layer 1 on the raw h' (LayerNorm folded), statistics alongside
      for (int job = uwarp; job < NM1 * NKC; job += 8) {
        const int mt = job / NKC, kc = job % NKC;
        mma_job<false>(s_w1, PITCH, HS, s_zrow, mt, kc * (D / 16) / NKC, (kc + 1) * (D / 16) / NKC, act(round), nullptr, PITCHX,
                       P + kc * NM1 * 128, lane);
      }
      row_stats(act(round), PITCHX);
      ++round;
      arm(round, (uint32_t)(K * H * 2));
      upd_sync();
      PP_T(5);
      // ============================================================ R4: all-gather the MLP hidden layer
      for (int i0 = uwarp * 32; i0 < K * HS; i0 += 256) {
        const int i = i0 + lane;
        float hid = 0.f;
        if (i < K * HS) {
          const int slot = i / HS, hl = i % HS;
          float acc = 0.f;
#pragma unroll
          for (int kc = 0; kc < NKC; ++kc) acc += P[(kc * NM1 * 16 + hl) * 8 + slot];
          const float pre = s_rstd[slot] * (acc - s_mean[slot] * s_c1[hl]) + s_b1f[hl];
          if (a.saved != nullptr) saved_at(img, t)[SL.off_pre() + slot * H + rank * HS + hl] = pre;
          hid = fmaxf(pre, 0.f);
        }
        quad_push(round, hid, i, HS, act(round), PITCHX);
      }
      xwait(round);
      for (int job = uwarp; job < NM2 * NKC; job += 8) {
        const int mt = job / NKC, kc = job % NKC;
        mma_job<false>(s_w2, PITCHH, DS, s_zrow, mt, kc * (H / 16) / NKC, (kc + 1) * (H / 16) / NKC, act(round), nullptr, PITCHX,
                       P + kc * NM2 * 128, lane);
      }
      ++round;
      if (!last) arm(round, (uint32_t)(K * D * 2));
      upd_sync();
      PP_T(6);
      // ============================================================ R5: all-gather the new slots (or write them out)
      const bool more = last && (m + 1 < nimg_of(l));
      for (int i0 = uwarp * 32; i0 < K * DS; i0 += 256) {
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
        if (!last) quad_push(round, sn, i, DS, slh_hi(l), PITCH);
      }
      if (!last) {
        xwait(round);
        ++round;
        q_phase(l, img, t + 1);  // R6: q of the lane's next iteration
      } else if (more) {
        upd_sync();  // every warp is done with `own` and P
        load_slots0(l, image_of(l, m + 1));
        q_phase(l, image_of(l, m + 1), 0);  // q of the lane's next image
      } else {
        upd_sync();  // P is rewritten by the next op
      }
      PP_T(7);
    }
  }
  __syncthreads();
  cluster.sync();  // no CTA leaves while a peer may still address its shared memory
}

template <int D, int H, int CL, int S, int NL, int KB>
static int launch_pipe(const IterFwdArgs& a, cudaStream_t stream) {
  using C = Cfg<D, H, CL, S, NL, KB>;
  auto kern = sa_iter_fwd_pipe_kernel<D, H, CL, S, NL, KB>;
  static_assert(C::SMEM_BYTES <= 227 * 1024, "shared memory budget");
  CUtensorMap tm_k, tm_v;
  if (!make_kv_map(&tm_k, a.k, (long long)a.B * a.N, D, C::TOK) || !make_kv_map(&tm_v, a.v, (long long)a.B * a.N, D, C::TOK)) {
    set_error("sa_iter_fwd(pipeline): cuTensorMapEncodeTiled failed");
    return OCRL_E_LAUNCH;
  }
  static bool configured = false;
  if (!configured) {
    OCRL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    if (CL > 8) OCRL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    configured = true;
  }
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
  static int max_clusters = -1;
  if (max_clusters < 0) {
    cfg.gridDim = dim3(CL * 148);
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
    if (e != cudaSuccess || n <= 0) {
      (void)cudaGetLastError();
      set_error("sa_iter_fwd(pipeline): cluster size %d with %d B of shared memory cannot be scheduled", CL, C::SMEM_BYTES);
      return OCRL_E_SHAPE;
    }
    max_clusters = n;
  }
  int ncl = max_clusters;
  if (a.max_clusters > 0) ncl = max(1, min(ncl, a.max_clusters));
  const int want = (a.B + NL - 1) / NL;  // NL lanes per cluster
  if (ncl > want) ncl = want;
  {  // fewest clusters that keep the same number of image rounds (frees SMs for concurrent work)
    const int per = (a.B + ncl - 1) / ncl;
    ncl = (a.B + per - 1) / per;
  }
  cfg.gridDim = dim3((unsigned)(ncl * CL));
  OCRL_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, a, tm_k, tm_v));
  ocrl::count_launch();
  return OCRL_OK;
}

}  // namespace pipe

// Returns OCRL_E_SHAPE (without launching) for shapes this kernel does not cover; the caller falls back.
int sa_iter_fwd_pipe_dispatch(const IterFwdArgs& a, cudaStream_t s) {
  if (a.K > 8) {
    set_error("sa_iter_fwd(pipeline): K <= 8");
    return OCRL_E_SHAPE;
  }
  if (a.D == 192 && a.H == 192) {
    if (a.K <= 6) {  // six slot rows in the exchange / staging buffers leave room for an eighth ring stage
      if (a.lanes == 2) return pipe::launch_pipe<192, 192, 8, 8, 2, 6>(a, s);
      return pipe::launch_pipe<192, 192, 8, 8, 3, 6>(a, s);
    }
    return pipe::launch_pipe<192, 192, 8, 7, 3, 8>(a, s);
  }
  if (a.D == 64 && a.H == 128) {  // the "Slot-Attention (small)" configuration: D = 64, H_mlp = 128
    // The whole weight set is 88 KB here, so clusters of four are enough: 33 of them fit (132 SMs against 120 with
    // clusters of eight) and the seven updates per image stop being the serial bottleneck of a 15-cluster grid.
    // One image is slower on four CTAs than on eight, so small batches stay on clusters of eight
    // (B = 64, T = 7: 146 against 191 us; B = 256: 545 against 653 us; B = 16: 145 against 119 us).
    if (a.lanes == 2 || (a.lanes == 0 && a.B >= 48)) return pipe::launch_pipe<64, 128, 4, 16, 2, 8>(a, s);
    if (a.K <= 6) return pipe::launch_pipe<64, 128, 8, 16, 3, 6>(a, s);
    return pipe::launch_pipe<64, 128, 8, 16, 3, 8>(a, s);
  }
  set_error("sa_iter_fwd(pipeline): D=%d H=%d not instantiated", a.D, a.H);
  return OCRL_E_SHAPE;
}

}  // namespace ocrl

// tcgen05 form of the fused T-iteration loop (slot_attn.py:64-102) for bf16 k/v: the token pass runs on the
// 5th-generation tensor cores with its accumulators in tensor memory.
//
// Structure as in sa_iter_fwd_pipe.cu: a cluster of CL CTAs works on NL images at a time ("lanes"); every CTA
// streams its share of the tokens (PASS engine, warps 0-7) while warps 8-15 run the slot update of another lane
// (UPDATE engine: reduce-scatter over the cluster, GRU, residual MLP, next q; weights stationary in shared memory).
// The pass engine here is
//   * warp 4 / warp 6, one lane each: TMA producers of the k ring and the v ring (64-token half tiles,
//     [D/64][64 rows][128 B] with the 128-byte swizzle = the canonical UMMA layouts: K-major for k, MN-major for v);
//   * warp 5, one lane: issues  logits[64 x 8] = k_half . q^T  (M = 64; the two halves of a 128-token pair land in
//     the lower / upper 16 lanes of every 32-lane quarter of one tensor-memory buffer) and
//     U^T[D x 16] += v_half^T . w  (M = 128 for features 0-127, M = 64 for 128-191), tcgen05.commit frees the ring
//     slots and signals the softmax warps;
//   * warps 0-3, one token per thread: tcgen05.ld of the token's logits, K-way softmax in registers (no shuffles),
//     attn_vis store in the last iteration, w = a + eps rounded to bf16 into the B-operand tile of the U product.
// q arrives from the update engine already in the UMMA K-major layout ([D/8][8 slots][8] bf16).
#include "pc_common.cuh"
#include "umma_common.cuh"

namespace ocrl {
namespace umma {

using namespace pc;

template <int D_, int H_, int CL_, int NL_, int KB_, int NKS_, int NVS_, int NWB_>
struct Cfg {
  // NL image lanes per cluster; KB rows in the slot-indexed buffers; NKS / NVS ring slots for k / v half tiles;
  // NWB buffers for the softmax weights of a 128-token pair
  static constexpr int D = D_, H = H_, CL = CL_, NL = NL_, KB = KB_, NKS = NKS_, NVS = NVS_, NWB = NWB_;
  static constexpr int KP = 8, NT = 512, HT = 64, NCH = D / 64;
  static constexpr int PITCH = D * 2 + 16, PITCHH = H * 2 + 16;
  static constexpr int LX = D > H ? D : H;
  static constexpr int PITCHX = LX * 2 + 16;
  static constexpr int CH_BYTES = HT * 128, HT_BYTES = NCH * CH_BYTES;  // one 64-wide feature chunk / one half tile
  static constexpr int WH_BYTES = 2048, WP_BYTES = 2 * WH_BYTES;        // w tile of a half ([8][16 slots][8] bf16) / a pair
  static constexpr int QOP_BYTES = D * 16;                              // q operand [D/8][8 slots][8] bf16
  static constexpr int DS = D / CL, HS = H / CL;
  static constexpr int NMG = (3 * DS + 15) / 16, NM1 = (HS + 15) / 16, NM2 = (DS + 15) / 16, NKC = 4;
  static constexpr int UP = D + 4;
  static constexpr int MA = NCH >= 2 ? 128 : 64;   // U product, features [0, 128) (or all 64)
  static constexpr bool HAS_B = NCH == 3;          // second U product, features [128, 192)
  static_assert(D % 64 == 0 && H % 64 == 0 && DS % 4 == 0 && HS % 4 == 0 && NCH <= 3, "shape");
  // tensor memory columns: two logit buffers, two U accumulators (features 0-127 | 128-191)
  static constexpr uint32_t COL_LG = 0, COL_U = 32, TMEM_COLS = 128;

  static constexpr int OFF_KRING = 0;
  static constexpr int OFF_VRING = OFF_KRING + NKS * HT_BYTES;
  static constexpr int OFF_WT = OFF_VRING + NVS * HT_BYTES;
  static constexpr int OFF_WIH = OFF_WT + NWB * WP_BYTES;
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
  static constexpr int OFF_UST = OFF_ACT + 2 * KB * PITCHX;        // U staging (pass -> update hand-off)
  static constexpr int OFF_LANE = (OFF_UST + KB * UP * 4 + 127) & ~127;  // per lane: q operand, slots (bf16), own slice (fp32)
  static constexpr int LANE_BYTES = (QOP_BYTES + KB * PITCH + KB * DS * 4 + 127) & ~127;
  // rows KB..7 of the slot-indexed buffers are read as (ignored) padding columns of the MMAs, so plain data follows
  // them: the MMA partial outputs and the token sums close the list
  static constexpr int OFF_P = (OFF_LANE + NL * LANE_BYTES + 15) & ~15;
  static constexpr int OFF_SRED = OFF_P + P_ROWS * 32;             // [4][8] token sums of the softmax warps
  // k_full k_empty [NKS] v_full v_empty [NVS] lg_full lg_empty w_full w_empty u_full u_accfree [2 each] xbar[2]
  // u_ready u_free q_ready[NL]
  static constexpr int OFF_BAR = OFF_SRED + 128;
  static constexpr int NBAR = 2 * NKS + 2 * NVS + 12 + 2 + 2 + NL;
  static constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
  static constexpr int SMEM_BYTES = OFF_TMEM + 16;
};

__device__ __forceinline__ void upd_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

template <int D, int H, int CL, int NL, int KB, int NKS, int NVS, int NWB>
__global__ void __launch_bounds__(512, 1)
sa_iter_fwd_umma_kernel(const IterFwdArgs a, const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v) {
  using C = Cfg<D, H, CL, NL, KB, NKS, NVS, NWB>;
  constexpr int PITCH = C::PITCH, PITCHH = C::PITCHH, PITCHX = C::PITCHX, DS = C::DS, HS = C::HS;
  constexpr int NMG = C::NMG, NM1 = C::NM1, NM2 = C::NM2, NKC = C::NKC, UP = C::UP, HT = C::HT, NCH = C::NCH;
  constexpr float LOG2E = 1.4426950408889634f;

  extern __shared__ __align__(1024) unsigned char sm[];  // swizzled TMA tiles need 1024-byte alignment
  if ((smem_u32(sm) & 1023u) != 0) __trap();
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / CL, ncl = gridDim.x / CL;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = a.K, N = a.N, T = a.T, B = a.B;
  const bool tracer = (a.trace != nullptr && blockIdx.x == 0);
#define PP_TRACE(i) do { if (tracer) a.trace[(i)] = clock64(); } while (0)

  unsigned char* kring = sm + C::OFF_KRING;
  unsigned char* vring = sm + C::OFF_VRING;
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
  float* ustage = reinterpret_cast<float*>(sm + C::OFF_UST);
  float* sred = reinterpret_cast<float*>(sm + C::OFF_SRED);
  auto qop = [&](int l) { return sm + C::OFF_LANE + l * C::LANE_BYTES; };
  auto slh_hi = [&](int l) { return sm + C::OFF_LANE + l * C::LANE_BYTES + C::QOP_BYTES; };
  auto own_of = [&](int l) { return reinterpret_cast<float*>(sm + C::OFF_LANE + l * C::LANE_BYTES + C::QOP_BYTES + KB * PITCH); };
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + C::OFF_BAR);
  uint64_t* k_full = bars;
  uint64_t* k_empty = k_full + NKS;
  uint64_t* v_full = k_empty + NKS;
  uint64_t* v_empty = v_full + NVS;
  uint64_t* lg_full = v_empty + NVS;   // [2] logits of a pair are in tensor memory
  uint64_t* lg_empty = lg_full + 2;    // [2] the softmax warps have read them
  uint64_t* w_full = lg_empty + 2;     // [2] softmax weights of a pair are in shared memory
  uint64_t* w_empty = w_full + 2;      // [2] the U products that read them are complete
  uint64_t* u_full = w_empty + 2;      // [2] U accumulator of an op is complete
  uint64_t* u_accfree = u_full + 2;    // [2] ... and has been drained
  uint64_t* xbar = u_accfree + 2;
  uint64_t* u_ready = xbar + 2;
  uint64_t* u_free = u_ready + 1;
  uint64_t* q_ready = u_free + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + C::OFF_TMEM);

  // ------------------------------------------------------------------ work assignment (NL lanes per cluster)
  // Images are dealt to clusters round-robin (newest first: the projection kernel left them in L2) and a cluster
  // deals its images to its lanes, so every cluster gets floor or ceil of B / #clusters images.
  const int ncimg = (B > cid) ? (B - cid + ncl - 1) / ncl : 0;
  // Lane l takes the cluster's images l, l + NL, ...; ops (one pass + one update of one lane) are streamed row by row:
  // row c holds op c of every lane that still has one, lanes in order (lane counts are non-increasing in l and
  // differ by at most one image).
  int nops[NL], row_start[NL + 1], row_width[NL + 1], seg_base[NL + 1];
  auto nimg_of = [&](int l) { return (ncimg > l) ? (ncimg - l + NL - 1) / NL : 0; };
#pragma unroll
  for (int l = 0; l < NL; ++l) nops[l] = nimg_of(l) * T;
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
  const int ntiles = (N + HT - 1) / HT;          // 64-token half tiles of an image
  const int TPC = (ntiles + CL - 1) / CL;
  const int tile0 = rank * TPC;
  const int TP = max(0, min(TPC, ntiles - tile0));  // half tiles of this CTA per pass
  const int NP = (TP + 1) / 2;                       // 128-token pairs
  const int total_ht = total_ops * TP;

  // One elected lane per ring: slot (j % NS) <- half tile j of the CTA's stream (ops in stream order).
  // L2 policy: tiles of iterations 0..T-2 are read again by the next pass (evict_last), the final pass's are dead
  // afterwards (evict_first).  For the first pass of an image the half tile `PFD` further on is prefetched into L2
  // at the rate tiles are consumed.
  const uint64_t pol_keep = l2_policy_evict_last(), pol_drop = l2_policy_evict_first();
  constexpr int PFD = 6;
  auto issue_half = [&](int j, bool is_v) {
    const int n = j / TP, tile = j - n * TP;
    int l, c;
    op_of(n, l, c);
    const int t = c % T;
    const int row0 = image_of(l, c / T) * N + (tile0 + tile) * HT;
    const uint64_t pol = (t == T - 1) ? pol_drop : pol_keep;
    const int s = j % (is_v ? NVS : NKS);
    unsigned char* dst = (is_v ? vring : kring) + (size_t)s * C::HT_BYTES;
    uint64_t* bar = is_v ? &v_full[s] : &k_full[s];
    const CUtensorMap* tm = is_v ? &tm_v : &tm_k;
    mbar_expect_tx(bar, (uint32_t)C::HT_BYTES);
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) tc::tma_load_2d_hint(dst + ch * C::CH_BYTES, tm, ch * 64, row0, bar, pol);
    const int jp = j + PFD;
    if (jp < total_ht) {
      const int pn = jp / TP, pt = jp - pn * TP;
      int pl, pc;
      op_of(pn, pl, pc);
      if (pc % T == 0) {
        const size_t off = ((size_t)image_of(pl, pc / T) * N + (size_t)(tile0 + pt) * HT) * D * 2;
        if (off + C::HT_BYTES <= (size_t)B * N * D * 2)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const char*>(is_v ? a.v : a.k) + off), "r"(C::HT_BYTES) : "memory");
      }
    }
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
    for (int s = 0; s < NKS; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); }
    for (int s = 0; s < NVS; ++s) { mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&lg_full[s], 1);
      mbar_init(&lg_empty[s], 4);
      mbar_init(&w_full[s], 4);
      mbar_init(&w_empty[s], 1);
      mbar_init(&u_full[s], 1);
      mbar_init(&u_accfree[s], 4);
    }
    mbar_init(&xbar[0], 1);
    mbar_init(&xbar[1], 1);
    mbar_init(u_ready, 4);
    mbar_init(u_free, 8);
    for (int l = 0; l < NL; ++l) mbar_init(&q_ready[l], 1);
    mbar_fence_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_v) : "memory");
  }
  if (warp == 7) tc::tmem_alloc<C::TMEM_COLS>(tmem_slot);
  // first half tiles: their HBM latency overlaps the weight load below (behind a barrier of the pass warps that
  // publishes the mbarrier init)
  if (warp < 8) {
    asm volatile("bar.sync 3, 256;" ::: "memory");
    if (lane == 0 && TP > 0) {
      if (warp == 4) for (int j = 0; j < NKS && j < total_ht; ++j) issue_half(j, false);
      if (warp == 6) for (int j = 0; j < NVS && j < total_ht; ++j) issue_half(j, true);
    }
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
    // staging rows of the padded slots are never written by the conversions; keep them finite.  The w tiles' slot rows
    // 8..15 and the q operands' rows K..7 stay zero for the whole kernel.
    for (int i = tid; i < (2 * KB * PITCHX) / 4; i += C::NT) reinterpret_cast<uint32_t*>(sm + C::OFF_ACT)[i] = 0u;
    for (int i = tid; i < (NL * C::LANE_BYTES) / 4; i += C::NT) reinterpret_cast<uint32_t*>(sm + C::OFF_LANE)[i] = 0u;
    for (int i = tid; i < (NWB * C::WP_BYTES) / 4; i += C::NT) reinterpret_cast<uint32_t*>(wtiles)[i] = 0u;
    fence_proxy_async();  // the zero rows are read by tcgen05.mma (async proxy)
  }
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tmem = *tmem_slot;
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
    if (warp < 4) {
      // ---- softmax warps: one token per thread.  Thread (warp w, lane) reads tensor-memory lane 32 w + lane:
      // lanes 0-15 hold rows 16 w .. 16 w + 15 of the pair's first half, lanes 16-31 the same rows of the second half.
      const int half = lane >> 4, r16 = lane & 15;
      const int tt = warp * 16 + r16;  // token inside its half
      const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
      uint32_t gp = 0;  // pairs handled so far (selects the logit / w buffers and their phases)
      for (int n = 0; n < total_ops; ++n) {
        int l, c;
        op_of(n, l, c);
        const int t = c % T, img = image_of(l, c / T);
        const bool last = (t == T - 1);
        float Sl[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) Sl[i] = 0.f;
        for (int p = 0; p < NP; ++p, ++gp) {
          const uint32_t lb = gp & 1, wb = gp % NWB;
          mbar_wait(&lg_full[lb], (gp >> 1) & 1);
          if (tid == 0 && p == 0 && n < 40) PP_TRACE(8 + n * 8);
          tc::fence_after();
          float x[8];
          tc::tmem_ld8(tlane + C::COL_LG + 16 * lb, x);
          tc::fence_before();
          __syncwarp();
          if (lane == 0) tc::arrive(&lg_empty[lb]);
          const int ht = 2 * p + half;
          const int tok = (tile0 + ht) * HT + tt;
          const bool tok_ok = (ht < TP) && (tok < N);
          float mx = -INFINITY;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            x[i] = (i < K) ? x[i] : -INFINITY;
            mx = fmaxf(mx, x[i]);
          }
          float sum = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            x[i] = ex2f(x[i] - mx);  // q carries log2(e)
            sum += x[i];
          }
          const float inv = __fdividef(1.f, sum);
#pragma unroll
          for (int i = 0; i < 8; ++i) x[i] *= inv;
          if (last && a.attn_out != nullptr && tok_ok) {
            float* ao = a.attn_out + ((size_t)img * N + tok) * K;
            if ((K & 1) == 0) {
#pragma unroll
              for (int i = 0; i < 8; i += 2)
                if (i < K) __stcs(reinterpret_cast<float2*>(ao + i), make_float2(x[i], x[i + 1]));
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (i < K) __stcs(ao + i, x[i]);
            }
          }
          // weights rounded to bf16 once; the same rounded values feed the numerator and the token sum
          if (gp >= (uint32_t)NWB) mbar_wait(&w_empty[wb], ((gp / NWB) - 1) & 1);
          unsigned char* wrow = wtiles + wb * C::WP_BYTES + half * C::WH_BYTES + (tt >> 3) * 256 + (tt & 7) * 2;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float wv = (tok_ok && i < K) ? x[i] + a.eps : 0.f;
            const __nv_bfloat16 wq = __float2bfloat16_rn(wv);
            *reinterpret_cast<__nv_bfloat16*>(wrow + i * 16) = wq;
            Sl[i] += __bfloat162float(wq);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) tc::arrive(&w_full[wb]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) Sl[i] = warp_sum(Sl[i]);
        // ---- drain the op's U accumulator into the staging buffer of the update engine
        float ua[8], ub[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { ua[i] = 0.f; ub[i] = 0.f; }
        if (NP > 0) {
          mbar_wait(&u_full[n & 1], (n >> 1) & 1);
          tc::fence_after();
          tc::tmem_ld8(tlane + C::COL_U + 32 * (n & 1), ua);
          if (C::HAS_B) tc::tmem_ld8(tlane + C::COL_U + 32 * (n & 1) + 16, ub);
          tc::fence_before();
          __syncwarp();
          if (lane == 0) tc::arrive(&u_accfree[n & 1]);
        }
        if (n > 0) mbar_wait(u_free, (uint32_t)((n - 1) & 1));  // the previous update has read sred / the U staging
        {
          // features: M = 128 block -> lane index 32 w + lane; M = 64 blocks -> lanes 0-15 of every quarter
          const int da = (C::MA == 128) ? warp * 32 + lane : warp * 16 + r16;
          const bool oka = (C::MA == 128) || lane < 16;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (i < KB && oka) ustage[i * UP + da] = ua[i];
          if (C::HAS_B && lane < 16) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (i < KB) ustage[i * UP + 128 + warp * 16 + r16] = ub[i];
          }
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (lane == i) sred[warp * 8 + i] = Sl[i];
        }
        if (tid == 0 && n < 40) PP_TRACE(8 + n * 8 + 1);
        __syncwarp();
        if (lane == 0) mbar_arrive(u_ready);
      }
    } else if (warp == 4 || warp == 6) {
      // ---- TMA producers: warp 4 the k ring, warp 6 the v ring
      if (lane == 0) {
        const bool is_v = (warp == 6);
        const int NS = is_v ? NVS : NKS;
        uint64_t* empty = is_v ? v_empty : k_empty;
        for (int j = NS; j < total_ht; ++j) {
          mbar_wait(&empty[j % NS], (uint32_t)(((j / NS) - 1) & 1));
          issue_half(j, is_v);
        }
      }
    } else if (warp == 5) {
      // ---- MMA issuer
      if (lane == 0) {
        constexpr uint32_t ID_LG = tc::idesc_bf16(64, 8);
        constexpr uint32_t ID_UA = tc::idesc_bf16(C::MA, 16, 1, 0);
        constexpr uint32_t ID_UB = tc::idesc_bf16(64, 16, 1, 0);
        uint32_t gp = 0, gu = 0;  // pairs whose logit / U products have been issued
        int jk = 0, jv = 0;       // half tiles consumed from the k / v rings
        for (int n = 0; n < total_ops; ++n) {
          int l, c;
          op_of(n, l, c);
          if (NP == 0) continue;
          mbar_wait(&q_ready[l], (uint32_t)(c & 1));
          fence_proxy_async();
          tc::fence_after();
          const uint32_t qa = smem_u32(qop(l));
          const uint32_t ucol = tmem + C::COL_U + 32 * (n & 1);
          auto logits = [&](int p) {
            const uint32_t lb = gp & 1;
            if (gp >= 2) {
              mbar_wait(&lg_empty[lb], ((gp >> 1) - 1) & 1);
              tc::fence_after();
            }
            for (int h = 0; h < 2 && 2 * p + h < TP; ++h) {
              const int s = jk % NKS;
              mbar_wait(&k_full[s], (uint32_t)((jk / NKS) & 1));
              tc::fence_after();
              const uint32_t ka = smem_u32(kring + (size_t)s * C::HT_BYTES);
              const uint32_t dcol = tmem + C::COL_LG + 16 * lb + ((uint32_t)(16 * h) << 16);
#pragma unroll
              for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  tc::mma_bf16(dcol, tc::smem_desc(ka + ch * C::CH_BYTES + ks * 32, 16, 1024, tc::SW_128),
                               tc::smem_desc(qa + (ch * 4 + ks) * 256, 128, 128, tc::SW_NONE), ID_LG, (ch | ks) != 0);
              tc::commit(&k_empty[s]);
              ++jk;
            }
            tc::commit(&lg_full[lb]);
            ++gp;
          };
          auto uprod = [&](int p) {
            const uint32_t wb = gu % NWB;
            mbar_wait(&w_full[wb], (gu / NWB) & 1);
            tc::fence_after();
            if (p == 0 && n >= 2) {  // the accumulator was last used by op n - 2
              mbar_wait(&u_accfree[n & 1], (uint32_t)(((n >> 1) - 1) & 1));
              tc::fence_after();
            }
            for (int h = 0; h < 2 && 2 * p + h < TP; ++h) {
              const int s = jv % NVS;
              mbar_wait(&v_full[s], (uint32_t)((jv / NVS) & 1));
              tc::fence_after();
              const uint32_t va = smem_u32(vring + (size_t)s * C::HT_BYTES);
              const uint32_t wa = smem_u32(wtiles + wb * C::WP_BYTES + h * C::WH_BYTES);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {  // 16 tokens per step: two 8-token groups 1024 B apart, next step +2048 B
                const uint32_t acc = (p | h | ks) != 0;
                const uint64_t db = tc::smem_desc(wa + ks * 512, 256, 128, tc::SW_NONE);
                tc::mma_bf16(ucol, tc::smem_desc(va + ks * 2048, C::CH_BYTES, 1024, tc::SW_128), db, ID_UA, acc);
                if (C::HAS_B)
                  tc::mma_bf16(ucol + 16, tc::smem_desc(va + 2 * C::CH_BYTES + ks * 2048, C::CH_BYTES, 1024, tc::SW_128), db, ID_UB, acc);
              }
              tc::commit(&v_empty[s]);
              ++jv;
            }
            tc::commit(&w_empty[wb]);
            ++gu;
          };
          for (int p = 0; p < NP; ++p) {
            logits(p);
            if (p > 0) uprod(p - 1);
          }
          uprod(NP - 1);
          tc::commit(&u_full[n & 1]);
        }
      }
    }
    __syncwarp();
  } else {
    // ==================================================================================== UPDATE ENGINE
    // Every exchange round is  push -> wait -> tensor-core product -> one barrier -> epilogue (= next push):
    // all-gather payloads travel as bf16 straight into the receivers' MMA staging buffers (act[parity], slh[lane],
    // the lane's q operand); the LayerNorms are folded into the products that follow them (statistics from the received rows).
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
    // same for the queries, into the K-major UMMA operand [D/8][8 slots][8 features] of every CTA
    auto quad_push_q = [&](uint32_t r, float val, int i, unsigned char* dst) {
      const int qb = lane & ~3;
      const float v0 = __shfl_sync(FULL, val, qb), v1 = __shfl_sync(FULL, val, qb + 1);
      const float v2 = __shfl_sync(FULL, val, qb + 2), v3 = __shfl_sync(FULL, val, qb + 3);
      if (i < K * DS) {
        const int slot = i / DS, f = rank * DS + ((i % DS) & ~3);
        const uint32_t lbuf = smem_u32(dst) + (uint32_t)((f >> 3) * 128 + slot * 16 + (f & 7) * 2);
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
    // q = W_q LN(slots) of lane l for the CTA's slice (slots = slh[l], bf16), all-gathered (times log2 e) into the lane's q operand
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
        quad_push_q(round, val, i, qop(l));
      }
      xwait(round);
      ++round;
      if (utid == 0) {  // the MMA issuer may read lane l's q operand (tcgen05.mma reads it through the async proxy)
        fence_proxy_async();
        mbar_arrive(&q_ready[l]);
      }
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
          const float4 val = *reinterpret_cast<const float4*>(ustage + slot * UP + d);
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
  tc::fence_before();
  __syncthreads();
  if (warp == 7) tc::tmem_dealloc<C::TMEM_COLS>(tmem);
  cluster.sync();  // no CTA leaves while a peer may still address its shared memory
}


template <int D, int H, int CL, int NL, int KB, int NKS, int NVS, int NWB>
static int launch_umma(const IterFwdArgs& a, cudaStream_t stream) {
  using C = Cfg<D, H, CL, NL, KB, NKS, NVS, NWB>;
  auto kern = sa_iter_fwd_umma_kernel<D, H, CL, NL, KB, NKS, NVS, NWB>;
  static_assert(C::SMEM_BYTES <= 227 * 1024, "shared memory budget");
  CUtensorMap tm_k, tm_v;
  const uint64_t rows = (uint64_t)a.B * a.N;
  if (!tc::make_map_bf16_sw128(&tm_k, a.k, D, rows, (uint64_t)D * 2, C::HT) ||
      !tc::make_map_bf16_sw128(&tm_v, a.v, D, rows, (uint64_t)D * 2, C::HT)) {
    set_error("sa_iter_fwd(tcgen05): cuTensorMapEncodeTiled failed");
    return OCRL_E_LAUNCH;
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
  static int max_clusters = -1;  // per instantiation; one device per process
  if (max_clusters < 0) {
    OCRL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    cfg.gridDim = dim3(CL * 148);
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
    if (e != cudaSuccess || n <= 0) {
      (void)cudaGetLastError();
      set_error("sa_iter_fwd(tcgen05): cluster size %d with %d B of shared memory cannot be scheduled", CL, C::SMEM_BYTES);
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
  return OCRL_OK;
}

}  // namespace umma

// Returns OCRL_E_SHAPE (without launching) for shapes this kernel does not cover.
int sa_iter_fwd_umma_dispatch(const IterFwdArgs& a, cudaStream_t s) {
  if (a.K > 8) {
    set_error("sa_iter_fwd(tcgen05): K <= 8");
    return OCRL_E_SHAPE;
  }
  if (a.D == 192 && a.H == 192) {
    if (a.K <= 6) {
      if (a.lanes == 3) return umma::launch_umma<192, 192, 8, 3, 6, 2, 2, 1>(a, s);
      return umma::launch_umma<192, 192, 8, 2, 6, 2, 2, 2>(a, s);
    }
    return umma::launch_umma<192, 192, 8, 2, 8, 2, 2, 1>(a, s);
  }
  if (a.D == 64 && a.H == 128) {  // the "Slot-Attention (small)" configuration
    if (a.B >= 48 && a.lanes != 3) return umma::launch_umma<64, 128, 4, 2, 8, 4, 4, 2>(a, s);
    return umma::launch_umma<64, 128, 8, 3, 8, 4, 4, 2>(a, s);
  }
  set_error("sa_iter_fwd(tcgen05): D=%d H=%d not instantiated", a.D, a.H);
  return OCRL_E_SHAPE;
}

}  // namespace ocrl

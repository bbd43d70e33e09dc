// tcgen05 form of the fused T-iteration loop (slot_attn.py:64-102) for bf16 k/v: the token pass runs on the
// 5th-generation tensor cores with its accumulators in tensor memory.
//
// Structure as in sa_iter_fwd_pipe.cu: a cluster of CL CTAs works on NL images at a time ("lanes"); every CTA
// streams its share of the tokens (PASS engine, warps 0-7) while warps 8-15 run the slot update of another lane
// (UPDATE engine: reduce-scatter over the cluster, GRU, residual MLP, next q; weights stationary in shared memory).
// The slot-update weights live in TENSOR MEMORY (A operand of tcgen05.mma, 2 x 96 columns of bf16 pairs), which
// leaves the shared memory to the k/v ring: seven 24 KB slots instead of four.  The pass engine is
//   * warp 4 / warp 6, one lane each: TMA producers of the k ring and the v ring (64-token half tiles,
//     [D/64][64 rows][128 B] with the 128-byte swizzle = the canonical UMMA layouts: K-major for k, MN-major for v);
//   * warps 5 and 7, one lane each: issue  logits[64 x 8] = k_half . q^T  (M = 64; the two halves of a 128-token pair land in
//     the lower / upper 16 lanes of every 32-lane quarter of one tensor-memory buffer) and
//     U^T[D x 16] += v_half^T . w  (M = 128 for features 0-127, M = 64 for 128-191), tcgen05.commit frees the ring
//     slots and signals the softmax warps;
//   * warps 0-3, one token per thread: tcgen05.ld of the token's logits, K-way softmax in registers (no shuffles),
//     attn_vis store in the last iteration, w = a + eps rounded to bf16 into the B-operand tile of the U product.
// The update engine's matrix-vector products are tcgen05.mma as well: D[weight rows x slots] = W (tensor memory) x
// activations (shared memory, K-major [features/8][8 slots][8] bf16 -- the layout the all-gathers write directly).
#include "pc_common.cuh"
#include "umma_common.cuh"

namespace ocrl {
namespace umma {

using namespace pc;

#ifndef TRACE_OP
#define TRACE_OP 2
#endif
// The clock64 phase stamps (ocrl_sa_launch_opts.trace, scripts/trace_umma.py) are compiled in only on request
//   OCRL_NVCC_FLAGS=-DOCRL_UMMA_TRACE=1 python -m ocrl_b200.build
// -- the factored kernel runs 640 threads at the 96-register limit and pays for every live value.
#ifndef OCRL_UMMA_TRACE
#define OCRL_UMMA_TRACE 0
#endif

template <int D_, int H_, int CL_, int NL_, int KB_, int NKS_, int NVS_, int NWB_, int NUS_, int DEFER_ = 0, int F_ = 0, int BT_ = 0>
struct Cfg {
  // NL image lanes per cluster; KB rows in the slot-indexed buffers; NKS / NVS ring slots for k / v half tiles;
  // NWB buffers for the softmax weights of a 128-token pair
  static constexpr int D = D_, H = H_, CL = CL_, NL = NL_, KB = KB_, NKS = NKS_, NVS = NVS_, NWB = NWB_;
  static constexpr int KS = KB > 8 ? 16 : 8;  // slot columns of the tensor-core operands and accumulators (K <= KS)
  // NUS update streams: the eight update warps are split into NUS groups that run the slot updates of different ops
  // concurrently (op n -> stream n % NUS); an update is a chain of exchange rounds, bound by latency, not by work
  // (up to two streams share the eight update warps of a 512-thread CTA; more streams bring four warps each -- a
  // stream needs one warp per tensor-memory lane quarter to read its accumulators)
  static constexpr int NUS = NUS_, UWT = NUS_ <= 2 ? 8 : 4 * NUS_, UW = UWT / NUS, UT = 32 * UW;
  static constexpr bool DEFER = DEFER_ != 0;  // drain an op's U accumulator after the next op's first softmax
  static_assert(NUS >= 1 && NUS <= 4 && NL <= 6, "update streams, lanes");
  // F > 0: the FACTORED form of the pass.  k = s W_k x^ and v = W_v x^ are rank-F functions of the normalised tokens
  // x^ [N, F] (slot_attn.py:54-61, no bias), so  k . q = x^ . (s W_k^T q)  and  sum_n w_n v_n = W_v (sum_n w_n x^_n):
  // the kernel streams x^ (F * 2 bytes per token instead of 4 D) through ONE ring -- the same tile is the K-major
  // operand of the logits and the MN-major operand of the weighted sum -- and the two projections are folded into the
  // update weights (W_q'' = s W_k^T W_q diag(gamma), W_ih'' = W_ih W_v; prepared once per parameter version).
  static constexpr int F = F_;
  static constexpr bool XH = F_ != 0;
  static constexpr int FW = XH ? F_ : D_;  // feature width of the streamed tiles
  static constexpr int FS = FW / CL_;      // ... and this CTA's slice of it in the exchanges
  static_assert(!XH || (NKS_ == NVS_ && FW % (4 * CL_) == 0), "factored pass: one ring, four features per push");
  // The factored pass is short (the softmax warps are its critical path), so an op's accumulator is read out of tensor
  // memory by the update stream that consumes it, not by the softmax warps
  static constexpr bool UDRAIN = XH;
  // BT: 256-token steps of the factored pass.  One ring tile, one barrier round and one softmax step cover 256 tokens
  // (two tokens per softmax thread, logits as two M = 128 products side by side in tensor memory): the per-step costs of
  // the three pass roles (mbarrier waits ~100 cycles each, commits, fences) are paid once per 256 instead of per 128 tokens
  static constexpr bool BT = BT_ != 0;
  static_assert(!BT || XH, "256-token steps belong to the factored pass");
  static constexpr int NT = 256 + 32 * UWT, HT = BT ? 256 : 64, NCH = FW / 64;
  static constexpr int LX = D > H ? D : H;
  static constexpr int CH_BYTES = HT * 128, HT_BYTES = NCH * CH_BYTES;  // one 64-wide feature chunk / one half tile
  static constexpr int WH_BYTES = 2048, WP_BYTES = (BT ? 4 : 2) * WH_BYTES;  // w tile of 64 tokens ([8][16 slots][8] bf16) / of a step
  static constexpr int OPD_BYTES = D * 2 * KS, OPX_BYTES = LX * 2 * KS;  // activation operands [features/8][KS slots][8] bf16
  static constexpr int DS = D / CL, HS = H / CL;
  // weight blocks in tensor memory (row = lane): X = W_ih (3 DS rows) | W1' (HS) | W2 (DS);  Y = W_hh (3 DS) | Wq' (DS)
  static constexpr int RX = 3 * DS + HS + DS, RY = 3 * DS + DS;
  static constexpr int WPB = LX / 2;  // 32-bit words (bf16 pairs) per weight row in the prepared copy
  static constexpr int UP = FW + 4;
  static constexpr int MA = NCH >= 2 ? 128 : 64;   // U product, features [0, 128) (or all 64)
  static constexpr bool HAS_B = NCH == 3;          // second U product, features [128, 192)
  static_assert(D % 64 == 0 && FW % 64 == 0 && H % 64 == 0 && DS % 4 == 0 && HS % 4 == 0 && NCH <= 3 && RX <= 128 && RY <= 128, "shape");
  // Tensor memory columns: two logit buffers, two U accumulators (features 0-127 | 128-191), the update engine's two
  // accumulators, the weight blocks (a column holds two bf16: K features take K / 2 columns).  (No need to spread a
  // product over several accumulators: back-to-back MMAs into the same columns issue at full rate, scripts/umma_time.cu.)
  static constexpr uint32_t COL_LG = 0, LG_STRIDE = BT ? 32 : 16, COL_U = 2 * LG_STRIDE, U_STRIDE = 32, U_B = 16;
  static constexpr uint32_t COL_GI = COL_U + 64, COL_GH = COL_GI + 16;
  static constexpr uint32_t COL_WX = COL_GH + 16, COL_WY = COL_WX + LX / 2, TMEM_COLS = 512;
  static constexpr uint32_t COL_GI2 = COL_WY + D / 2, COL_GH2 = COL_GI2 + 16;  // accumulators of the update streams 1, 2, ...: 32 columns each
  static_assert(COL_GI2 + 32 * (NUS - 1) <= 512, "tensor memory");

  static constexpr int OFF_KRING = 0;
  static constexpr int OFF_VRING = OFF_KRING + NKS * HT_BYTES;
  static constexpr int OFF_WT = OFF_VRING + (XH ? 0 : NVS * HT_BYTES);  // (factored pass: one ring)
  static constexpr int OFF_BIAS = OFF_WT + NWB * WP_BYTES;  // (the LayerNorm affine parameters are folded into W1', Wq')
  // fp32 constants: b_ih[3DS] b_hh[3DS] b1'[HS] b2[DS] c1[HS] cq[DS] bq'[DS] + LayerNorm stats mean[8] rstd[8]
  static constexpr int NCONST = 9 * DS + 2 * HS + 32 * NUS;
  static constexpr int XBUF_BYTES = KB * D * 4 + 4 * KS * CL;      // R1 receive buffer (fp32 partial sums), single
  static constexpr int OFF_XBUF = (OFF_BIAS + NCONST * 4 + 15) & ~15;
  static constexpr int OFF_ACT = (OFF_XBUF + NUS * XBUF_BYTES + 127) & ~127;  // bf16 all-gather targets (operands), by round parity
  static constexpr int OFF_UST = OFF_ACT + NUS * 2 * OPX_BYTES;    // U staging (pass -> update hand-off), one per stream
  static constexpr int UST_BYTES = (KB * UP * 4 + 127) & ~127;
  static constexpr int OFF_LANE = (OFF_UST + NUS * UST_BYTES + 127) & ~127;  // per lane: q operand, slots operand, own slice (fp32)
  static constexpr int LANE_BYTES = (2 * OPD_BYTES + KB * DS * 4 + 127) & ~127;
  static constexpr int OFF_P = (OFF_LANE + NL * LANE_BYTES + 15) & ~15;   // product outputs, fp32 [2][128 rows][8 slots]
  static constexpr int OFF_SRED = OFF_P + NUS * 2 * 128 * 4 * KS;  // [4][KS] token sums of the softmax warps, per stream
  // k_full k_empty [NKS] v_full v_empty [NVS] lg_full lg_empty w_full w_empty u_full u_accfree [2 each]
  // per stream: xbar[2] u_ready u_free ubar[2];  q_ready[NL]
  static constexpr int OFF_BAR = OFF_SRED + NUS * 16 * KS;
  static constexpr int NBAR = 2 * NKS + 2 * NVS + 12 + 6 * NUS + NL;
  static constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
  // op table (factored pass): op n -> lane | iteration << 4 | image round << 12.  Finding an op's lane / iteration /
  // image takes integer divisions by run-time values in every role of the kernel (~2 k cycles per op in the softmax
  // warps alone: a third of a factored pass); the table is filled once, in parallel, during the setup
  static constexpr int OPT_MAX = XH ? 2048 : 0;
  static constexpr int OFF_OPTAB = OFF_TMEM + 32;  // (tensor-memory base, NL <= 6 query counters)
  static constexpr int SMEM_BYTES = OFF_OPTAB + 4 * OPT_MAX;
};

__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// barrier of one update stream (ids 1, 2)
__device__ __forceinline__ void upd_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

template <int D, int H, int CL, int NL, int KB, int NKS, int NVS, int NWB, int NUS, int DEFER, int F, int BT>
__global__ void __launch_bounds__((NUS <= 2 ? 512 : 256 + 128 * NUS), 1)
sa_iter_fwd_umma_kernel(const IterFwdArgs a, const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v) {
  using C = Cfg<D, H, CL, NL, KB, NKS, NVS, NWB, NUS, DEFER, F, BT>;
  constexpr int DS = C::DS, HS = C::HS, UP = C::UP, HT = C::HT, NCH = C::NCH, UW = C::UW, UT = C::UT, KS = C::KS;
  constexpr int FW = C::FW, FS = C::FS;
  constexpr bool XH = C::XH;
  constexpr float LOG2E = 1.4426950408889634f;

  extern __shared__ __align__(1024) unsigned char sm[];  // swizzled TMA tiles need 1024-byte alignment
  if ((smem_u32(sm) & 1023u) != 0) __trap();
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / CL, ncl = gridDim.x / CL;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = a.K, N = a.N, T = a.T, B = a.B;
  const bool tracer = OCRL_UMMA_TRACE && (a.trace != nullptr && blockIdx.x == 0);
#define PP_TRACE(i) do { if (tracer) a.trace[(i)] = clock64(); } while (0)
  // kernel start / end of setup / kernel end of the first and the last cluster's CTA 0: trace[1..3], trace[4..6]
  const bool tracer2 = OCRL_UMMA_TRACE && (a.trace != nullptr && blockIdx.x == gridDim.x - CL && tid == 0);
  if (tracer && tid == 0) a.trace[1] = clock64();
  if (tracer2) a.trace[4] = clock64();

  unsigned char* kring = sm + C::OFF_KRING;
  unsigned char* vring = XH ? kring : sm + C::OFF_VRING;  // factored pass: one ring, two readers
  unsigned char* wtiles = sm + C::OFF_WT;
  float* s_bias = reinterpret_cast<float*>(sm + C::OFF_BIAS);
  const float* s_bih = s_bias;
  const float* s_bhh = s_bias + 3 * DS;
  float* s_b1f = s_bias + 6 * DS;            // b1 + W1 . beta_mlp
  const float* s_b2 = s_bias + 6 * DS + HS;
  float* s_c1 = s_bias + 7 * DS + HS;        // row sums of the folded bf16 W1'
  float* s_cq = s_bias + 7 * DS + 2 * HS;    // row sums of the folded bf16 Wq'
  float* s_bqf = s_bias + 8 * DS + 2 * HS;   // Wq . beta_slots
  // update stream of this warp (warps 8..15; the pass warps address the per-stream hand-off buffers by op)
  const int us = (warp >= 8) ? (warp - 8) / UW : 0;
  float* s_mean = s_bias + 9 * DS + 2 * HS + 32 * us;  // LayerNorm statistics of the round in flight
  float* s_rstd = s_mean + 16;
  unsigned char* xbuf = sm + C::OFF_XBUF + us * C::XBUF_BYTES;
  auto act = [&](uint32_t r) { return sm + C::OFF_ACT + (2 * us + (r & 1)) * C::OPX_BYTES; };  // operand [LX/8][8 slots][8] bf16
  float* P_GI = reinterpret_cast<float*>(sm + C::OFF_P) + us * 2 * 128 * KS;  // [128 rows of block X][KS slots]
  float* P_GH = P_GI + 128 * KS;                                              // [128 rows of block Y][KS slots]
  auto ustage_of = [&](int s) { return reinterpret_cast<float*>(sm + C::OFF_UST + s * C::UST_BYTES); };
  auto sred_of = [&](int s) { return reinterpret_cast<float*>(sm + C::OFF_SRED + s * 16 * KS); };
  auto qop = [&](int l) { return sm + C::OFF_LANE + l * C::LANE_BYTES; };
  auto slh_hi = [&](int l) { return sm + C::OFF_LANE + l * C::LANE_BYTES + C::OPD_BYTES; };  // the lane's slots, operand layout
  auto own_of = [&](int l) { return reinterpret_cast<float*>(sm + C::OFF_LANE + l * C::LANE_BYTES + 2 * C::OPD_BYTES); };
  // byte offset of (slot, feature f) inside an operand buffer
  auto opnd_off = [](int slot, int f) { return (f >> 3) * (16 * KS) + slot * 16 + (f & 7) * 2; };
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + C::OFF_BAR);
  uint64_t* k_full = bars;
  uint64_t* k_empty = k_full + NKS;
  uint64_t* v_full = XH ? k_full : k_empty + NKS;
  uint64_t* v_empty = XH ? k_empty : k_empty + NKS + NVS;
  uint64_t* lg_full = k_empty + NKS + 2 * NVS;   // [2] logits of a pair are in tensor memory
  uint64_t* lg_empty = lg_full + 2;    // [2] the softmax warps have read them
  uint64_t* w_full = lg_empty + 2;     // [2] softmax weights of a pair are in shared memory
  uint64_t* w_empty = w_full + 2;      // [2] the U products that read them are complete
  uint64_t* u_full = w_empty + 2;      // [2] U accumulator of an op is complete
  uint64_t* u_accfree = u_full + 2;    // [2] ... and has been drained
  uint64_t* sbars = u_accfree + 2;     // per update stream: xbar[2] u_ready u_free ubar[2]
  uint64_t* xbar = sbars + 6 * us;
  auto u_ready_of = [&](int s) { return sbars + 6 * s + 2; };
  auto u_free_of = [&](int s) { return sbars + 6 * s + 3; };
  uint64_t* ubar = sbars + 6 * us + 4;  // [2] a product of the stream is complete ([1]: W_hh h, which overlaps others)
  uint64_t* q_ready = sbars + 6 * NUS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + C::OFF_TMEM);
  uint32_t* q_count = tmem_slot + 1;  // [NL <= 6] queries published per lane (read by the update streams)

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
  auto op_of_div = [&](int n, int& l, int& c) {
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
  const uint32_t* optab = reinterpret_cast<const uint32_t*>(sm + C::OFF_OPTAB);
  // (the launcher splits a batch whose clusters would hold more ops than the table: with the table the segment arrays
  // above are dead after the setup -- in a kernel that runs 640 threads at the 96-register limit that is what matters)
  constexpr bool use_tab = C::OPT_MAX > 0;
  if (use_tab) {
    uint32_t* tab = reinterpret_cast<uint32_t*>(sm + C::OFF_OPTAB);
    for (int n = threadIdx.x; n < total_ops; n += C::NT) {
      int l, c;
      op_of_div(n, l, c);
      tab[n] = (uint32_t)l | ((uint32_t)(c % T) << 4) | ((uint32_t)(c / T) << 12);
    }
    __syncthreads();
  }
  // op n -> lane l, op index c of the lane, iteration t = c % T, image round m = c / T
  auto op_of4 = [&](int n, int& l, int& c, int& t, int& m) {
    if (use_tab) {
      const uint32_t e = optab[n];
      l = (int)(e & 15u); t = (int)((e >> 4) & 255u); m = (int)(e >> 12);
      c = m * T + t;
    } else {
      op_of_div(n, l, c);
      t = c % T; m = c / T;
    }
  };
  auto op_of = [&](int n, int& l, int& c) {
    int t, m;
    op_of4(n, l, c, t, m);
  };
  auto image_of = [&](int l, int m) { return B - 1 - (cid + (NL * m + l) * ncl); };
  const int ntiles = (N + HT - 1) / HT;          // 64-token half tiles of an image
  const int TPC = (ntiles + CL - 1) / CL;
  const int tile0 = rank * TPC;
  const int TP = max(0, min(TPC, ntiles - tile0));  // half tiles of this CTA per pass
  const int NP = C::BT ? TP : (TP + 1) / 2;          // steps per pass: 128-token pairs (256-token tiles with BT)
  const int total_ht = total_ops * TP;

  // TMA producers (warp 4: k ring, warp 6: v ring).  The whole warp walks the half-tile stream of the CTA (ops in
  // stream order) so that addresses and coordinates stay warp-uniform; one elected lane issues.  All per-tile state is
  // kept incrementally -- an integer division costs the issuing warp more than the three TMA instructions.
  // L2 policy: tiles of iterations 0..T-2 are read again by the next pass (evict_last), the final pass's are dead
  // afterwards (evict_first).  For the first pass of an image the half tile `PFD` further on is prefetched into L2
  // at the rate tiles are consumed.
  const uint64_t pol_keep = l2_policy_evict_last(), pol_drop = l2_policy_evict_first();
  constexpr int PFD = 6;
  struct OpInfo { int row_base; int cold; uint64_t pol; };
  auto op_info = [&](int n) {
    OpInfo o;
    o.row_base = 0; o.cold = 0; o.pol = pol_keep;
    if (n < total_ops) {
      int l, c, t, m;
      op_of4(n, l, c, t, m);
      o.row_base = image_of(l, m) * N + tile0 * HT;
      o.cold = (t == 0);
      o.pol = (t == T - 1) ? pol_drop : pol_keep;
    }
    return o;
  };
  const bool prod_v = (warp == 6);
  const bool is_prod = (warp == 4) || (warp == 6 && !XH);
  const int PNS = prod_v ? NVS : NKS;
  int pj = 0, p_n = 0, p_tile = 0, p_slot = 0, p_round = 0;  // next half tile to issue, its ring slot and round
  int f_n = 0, f_tile = 0;                                   // the half tile PFD further on (L2 prefetch)
  OpInfo p_cur = op_info(0), f_cur = p_cur;
  if (is_prod && TP > 0) {
    f_n = PFD / TP;
    f_tile = PFD - f_n * TP;
    f_cur = op_info(f_n);
  }
  auto produce = [&]() {
    unsigned char* dst = (prod_v ? vring : kring) + (size_t)p_slot * C::HT_BYTES;
    uint64_t* full = (prod_v ? v_full : k_full) + p_slot;
    const bool tt = tracer && lane == 0 && p_n == TRACE_OP && p_tile < 16;
    if (tt) a.trace[400 + p_tile * 4 + (prod_v ? 2 : 0)] = clock64();
    if (p_round > 0) mbar_wait((prod_v ? v_empty : k_empty) + p_slot, (uint32_t)((p_round - 1) & 1));
    if (tt) a.trace[400 + p_tile * 4 + (prod_v ? 3 : 1)] = clock64();
    const CUtensorMap* tm = prod_v ? &tm_v : &tm_k;
    const int row0 = p_cur.row_base + p_tile * HT;
    const bool leader = tc::elect_one();
    if (leader) {
      mbar_expect_tx(full, (uint32_t)C::HT_BYTES);
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) tc::tma_load_2d_hint(dst + ch * C::CH_BYTES, tm, ch * 64, row0, full, p_cur.pol);
      if (f_n < total_ops && f_cur.cold) {
        const size_t off = (size_t)(f_cur.row_base + f_tile * HT) * (FW * 2);
        if (off + C::HT_BYTES <= (size_t)B * N * FW * 2)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const char*>(prod_v ? a.v : a.k) + off), "r"(C::HT_BYTES) : "memory");
      }
    }
    __syncwarp();
    ++pj;
    if (++p_slot == PNS) { p_slot = 0; ++p_round; }
    if (++p_tile == TP) { p_tile = 0; ++p_n; p_cur = op_info(p_n); }
    if (++f_tile == TP) { f_tile = 0; ++f_n; f_cur = op_info(f_n); }
  };

  // ------------------------------------------------------------------ one-time setup
  // The slots of lane 0's first image are fetched now and written to shared memory at the end of the setup: their global
  // round trip hides under the weight load instead of standing in front of the first query
  constexpr int PRE2 = (KS * (D / 2) + 255) / 256, PRE1 = (KS * DS + 255) / 256;
  float2 pre2[PRE2];
  float pre1[PRE1];
  if (tid >= 256 && tid < 512 && nops[0] > 0) {
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
    for (int s = 0; s < NKS; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], XH ? 2 : 1); }  // (factored: both issuers release a slot)
    if (!XH)
      for (int s = 0; s < NVS; ++s) { mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&lg_full[s], 1);
      mbar_init(&lg_empty[s], 4);
      mbar_init(&w_full[s], 4);
      mbar_init(&w_empty[s], 1);
      mbar_init(&u_full[s], 1);
      mbar_init(&u_accfree[s], 4);
    }
    for (int s = 0; s < NUS; ++s) {
      mbar_init(sbars + 6 * s + 0, 1);
      mbar_init(sbars + 6 * s + 1, 1);
      mbar_init(u_ready_of(s), 4);
      mbar_init(u_free_of(s), UW);
      mbar_init(sbars + 6 * s + 4, 1);
      mbar_init(sbars + 6 * s + 5, 1);
    }
    for (int l = 0; l < NL; ++l) { mbar_init(&q_ready[l], 1); q_count[l] = 0u; }
    mbar_fence_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_v) : "memory");
  }
  if (warp == 7) tc::tmem_alloc<C::TMEM_COLS>(tmem_slot);
  // first half tiles: their HBM latency overlaps the weight load below (behind a barrier of the pass warps that
  // publishes the mbarrier init)
  if (warp < 8) {
    asm volatile("bar.sync 8, 256;" ::: "memory");
    if (is_prod)
      while (pj < PNS && pj < total_ht) produce();
  }
  tc::fence_before();
  __syncthreads();  // tensor memory is allocated
  tc::fence_after();
  const uint32_t tmem = *tmem_slot;
  {
    // The CTA's weight slices -> tensor memory, one weight row per thread (= lane): warps 0-3 fill block X, warps 8-11
    // block Y, from the bf16 pairs umma_prep_kernel laid out column-major ([word][row]: coalesced 4-byte loads).
    const bool fill_x = warp < 4, fill_y = (warp >= 8 && warp < 12);
    if (fill_x || fill_y) {
      const int r = (warp & 3) * 32 + lane;
      const uint32_t* wsrc = a.wprep + ((size_t)(rank * 2 + (fill_x ? 0 : 1)) * C::WPB) * 128 + r;
      const int nwords = fill_x ? C::LX / 2 : D / 2;
      const uint32_t tcol = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (fill_x ? C::COL_WX : C::COL_WY);
      for (int w0 = 0; w0 < nwords; w0 += 32) {  // 32 loads in flight, then four stores of one k step (8 columns) each
        uint32_t wv[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) wv[j] = __ldg(wsrc + (size_t)(w0 + j) * 128);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tcol + (uint32_t)(w0 + 8 * q)),
                       "r"(wv[8 * q]), "r"(wv[8 * q + 1]), "r"(wv[8 * q + 2]), "r"(wv[8 * q + 3]), "r"(wv[8 * q + 4]),
                       "r"(wv[8 * q + 5]), "r"(wv[8 * q + 6]), "r"(wv[8 * q + 7])
                       : "memory");
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    {  // constants of the folded LayerNorms: c1[HS] b1'[HS] cq[DS] bq'[DS]
      const float* cs = a.wprep_consts + (size_t)rank * (2 * HS + 2 * DS);
      for (int i = tid; i < HS; i += C::NT) { s_c1[i] = __ldg(cs + i); s_b1f[i] = __ldg(cs + HS + i); }
      for (int i = tid; i < DS; i += C::NT) { s_cq[i] = __ldg(cs + 2 * HS + i); s_bqf[i] = __ldg(cs + 2 * HS + DS + i); }
    }
    for (int i = tid; i < 3 * DS; i += C::NT) {
      const int gate = i / DS, dl = i % DS;
      s_bias[i] = a.w.b_ih[gate * D + rank * DS + dl];
      s_bias[3 * DS + i] = a.w.b_hh[gate * D + rank * DS + dl];
    }
    for (int i = tid; i < DS; i += C::NT) s_bias[6 * DS + HS + i] = a.w.b2[rank * DS + i];
    // operand rows of the padded slots are never written; keep them zero.  The w tiles' slot rows 8..15 and the
    // operands' rows K..7 stay zero for the whole kernel.
    for (int i = tid; i < (NUS * 2 * C::OPX_BYTES) / 4; i += C::NT) reinterpret_cast<uint32_t*>(sm + C::OFF_ACT)[i] = 0u;
    for (int i = tid; i < (NL * C::LANE_BYTES) / 4; i += C::NT) reinterpret_cast<uint32_t*>(sm + C::OFF_LANE)[i] = 0u;
    for (int i = tid; i < (NWB * C::WP_BYTES) / 4; i += C::NT) reinterpret_cast<uint32_t*>(wtiles)[i] = 0u;
    fence_proxy_async();  // the zero rows are read by tcgen05.mma (async proxy)
  }
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  if (tid >= 256 && tid < 512 && nops[0] > 0) {  // after the zero fill of the lane buffers; the cluster barrier below publishes it
    unsigned char* dst = slh_hi(0);
    float* own = own_of(0);
#pragma unroll
    for (int u = 0; u < PRE2; ++u) {
      const int i = tid - 256 + 256 * u;
      if (i < K * (D / 2)) *reinterpret_cast<uint32_t*>(dst + opnd_off(i / (D / 2), 2 * (i % (D / 2)))) = pack_bf16x2(pre2[u].x, pre2[u].y);
    }
#pragma unroll
    for (int u = 0; u < PRE1; ++u) {
      const int i = tid - 256 + 256 * u;
      if (i < K * DS) own[i] = pre1[u];
    }
    fence_proxy_async();  // the slots are an MMA operand
  }
  cluster.sync();  // every CTA's barriers are initialised before any peer signals them
  if (tid == 0) PP_TRACE(0);
  if (tracer2) a.trace[5] = clock64();

  if (warp < 8) {
    // ==================================================================================== PASS ENGINE
    if (warp < 4) {
      // ---- softmax warps: one token per thread.  Thread (warp w, lane) reads tensor-memory lane 32 w + lane:
      // lanes 0-15 hold rows 16 w .. 16 w + 15 of the pair's first half, lanes 16-31 the same rows of the second half.
      const int half = lane >> 4, r16 = lane & 15;
      const int tih = warp * 16 + r16;  // token inside its half
      const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
      uint32_t gp = 0;  // pairs handled so far (selects the logit / w buffers and their phases)
      // Drain of an op's U accumulator into the staging buffer of the update engine.  It is DEFERRED: the softmax of the
      // next op's first pair runs first, so the U issuer always has softmax weights waiting when the previous op's
      // products retire and the tensor pipe does not idle across op boundaries.
      auto drain = [&](int n, const float (&S)[KS]) {
        if (C::UDRAIN) {  // only the token sums; the update stream fetches the accumulator itself
          const int sx = n % NUS;
          float* sred = sred_of(sx);
          if (tracer && tid == 0 && n == TRACE_OP) a.trace[364 + 11] = clock64();
          if (n >= NUS) mbar_wait(u_free_of(sx), (uint32_t)((n / NUS - 1) & 1));
          if (tracer && tid == 0 && n == TRACE_OP) a.trace[376 + 11] = clock64();
          {  // lane i keeps slot i's sum (a select chain: `if (lane == i) store` compiles to a divergent jump table)
            float sv = S[0];
#pragma unroll
            for (int i = 1; i < KS; ++i) sv = (lane == i) ? S[i] : sv;
            if (lane < KS) sred[warp * KS + lane] = sv;
          }
          if (tid == 0 && n < 40) PP_TRACE(8 + n * 8 + 1);
          __syncwarp();
          if (lane == 0) mbar_arrive(u_ready_of(sx));
          return;
        }
        float ua[KS], ub[KS];
#pragma unroll
        for (int i = 0; i < KS; ++i) { ua[i] = 0.f; ub[i] = 0.f; }
        if (NP > 0) {
          mbar_wait(&u_full[n & 1], (n >> 1) & 1);
          tc::fence_after();
          const uint32_t ucol = tlane + C::COL_U + C::U_STRIDE * (n & 1);
          tc::tmem_ldn(ucol, ua);
          if (C::HAS_B) tc::tmem_ldn(ucol + C::U_B, ub);
          tc::fence_before();
          __syncwarp();
          if (lane == 0) tc::arrive(&u_accfree[n & 1]);
        }
        const int sx = n % NUS;  // the op's update stream
        float* ustage = ustage_of(sx);
        float* sred = sred_of(sx);
        if (n >= NUS) mbar_wait(u_free_of(sx), (uint32_t)((n / NUS - 1) & 1));  // the stream's previous update has read sred / the U staging
        {
          // features: M = 128 block -> lane index 32 w + lane; M = 64 blocks -> lanes 0-15 of every quarter
          const int da = (C::MA == 128) ? warp * 32 + lane : warp * 16 + r16;
          const bool oka = (C::MA == 128) || lane < 16;
#pragma unroll
          for (int i = 0; i < KS; ++i)
            if (i < KB && oka) ustage[i * UP + da] = ua[i];
          if (C::HAS_B && lane < 16) {
#pragma unroll
            for (int i = 0; i < KS; ++i)
              if (i < KB) ustage[i * UP + 128 + warp * 16 + r16] = ub[i];
          }
          {  // lane i keeps slot i's sum (a select chain: `if (lane == i) store` compiles to a divergent jump table)
            float sv = S[0];
#pragma unroll
            for (int i = 1; i < KS; ++i) sv = (lane == i) ? S[i] : sv;
            if (lane < KS) sred[warp * KS + lane] = sv;
          }
        }
        if (tid == 0 && n < 40) PP_TRACE(8 + n * 8 + 1);
        __syncwarp();
        if (lane == 0) mbar_arrive(u_ready_of(sx));
      };
      float Sprev[KS];
      int n_prev = -1;
      for (int n = 0; n < total_ops; ++n) {
        int l, c, t, m_;
        op_of4(n, l, c, t, m_);
        const int img = image_of(l, m_);
        const bool last = (t == T - 1);
        float Sl[KS];
#pragma unroll
        for (int i = 0; i < KS; ++i) Sl[i] = 0.f;
        for (int p = 0; p < NP; ++p, ++gp) {
          const uint32_t lb = gp & 1, wb = gp % NWB;
          const bool tt = tracer && tid == 0 && n == TRACE_OP && p < 4;
          if (tt) a.trace[340 + p * 12 + 0] = clock64();
          if (C::BT) {
            // ---- 256-token step: this thread's two tokens are rows 32 w + lane and 128 + 32 w + lane of the tile; their
            // logits sit side by side in the thread's tensor-memory lane (columns [0, KS) and [16, 16 + KS) of the buffer)
            mbar_wait(&lg_full[lb], (gp >> 1) & 1);
            if (tt) a.trace[340 + p * 12 + 1] = clock64();
            if (tid == 0 && p == 0 && n < 40) PP_TRACE(8 + n * 8);
            tc::fence_after();
            uint32_t r[32];
            tc::tmem_ld32_nowait(tlane + C::COL_LG + C::LG_STRIDE * lb, r);
            tc::tmem_ld_wait();
            tc::fence_before();
            __syncwarp();
            if (lane == 0) tc::arrive(&lg_empty[lb]);
            if (tt) a.trace[340 + p * 12 + 2] = clock64();
            const int tp = warp * 32 + lane;
            float xs[2][KS];
            bool ok[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int tok = (tile0 + p) * HT + 128 * h + tp;
              ok[h] = tok < N;
              float mx = -INFINITY;
#pragma unroll
              for (int i = 0; i < KS; ++i) {
                xs[h][i] = (i < K) ? __uint_as_float(r[16 * h + i]) : -INFINITY;
                mx = fmaxf(mx, xs[h][i]);
              }
              float sum = 0.f;
#pragma unroll
              for (int i = 0; i < KS; ++i) {
                xs[h][i] = ex2f(xs[h][i] - mx);
                sum += xs[h][i];
              }
              const float inv = __fdividef(1.f, sum);
#pragma unroll
              for (int i = 0; i < KS; ++i) xs[h][i] *= inv;
              if (last && a.attn_out != nullptr && ok[h]) {
                float* ao = a.attn_out + ((size_t)img * N + tok) * K;
                if ((K & 1) == 0) {
#pragma unroll
                  for (int i = 0; i < KS; i += 2)
                    if (i < K) __stcs(reinterpret_cast<float2*>(ao + i), make_float2(xs[h][i], xs[h][i + 1]));
                } else {
#pragma unroll
                  for (int i = 0; i < KS; ++i)
                    if (i < K) __stcs(ao + i, xs[h][i]);
                }
              }
            }
            if (tt) a.trace[340 + p * 12 + 3] = clock64();
            if (gp >= (uint32_t)NWB) mbar_wait(&w_empty[wb], ((gp / NWB) - 1) & 1);
            if (tt) a.trace[340 + p * 12 + 4] = clock64();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              unsigned char* wrow = wtiles + wb * C::WP_BYTES + (2 * h + (tp >> 6)) * C::WH_BYTES + ((tp & 63) >> 3) * 256 + (tp & 7) * 2;
#pragma unroll
              for (int i = 0; i < KS; ++i) {
                const float wv = (ok[h] && i < K) ? xs[h][i] + a.eps : 0.f;
                const __nv_bfloat16 wq = __float2bfloat16_rn(wv);
                *reinterpret_cast<__nv_bfloat16*>(wrow + i * 16) = wq;
                Sl[i] += __bfloat162float(wq);
              }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) tc::arrive(&w_full[wb]);
            if (tt) a.trace[340 + p * 12 + 5] = clock64();
            continue;
          }
          mbar_wait(&lg_full[lb], (gp >> 1) & 1);
          if (tt) a.trace[340 + p * 12 + 1] = clock64();
          if (tid == 0 && p == 0 && n < 40) PP_TRACE(8 + n * 8);
          tc::fence_after();
          float x[KS];
          tc::tmem_ldn(tlane + C::COL_LG + C::LG_STRIDE * lb, x);
          tc::fence_before();
          __syncwarp();
          if (lane == 0) tc::arrive(&lg_empty[lb]);
          if (tt) a.trace[340 + p * 12 + 2] = clock64();
          const int ht = 2 * p + half;
          const int tok = (tile0 + ht) * HT + tih;
          const bool tok_ok = (ht < TP) && (tok < N);
          float mx = -INFINITY;
#pragma unroll
          for (int i = 0; i < KS; ++i) {
            x[i] = (i < K) ? x[i] : -INFINITY;
            mx = fmaxf(mx, x[i]);
          }
          float sum = 0.f;
#pragma unroll
          for (int i = 0; i < KS; ++i) {
            x[i] = ex2f(x[i] - mx);  // q carries log2(e)
            sum += x[i];
          }
          const float inv = __fdividef(1.f, sum);
#pragma unroll
          for (int i = 0; i < KS; ++i) x[i] *= inv;
          if (last && a.attn_out != nullptr && tok_ok) {
            float* ao = a.attn_out + ((size_t)img * N + tok) * K;
            if ((K & 1) == 0) {
#pragma unroll
              for (int i = 0; i < KS; i += 2)
                if (i < K) __stcs(reinterpret_cast<float2*>(ao + i), make_float2(x[i], x[i + 1]));
            } else {
#pragma unroll
              for (int i = 0; i < KS; ++i)
                if (i < K) __stcs(ao + i, x[i]);
            }
          }
          // weights rounded to bf16 once; the same rounded values feed the numerator and the token sum
          if (tt) a.trace[340 + p * 12 + 3] = clock64();
          if (gp >= (uint32_t)NWB) mbar_wait(&w_empty[wb], ((gp / NWB) - 1) & 1);
          if (tt) a.trace[340 + p * 12 + 4] = clock64();
          unsigned char* wrow = wtiles + wb * C::WP_BYTES + half * C::WH_BYTES + (tih >> 3) * 256 + (tih & 7) * 2;
#pragma unroll
          for (int i = 0; i < KS; ++i) {
            const float wv = (tok_ok && i < K) ? x[i] + a.eps : 0.f;
            const __nv_bfloat16 wq = __float2bfloat16_rn(wv);
            *reinterpret_cast<__nv_bfloat16*>(wrow + i * 16) = wq;
            Sl[i] += __bfloat162float(wq);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) tc::arrive(&w_full[wb]);
          if (tt) a.trace[340 + p * 12 + 5] = clock64();
          if (n_prev >= 0 && p == 0) {  // the previous op's accumulator, one softmax later
            drain(n_prev, Sprev);
            n_prev = -1;
          }
        }
        const bool te = tracer && tid == 0 && n == TRACE_OP;
        if (te) a.trace[340 + 11] = clock64();
        // (stage by stage over all slots: KS independent shuffles in flight instead of KS serial 5-step chains)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          float o[KS];
#pragma unroll
          for (int i = 0; i < KS; ++i) o[i] = __shfl_xor_sync(FULL, Sl[i], off);
#pragma unroll
          for (int i = 0; i < KS; ++i) Sl[i] += o[i];
        }
        if (te) a.trace[352 + 11] = clock64();
        if (n_prev >= 0) {  // (only when this op had no pair)
          drain(n_prev, Sprev);
          n_prev = -1;
        }
        // defer only when the next op belongs to another lane: a lane's next pass needs the query its own update produces
        int l_next = -1, c_next = 0;
        if (n + 1 < total_ops) op_of(n + 1, l_next, c_next);
        if (C::DEFER && !C::UDRAIN && l_next >= 0 && l_next != l && NP > 0) {
#pragma unroll
          for (int i = 0; i < KS; ++i) Sprev[i] = Sl[i];
          n_prev = n;
        } else {
          drain(n, Sl);
        }
      }
    } else if (warp == 4 || warp == 6) {
      // ---- TMA producers: the rest of the stream (the first ring-full was issued during the setup)
      if (is_prod)
        while (pj < total_ht) produce();
    } else if (warp == 5) {
      // ---- MMA issuer of the logits: the whole warp runs the control flow (uniform operands), one elected lane issues.
      // (The U products have an issuer warp of their own: an mbarrier wait costs ~100 cycles even when it is already
      // complete, and a single issuer spent as long in its six waits per 128-token pair as in the MMAs.)
      const bool leader = tc::elect_one();
      constexpr uint32_t ID_LG = tc::idesc_bf16(64, KS);
      uint32_t gp = 0;  // pairs whose logit products have been issued
      int jk = 0;       // half tiles consumed from the k ring
      for (int n = 0; n < total_ops; ++n) {
        int l, c;
        op_of(n, l, c);
        if (NP == 0) continue;
        mbar_wait(&q_ready[l], (uint32_t)(c & 1));
        fence_proxy_async();
        tc::fence_after();
        const uint32_t qa = smem_u32(qop(l));
        const bool tn = tracer && lane == 0 && n == TRACE_OP;
        for (int p = 0; p < NP; ++p, ++gp) {
          const uint32_t lb = gp & 1;
          if (tn && p < 4) a.trace[340 + p * 12 + 6] = clock64();
          if (gp >= 2) {
            mbar_wait(&lg_empty[lb], ((gp >> 1) - 1) & 1);
            tc::fence_after();
          }
          if (C::BT) {  // one 256-token tile: two M = 128 products (N = 16: with 8 slot columns the second group aliases the first)
            constexpr uint32_t ID_LG2 = tc::idesc_bf16(128, 16);
            const int s = jk % NKS;
            mbar_wait(&k_full[s], (uint32_t)((jk / NKS) & 1));
            tc::fence_after();
            const uint32_t ka = smem_u32(kring + (size_t)s * C::HT_BYTES);
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                if (leader)
                  tc::mma_bf16(tmem + C::COL_LG + C::LG_STRIDE * lb + 16 * m, tc::smem_desc(ka + m * (128 * 128) + ks * 32, 16, 1024, tc::SW_128),
                               tc::smem_desc(qa + ks * (32 * KS), 16 * KS, KS == 16 ? 128 : 0, tc::SW_NONE), ID_LG2, ks != 0);
            if (leader) {
              tc::commit(&k_empty[s]);
              tc::commit(&lg_full[lb]);
            }
            __syncwarp();
            ++jk;
            if (tn && p < 4) a.trace[340 + p * 12 + 7] = clock64();
            continue;
          }
          for (int h = 0; h < 2 && 2 * p + h < TP; ++h) {
            const int s = jk % NKS;
            mbar_wait(&k_full[s], (uint32_t)((jk / NKS) & 1));
            tc::fence_after();
            const uint32_t ka = smem_u32(kring + (size_t)s * C::HT_BYTES);
            const uint32_t dcol = tmem + C::COL_LG + C::LG_STRIDE * lb + ((uint32_t)(16 * h) << 16);
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                if (leader)
                  tc::mma_bf16(dcol, tc::smem_desc(ka + ch * C::CH_BYTES + ks * 32, 16, 1024, tc::SW_128),
                               tc::smem_desc(qa + (ch * 4 + ks) * (32 * KS), 16 * KS, 128, tc::SW_NONE), ID_LG, (ch | ks) != 0);
            if (leader) tc::commit(&k_empty[s]);
            __syncwarp();
            ++jk;
          }
          if (leader) tc::commit(&lg_full[lb]);
          __syncwarp();
          if (tn && p < 4) a.trace[340 + p * 12 + 7] = clock64();
        }
      }
    } else if (warp == 7) {
      // ---- MMA issuer of U^T += v^T w
      const bool leader = tc::elect_one();
      constexpr uint32_t ID_UA = tc::idesc_bf16(C::MA, 16, 1, 0);
      constexpr uint32_t ID_UB = tc::idesc_bf16(64, 16, 1, 0);
      uint32_t gu = 0;  // pairs whose U products have been issued
      int jv = 0;       // half tiles consumed from the v ring
      for (int n = 0; n < total_ops; ++n) {
        if (NP == 0) continue;
        const uint32_t ucol = tmem + C::COL_U + C::U_STRIDE * (n & 1);
        const bool tn = tracer && lane == 0 && n == TRACE_OP;
        for (int p = 0; p < NP; ++p, ++gu) {
          const uint32_t wb = gu % NWB;
          if (tn && p < 4) a.trace[340 + p * 12 + 8] = clock64();
          mbar_wait(&w_full[wb], (gu / NWB) & 1);
          if (tn && p < 4) a.trace[340 + p * 12 + 9] = clock64();
          tc::fence_after();
          if (p == 0 && n >= 2) {  // the accumulator was last used by op n - 2
            mbar_wait(&u_accfree[n & 1], (uint32_t)(((n >> 1) - 1) & 1));
            tc::fence_after();
          }
          if (C::BT) {  // 256 tokens = 16 K steps on one ring tile and one w tile
            const int s = jv % NVS;
            mbar_wait(&v_full[s], (uint32_t)((jv / NVS) & 1));
            tc::fence_after();
            const uint32_t va = smem_u32(vring + (size_t)s * C::HT_BYTES);
            const uint32_t wa = smem_u32(wtiles + wb * C::WP_BYTES);
#pragma unroll
            for (int ks = 0; ks < 16; ++ks)
              if (leader)
                tc::mma_bf16(ucol, tc::smem_desc(va + ks * 2048, 1024, 1024, tc::SW_128), tc::smem_desc(wa + ks * 512, 256, 128, tc::SW_NONE), ID_UA,
                             (uint32_t)((p | ks) != 0));
            if (leader) {
              tc::commit(&v_empty[s]);
              tc::commit(&w_empty[wb]);
            }
            __syncwarp();
            ++jv;
            if (tn && p < 4) a.trace[340 + p * 12 + 10] = clock64();
            continue;
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
              if (leader) {
                tc::mma_bf16(ucol, tc::smem_desc(va + ks * 2048, C::CH_BYTES, 1024, tc::SW_128), db, ID_UA, acc);
                if (C::HAS_B)
                  tc::mma_bf16(ucol + C::U_B, tc::smem_desc(va + 2 * C::CH_BYTES + ks * 2048, C::CH_BYTES, 1024, tc::SW_128), db, ID_UB, acc);
              }
            }
            if (leader) tc::commit(&v_empty[s]);
            __syncwarp();
            ++jv;
          }
          if (leader) tc::commit(&w_empty[wb]);
          __syncwarp();
          if (tn && p < 4) a.trace[340 + p * 12 + 10] = clock64();
        }
        if (leader) tc::commit(&u_full[n & 1]);
        __syncwarp();
      }
    }
    __syncwarp();
  } else {
    // ==================================================================================== UPDATE ENGINE
    // Every exchange round is  push -> wait -> tensor-core product -> accumulator rows to shared memory -> one barrier
    // -> epilogue (= next push): all-gather payloads travel as bf16 straight into the receivers' MMA operand buffers
    // (act[parity], the lane's slots and q operands); the LayerNorms are folded into the products that follow them
    // (statistics from the received rows).
    const int utid = tid - 256 - us * UT, uwarp = warp - 8 - us * UW;  // thread / warp inside the stream
    auto upd_sync = [&]() { ocrl::umma::upd_sync(1 + us, UT); };
    const uint32_t col_gi = us ? C::COL_GI2 + 32u * (uint32_t)(us - 1) : C::COL_GI, col_gh = us ? C::COL_GH2 + 32u * (uint32_t)(us - 1) : C::COL_GH;
    float* ustage = ustage_of(us);
    float* sred = sred_of(us);
    uint64_t* u_ready = u_ready_of(us);
    uint64_t* u_free = u_free_of(us);
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
    // of a quad assemble 4 features (8 bytes of bf16) and each sends them to CL/4 CTAs, into the operand buffer `dst`.
    auto quad_push = [&](uint32_t r, float val, int i, int SL, unsigned char* dst) {
      const int qb = lane & ~3;
      const float v0 = __shfl_sync(FULL, val, qb), v1 = __shfl_sync(FULL, val, qb + 1);
      const float v2 = __shfl_sync(FULL, val, qb + 2), v3 = __shfl_sync(FULL, val, qb + 3);
      if (i < K * SL) {
        const int slot = i / SL, f = rank * SL + ((i % SL) & ~3);
        const uint32_t lbuf = smem_u32(dst) + (uint32_t)opnd_off(slot, f);
        const uint32_t lbar = smem_u32(&xbar[r & 1]);
        const uint32_t lo = pack_bf16x2(v0, v1), hi = pack_bf16x2(v2, v3);
#pragma unroll
        for (int q = 0; q < CL / 4; ++q) {
          const int dest = (lane & 3) + 4 * q;
          st_async_v2(mapa_u32(lbuf, dest), lo, hi, mapa_u32(lbar, dest));
        }
      }
    };
    // mean / rstd of the K rows (bf16, length D, operand layout) that just arrived: one warp per row
    auto row_stats = [&](const unsigned char* opnd) {
      // The round sits on the critical path of every update twice (traced: 1.7-2.3 k cycles as two rounds of two serial
      // butterfly reductions each).  One pass (sum and sum of squares), two rows per warp and round: four independent
      // shuffle chains in flight, K <= 8 rows in a single round.
      constexpr int NCD = D / 64;  // (rows of slot features; NCH counts the chunks of the streamed tiles)
      for (int row0 = uwarp; row0 < K; row0 += 2 * UW) {
        const int row1 = row0 + UW;
        const bool two = row1 < K;
        float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
#pragma unroll
        for (int c = 0; c < NCD; ++c) {
          const __nv_bfloat162 v0 = *reinterpret_cast<const __nv_bfloat162*>(opnd + opnd_off(row0, 64 * c + 2 * lane));
          const __nv_bfloat162 v1 = *reinterpret_cast<const __nv_bfloat162*>(opnd + opnd_off(two ? row1 : row0, 64 * c + 2 * lane));
          const float x0 = __low2float(v0), y0 = __high2float(v0), x1 = __low2float(v1), y1 = __high2float(v1);
          s0 += x0 + y0;
          q0 = fmaf(x0, x0, fmaf(y0, y0, q0));
          s1 += x1 + y1;
          q1 = fmaf(x1, x1, fmaf(y1, y1, q1));
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          const float a0 = __shfl_xor_sync(FULL, s0, off), b0 = __shfl_xor_sync(FULL, q0, off);
          const float a1 = __shfl_xor_sync(FULL, s1, off), b1 = __shfl_xor_sync(FULL, q1, off);
          s0 += a0; q0 += b0; s1 += a1; q1 += b1;
        }
        const float m0 = s0 * (1.f / D), m1 = s1 * (1.f / D);
        const float r0 = rsqrtf(fmaxf(q0 * (1.f / D) - m0 * m0, 0.f) + a.ln_eps);
        const float r1 = rsqrtf(fmaxf(q1 * (1.f / D) - m1 * m1, 0.f) + a.ln_eps);
        if (lane == 0) {
          s_mean[row0] = m0; s_rstd[row0] = r0;
          if (two) { s_mean[row1] = m1; s_rstd[row1] = r1; }
        }
      }
    };
    // Products: D[128 weight rows x 16] = W block (tensor memory, `wcol`) x operand (K-major [k/8][8 slots][8] bf16,
    // with 8 slot columns the second 8-row group is aliased onto the first: SBO = 0), `ksteps` of 16 features.  Warp 0 of the engine issues
    // (uniform control flow, one elected lane) and commits to ubar[which]; every thread of the engine waits for the
    // commits of a barrier in order (`uph` counts them; a barrier is never more than one phase ahead of its waiters:
    // W_hh h, which stays in flight across other products, has a barrier of its own); warps 0-3 / 4-7 then move
    // accumulator rows to P_GI / P_GH.
    constexpr uint32_t ID_UPD = tc::idesc_bf16(128, 16);
    uint32_t uph[2] = {0u, 0u};
    auto product = [&](uint32_t wcol, const unsigned char* opnd, int ksteps, uint32_t dcol, int which) {
      if (uwarp == 0) {
        fence_proxy_async();  // the operand was written through the generic proxy (st.async / st.shared)
        tc::fence_after();
        const uint32_t oa = smem_u32(opnd);
        if (tc::elect_one()) {
          for (int ks = 0; ks < ksteps; ++ks)
            tc::mma_bf16_ts(tmem + dcol, tmem + wcol + 8 * ks, tc::smem_desc(oa + ks * (32 * KS), 16 * KS, KS == 16 ? 128 : 0, tc::SW_NONE), ID_UPD, ks != 0);
          tc::commit(&ubar[which]);
        }
        __syncwarp();
      }
    };
    auto product_wait = [&](int which) {
      mbar_wait(&ubar[which], uph[which] & 1);
      ++uph[which];
      tc::fence_after();
    };
    const uint32_t tq = tmem + ((uint32_t)((uwarp & 3) * 32) << 16);  // this warp's quarter of the lanes
    auto unload = [&](uint32_t dcol, float* Pr, int group) {
      if (UW == 4 || (uwarp >> 2) == group) {  // four warps cover the 128 lanes; eight split the two accumulators
        float vs[KS];
        tc::tmem_ldn(tq + dcol, vs);
        float4* o = reinterpret_cast<float4*>(Pr + ((uwarp & 3) * 32 + lane) * KS);
#pragma unroll
        for (int i = 0; i < KS / 4; ++i) o[i] = make_float4(vs[4 * i], vs[4 * i + 1], vs[4 * i + 2], vs[4 * i + 3]);
        tc::fence_before();
      }
    };
    // slots of image `img` (fp32, global) -> this CTA's bf16 operand slh[l] and the fp32 own slice
    auto load_slots0 = [&](int l, int img) {
      const float* src = a.slots0 + (size_t)img * K * D;
      unsigned char* dst = slh_hi(l);
      for (int i = utid; i < K * (D / 2); i += UT) {
        const int slot = i / (D / 2), c2 = i % (D / 2);
        const float2 x = __ldg(reinterpret_cast<const float2*>(src + slot * D + 2 * c2));
        *reinterpret_cast<uint32_t*>(dst + opnd_off(slot, 2 * c2)) = pack_bf16x2(x.x, x.y);
      }
      float* own = own_of(l);
      for (int i = utid; i < K * DS; i += UT) own[i] = __ldg(src + (i / DS) * D + rank * DS + i % DS);
      fence_proxy_async();
      upd_sync();
    };
    // q = W_q LN(slots) of lane l for the CTA's slice (slots = slh[l], bf16), all-gathered (times log2 e) into the lane's q operand
    auto q_phase = [&](int l, int q_img, int q_t) {
      product(C::COL_WY, slh_hi(l), D / 16, col_gh, 0);
      row_stats(slh_hi(l));
      arm(round, (uint32_t)(K * FW * 2));
      product_wait(0);
      unload(col_gh, P_GH, 1);
      upd_sync();
      for (int i0 = uwarp * 32; i0 < K * FS; i0 += UT) {  // (factored: q'' = s W_k^T q, FS features per CTA)
        const int i = i0 + lane;
        float val = 0.f;
        if (i < K * FS) {
          const int slot = i / FS, dl = i % FS;
          const float acc = P_GH[(3 * DS + dl) * KS + slot];
          const float qv = s_rstd[slot] * (acc - s_mean[slot] * s_cq[dl]) + s_bqf[dl];
          if (!XH && a.saved != nullptr) saved_at(q_img, q_t)[SL.off_q() + slot * D + rank * DS + dl] = qv;
          val = qv * LOG2E;
        }
        quad_push(round, val, i, FS, qop(l));
      }
      xwait(round);
      ++round;
      if (utid == 0) {  // the MMA issuer may read lane l's q operand (tcgen05.mma reads it through the async proxy)
        fence_proxy_async();
        if (NUS > 1) {  // one writer at a time per lane (a lane's updates are ordered), so a plain increment is enough
          const uint32_t cnt = q_count[l] + 1u;
          asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(&q_count[l])), "r"(cnt) : "memory");
        }
        mbar_arrive(&q_ready[l]);
      }
    };

    // initial queries of the lanes
    for (int l = us; l < NL; l += NUS)
      if (nops[l] > 0) {
        if (l > 0) load_slots0(l, image_of(l, 0));  // lane 0's were staged during the setup
        q_phase(l, image_of(l, 0), 0);
      }

    for (int n = us; n < total_ops; n += NUS) {  // the stream's ops
      const int ns = n / NUS;  // ... counted per stream (barrier phases)
      int l, c, t, m;
      op_of4(n, l, c, t, m);
      const int img = image_of(l, m);
      const bool last = (t == T - 1);
      float* own = own_of(l);
      const bool tr_on = tracer && utid == 0 && n < 40;
#define PP_T(i) do { if (tr_on) a.trace[8 + n * 8 + (i)] = clock64(); } while (0)
      // ============================================================ R1: reduce-scatter of sum w v, all-reduce of sum w
      arm(round, (uint32_t)(K * FW * 4 + 4 * KS * CL));
      // gh = W_hh h only needs the slots that entered the iteration: it runs under the pass and the R1 round trip.
      // The lane's previous update may have run on the other stream: its last act was the lane's query for this op
      // (phase c of q_ready[l]).  A parity wait cannot tell how far the barrier is -- this stream may come here one phase
      // early or one phase late -- so the streams follow a per-lane count of the queries published so far.
      if (NUS > 1) {
        uint32_t seen;
        do {
          asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(seen) : "r"(smem_u32(&q_count[l])) : "memory");
        } while (seen < (uint32_t)(c + 1));
      }
      product(C::COL_WY, slh_hi(l), D / 16, col_gh, 1);
      mbar_wait(u_ready, (uint32_t)(ns & 1));
      if (C::UDRAIN) {  // the op's accumulator: tensor memory -> the stream's staging buffer
        if (uwarp < 4) {
          float ua[KS], ub[KS];
#pragma unroll
          for (int i = 0; i < KS; ++i) { ua[i] = 0.f; ub[i] = 0.f; }
          if (NP > 0) {
            mbar_wait(&u_full[n & 1], (uint32_t)((n >> 1) & 1));
            tc::fence_after();
            const uint32_t ucol = tq + C::COL_U + C::U_STRIDE * (n & 1);
            tc::tmem_ldn(ucol, ua);
            if (C::HAS_B) tc::tmem_ldn(ucol + C::U_B, ub);
            tc::fence_before();
            __syncwarp();
            if (lane == 0) tc::arrive(&u_accfree[n & 1]);
          }
          const int r16 = lane & 15;
          const int da = (C::MA == 128) ? uwarp * 32 + lane : uwarp * 16 + r16;
          const bool oka = (C::MA == 128) || lane < 16;
#pragma unroll
          for (int i = 0; i < KS; ++i)
            if (i < KB && oka) ustage[i * UP + da] = ua[i];
          if (C::HAS_B && lane < 16) {
#pragma unroll
            for (int i = 0; i < KS; ++i)
              if (i < KB) ustage[i * UP + 128 + uwarp * 16 + r16] = ub[i];
          }
        }
        upd_sync();
      }
      PP_T(2);
      {
        const uint32_t lbuf = smem_u32(xb), lbar = smem_u32(&xbar[round & 1]);
        constexpr int QPR = FW / 4;
        for (int i = utid; i < K * QPR; i += UT) {
          const int slot = i / QPR, d = 4 * (i % QPR);
          const int dest = d / FS, dl = d % FS;
          const float4 val = *reinterpret_cast<const float4*>(ustage + slot * UP + d);
          const uint32_t off = (uint32_t)(((rank * KB + slot) * FS + dl) * 4);
          st_async_v4(mapa_u32(lbuf + off, dest), val, mapa_u32(lbar, dest));
        }
        static_assert(KS * CL <= UT, "one thread per (slot, destination) of the token sums");
        if (utid < KS * CL) {
          const int slot = utid % KS, dest = utid / KS;
          const float s = sred[slot] + sred[KS + slot] + sred[2 * KS + slot] + sred[3 * KS + slot];
          const uint32_t off = (uint32_t)(KB * D * 4 + (rank * KS + slot) * 4);
          st_async_b32(mapa_u32(lbuf + off, dest), __float_as_uint(s), mapa_u32(lbar, dest));
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(u_free);  // the pass engine may overwrite the staging buffers
      xwait(round);
      ++round;
      PP_T(3);
      // ============================================================ R2: all-gather updates = sum over CTAs / token sum
      arm(round, (uint32_t)(K * FW * 2));
      {
        const float* rs = reinterpret_cast<const float*>(xb);
        const float* ss = rs + KB * D;
        for (int i0 = uwarp * 32; i0 < K * FS; i0 += UT) {
          const int i = i0 + lane;
          float val = 0.f;
          if (i < K * FS) {
            const int slot = i / FS, dl = i % FS;
            float u = 0.f, sw = 0.f;
#pragma unroll
            for (int src = 0; src < CL; ++src) {
              u += rs[(src * KB + slot) * FS + dl];
              sw += ss[src * KS + slot];
            }
            val = u / sw;  // factored: the weighted mean of x^ (W_v is folded into W_ih'')
            if (!XH && a.saved != nullptr) {
              float* sv = saved_at(img, t);
              sv[SL.off_h() + slot * D + rank * DS + dl] = own[i];  // slots entering the iteration
              sv[SL.off_u() + slot * D + rank * DS + dl] = val;
              if (rank == 0 && dl == 0) sv[SL.off_s() + slot] = sw;
            }
          }
          quad_push(round, val, i, FS, act(round));
        }
      }
      xwait(round);
      // ---- GRU: gi = W_ih u
      product(C::COL_WX, act(round), FW / 16, col_gi, 0);
      ++round;
      arm(round, (uint32_t)(K * D * 2));
      product_wait(1);  // gh
      product_wait(0);  // gi
      unload(col_gi, P_GI, 0);
      unload(col_gh, P_GH, 1);
      upd_sync();
      PP_T(4);
      // ============================================================ R3: all-gather h'
      // Factored form: the K * DS elements of the slice are one and a bit rounds of the stream's threads (144 on 128);
      // two elements per thread as ONE straight-line block (clamped indices, results masked afterwards) let the second
      // round overlap the first instead of following it (traced: 2.2 k cycles for the two serial rounds), and the gates
      // take the single-instruction tanh (tanh.approx.f32: 2^-11 relative, below the bf16 rounding of their operands).
      if (XH) {
        const int NE = K * DS;
        auto gate = [&](int i) {
          const int slot = i / DS, dl = i - slot * DS;
          const float gir = P_GI[(dl) * KS + slot] + s_bih[dl], ghr = P_GH[(dl) * KS + slot] + s_bhh[dl];
          const float giz = P_GI[(DS + dl) * KS + slot] + s_bih[DS + dl];
          const float ghz = P_GH[(DS + dl) * KS + slot] + s_bhh[DS + dl];
          const float gin = P_GI[(2 * DS + dl) * KS + slot] + s_bih[2 * DS + dl];
          const float ghn = P_GH[(2 * DS + dl) * KS + slot] + s_bhh[2 * DS + dl];
          const float r = fmaf(0.5f, tanh_approx(0.5f * (gir + ghr)), 0.5f), z = fmaf(0.5f, tanh_approx(0.5f * (giz + ghz)), 0.5f);
          const float nn = tanh_approx(fmaf(r, ghn, gin));
          return fmaf(z, own[i] - nn, nn);  // (1 - z) n + z h
        };
        for (int i0 = uwarp * 32; i0 < NE; i0 += 2 * UT) {
          const int iA = i0 + lane, iB = iA + UT;
          const bool second = (i0 + UT) < NE;  // warp-uniform
          const float hA = gate(min(iA, NE - 1));
          const float hB = gate(min(iB, NE - 1));
          if (iA < NE) own[iA] = hA;
          if (iB < NE) own[iB] = hB;
          quad_push(round, iA < NE ? hA : 0.f, iA, DS, act(round));
          if (second) quad_push(round, iB < NE ? hB : 0.f, iB, DS, act(round));
        }
      } else
      for (int i0 = uwarp * 32; i0 < K * DS; i0 += UT) {
        const int i = i0 + lane;
        float hp = 0.f;
        if (i < K * DS) {
          const int slot = i / DS, dl = i % DS;
          const float gir = P_GI[(dl) * KS + slot] + s_bih[dl], ghr = P_GH[(dl) * KS + slot] + s_bhh[dl];
          const float giz = P_GI[(DS + dl) * KS + slot] + s_bih[DS + dl];
          const float ghz = P_GH[(DS + dl) * KS + slot] + s_bhh[DS + dl];
          const float gin = P_GI[(2 * DS + dl) * KS + slot] + s_bih[2 * DS + dl];
          const float ghn = P_GH[(2 * DS + dl) * KS + slot] + s_bhh[2 * DS + dl];
          const float r = sigmoidf_(gir + ghr), z = sigmoidf_(giz + ghz);
          const float nn = tanhf(gin + r * ghn);
          hp = (1.f - z) * nn + z * own[i];
          own[i] = hp;
          if (!XH && a.saved != nullptr) {
            float* sv = saved_at(img, t);
            const int f = slot * D + rank * DS + dl;
            sv[SL.off_r() + f] = r;
            sv[SL.off_z() + f] = z;
            sv[SL.off_n() + f] = nn;
            sv[SL.off_ghn() + f] = ghn;
            sv[SL.off_hp() + f] = hp;
          }
        }
        quad_push(round, hp, i, DS, act(round));
      }
      xwait(round);
      // ---- MLP layer 1 on the raw h' (LayerNorm folded), statistics alongside
      product(C::COL_WX, act(round), D / 16, col_gi, 0);
      row_stats(act(round));
      ++round;
      arm(round, (uint32_t)(K * H * 2));
      product_wait(0);
      unload(col_gi, P_GI, 0);
      upd_sync();
      PP_T(5);
      // ============================================================ R4: all-gather the MLP hidden layer
      for (int i0 = uwarp * 32; i0 < K * HS; i0 += UT) {
        const int i = i0 + lane;
        float hid = 0.f;
        if (i < K * HS) {
          const int slot = i / HS, hl = i % HS;
          const float acc = P_GI[(3 * DS + hl) * KS + slot];
          const float pre = s_rstd[slot] * (acc - s_mean[slot] * s_c1[hl]) + s_b1f[hl];
          if (!XH && a.saved != nullptr) saved_at(img, t)[SL.off_pre() + slot * H + rank * HS + hl] = pre;
          hid = fmaxf(pre, 0.f);
        }
        quad_push(round, hid, i, HS, act(round));
      }
      xwait(round);
      product(C::COL_WX, act(round), H / 16, col_gi, 0);
      ++round;
      if (!last) arm(round, (uint32_t)(K * D * 2));
      product_wait(0);
      unload(col_gi, P_GI, 0);
      upd_sync();
      PP_T(6);
      // ============================================================ R5: all-gather the new slots (or write them out)
      const bool more = last && (m + 1 < nimg_of(l));
      for (int i0 = uwarp * 32; i0 < K * DS; i0 += UT) {
        const int i = i0 + lane;
        float sn = 0.f;
        if (i < K * DS) {
          const int slot = i / DS, dl = i % DS;
          sn = own[i] + (s_b2[dl] + P_GI[(3 * DS + HS + dl) * KS + slot]);
          own[i] = sn;
          if (last) a.slots_out[((size_t)img * K + slot) * D + rank * DS + dl] = sn;
        }
        if (!last) quad_push(round, sn, i, DS, slh_hi(l));
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
  if (tracer && tid == 0) a.trace[3] = clock64();
  if (tracer2) a.trace[6] = clock64();
  if (warp == 7) tc::tmem_dealloc<C::TMEM_COLS>(tmem);
  cluster.sync();  // no CTA leaves while a peer may still address its shared memory
}


// Weight preparation, once per call: for every CTA rank the two tensor-memory blocks as bf16 pairs, column-major
// ([rank][block][word][128 rows]) so that the main kernel's row threads read them coalesced, plus the constants of the
// folded LayerNorms (W LN(x) = rstd (W' x - mean c) + W beta, W' = W diag(gamma), c = row sums of the rounded W').
// One warp per weight row.
// Factored pass (F > 0): the W_ih rows become W_ih'' = W_ih W_v (length F) and the W_q rows become the CTA's F / CL
// rows of W_q'' = s W_k^T W_q (length D, LayerNorm folded as before): a folded row is sum_d coef[d] M[d][:].
template <int D, int H, int CL, int F>
__global__ void __launch_bounds__(256) umma_prep_kernel(const ocrl_sa_weights w, const float* __restrict__ wk, const float* __restrict__ wv,
                                                        uint32_t* __restrict__ words, float* __restrict__ consts) {
  constexpr int DS = D / CL, HS = H / CL, LX = D > H ? D : H, WPB = LX / 2;
  constexpr int FS = F / CL;
  constexpr int RX = 3 * DS + HS + DS, RY = 3 * DS + DS;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= CL * 2 * 128) return;
  const int r = gw & 127, blk = (gw >> 7) & 1, rank = gw >> 8;
  const float* src = nullptr;
  const float* gam = nullptr;
  const float* bet = nullptr;
  int len = 0, fold = 0;
  // folded rows: row = cscale * sum_d coef[d * cstride] * mat[d * mstride + :]
  const float* coef = nullptr;
  const float* mat = nullptr;
  int cstride = 0, mstride = 0;
  float cscale = 1.f;
  if (blk == 0) {
    if (F > 0 && r < 3 * DS) { coef = w.w_ih + ((size_t)(r / DS) * D + rank * DS + r % DS) * D; cstride = 1; mat = wv; mstride = F; len = F; }
    else if (r < 3 * DS) { src = w.w_ih + ((size_t)(r / DS) * D + rank * DS + r % DS) * D; len = D; }
    else if (r < 3 * DS + HS) { src = w.w1 + ((size_t)rank * HS + (r - 3 * DS)) * D; len = D; fold = 1; gam = w.ln_mlp_w; bet = w.ln_mlp_b; }
    else if (r < RX) { src = w.w2 + ((size_t)rank * DS + (r - 3 * DS - HS)) * H; len = H; }
  } else {
    if (r < 3 * DS) { src = w.w_hh + ((size_t)(r / DS) * D + rank * DS + r % DS) * D; len = D; }
    else if (F > 0) {
      if (r < 3 * DS + FS) {
        coef = wk + rank * FS + (r - 3 * DS); cstride = F; mat = w.wq; mstride = D; len = D; cscale = rsqrtf((float)D);
        fold = 2; gam = w.ln_slots_w; bet = w.ln_slots_b;
      }
    }
    else if (r < RY) { src = w.wq + ((size_t)rank * DS + (r - 3 * DS)) * D; len = D; fold = 2; gam = w.ln_slots_w; bet = w.ln_slots_b; }
  }
  uint32_t* dst = words + ((size_t)(rank * 2 + blk) * WPB) * 128 + r;
  float csum = 0.f, bsum = 0.f;
  for (int w0 = 0; w0 < WPB; w0 += 32) {
    const int f = 2 * (w0 + lane);
    float2 x = make_float2(0.f, 0.f);
    if (f < len) {
      if (coef != nullptr) {
        float ax = 0.f, ay = 0.f;
#pragma unroll 8
        for (int d = 0; d < D; ++d) {
          const float c = __ldg(coef + (size_t)d * cstride);
          const float2 m = __ldg(reinterpret_cast<const float2*>(mat + (size_t)d * mstride + f));
          ax = fmaf(c, m.x, ax);
          ay = fmaf(c, m.y, ay);
        }
        x = make_float2(ax * cscale, ay * cscale);
      } else {
        x = __ldg(reinterpret_cast<const float2*>(src + f));
      }
      if (fold) {
        const float2 g = __ldg(reinterpret_cast<const float2*>(gam + f)), b = __ldg(reinterpret_cast<const float2*>(bet + f));
        bsum = fmaf(x.x, b.x, fmaf(x.y, b.y, bsum));
        x = make_float2(x.x * g.x, x.y * g.y);
      }
    }
    const __nv_bfloat162 pr = __floats2bfloat162_rn(x.x, x.y);
    csum += __low2float(pr) + __high2float(pr);
    dst[(size_t)(w0 + lane) * 128] = *reinterpret_cast<const uint32_t*>(&pr);
  }
  if (fold) {
    csum = warp_sum(csum);
    bsum = warp_sum(bsum);
    if (lane == 0) {
      float* cs = consts + (size_t)rank * (2 * HS + 2 * DS);
      const int i = r - 3 * DS;
      if (fold == 1) { cs[i] = csum; cs[HS + i] = bsum + w.b1[rank * HS + i]; }
      else { cs[2 * HS + i] = csum; cs[2 * HS + DS + i] = bsum; }
    }
  }
}

template <int D, int H, int CL>
static size_t umma_prep_bytes() {
  constexpr int LX = D > H ? D : H;
  return (size_t)CL * 2 * (LX / 2) * 128 * 4 + (size_t)CL * (2 * (H / CL) + 2 * (D / CL)) * 4;
}

template <int D, int H, int CL, int NL, int KB, int NKS, int NVS, int NWB, int NUS, int DEFER = 0, int F = 0, int BT = 0>
static int launch_umma(const IterFwdArgs& a_in, cudaStream_t stream) {
  using C = Cfg<D, H, CL, NL, KB, NKS, NVS, NWB, NUS, DEFER, F, BT>;
  IterFwdArgs a = a_in;
  if (C::XH) {
    if (a.saved != nullptr || a.xhat == nullptr || a.wk == nullptr || a.wv == nullptr) {
      set_error("sa_iter_fwd(tcgen05, factored): inference only (saved == NULL), needs x^, W_k, W_v");
      return OCRL_E_SHAPE;
    }
    a.k = a.xhat;  // one stream of tiles: the kernel reads `k` only
    a.v = a.xhat;
  }
  if (a.workspace == nullptr || a.workspace_bytes < umma_prep_bytes<D, H, CL>() + 4096 + 256) {
    set_error("sa_iter_fwd(tcgen05): needs the workspace of ocrl_sa_query_workspace (bf16 weight copies)");
    return OCRL_E_SHAPE;
  }
  {
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(a.workspace) + 255) & ~uintptr_t(255));
    a.wprep = reinterpret_cast<uint32_t*>(base);
    a.wprep_consts = reinterpret_cast<float*>(base + (size_t)CL * 2 * C::WPB * 128 * 4);
    if (!a.prepared) {
      umma_prep_kernel<D, H, CL, F><<<CL * 2 * 128 / 8, 256, 0, stream>>>(a.w, a.wk, a.wv, const_cast<uint32_t*>(a.wprep), const_cast<float*>(a.wprep_consts));
      ocrl::count_launch();
      OCRL_CHECK_CUDA(cudaGetLastError());
    }
  }
  auto kern = sa_iter_fwd_umma_kernel<D, H, CL, NL, KB, NKS, NVS, NWB, NUS, DEFER, F, BT>;
  static_assert(C::SMEM_BYTES <= 227 * 1024, "shared memory budget");
  CUtensorMap tm_k, tm_v;
  const uint64_t rows = (uint64_t)a.B * a.N;
  if (!tc::make_map_bf16_sw128(&tm_k, a.k, C::FW, rows, (uint64_t)C::FW * 2, C::HT) ||
      !tc::make_map_bf16_sw128(&tm_v, a.v, C::FW, rows, (uint64_t)C::FW * 2, C::HT)) {
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
  int want = (a.B + NL - 1) / NL;  // NL lanes per cluster
  if (a.B <= ncl) want = a.B;      // clusters to spare: one image each (a lane's iterations are serial: sharing a cluster only adds latency)
  // factored pass: a cluster's passes are serial and no longer hidden behind memory, so its time grows with its images --
  // spread the batch over every cluster (B = 16 images of 16 384 tokens: 137 us on four clusters, 5 lanes each)
  if (C::XH) want = a.B;
  if (ncl > want) ncl = want;
  {  // fewest clusters that keep the same number of image rounds (frees SMs for concurrent work)
    const int per = (a.B + ncl - 1) / ncl;
    ncl = (a.B + per - 1) / per;
  }
  if (C::OPT_MAX > 0) {  // the kernel's op table holds OPT_MAX ops per cluster: larger batches go in several launches
    if (a.T >= 256) {
      set_error("sa_iter_fwd(tcgen05, factored): T = %d (< 256)", a.T);
      return OCRL_E_SHAPE;
    }
    const int per = (a.B + ncl - 1) / ncl;
    if (per * a.T > C::OPT_MAX) {
      const int chunk = ncl * (C::OPT_MAX / a.T);
      for (int b0 = 0; b0 < a_in.B; b0 += chunk) {
        IterFwdArgs c = a_in;
        c.B = min(chunk, a_in.B - b0);
        c.xhat = reinterpret_cast<const unsigned char*>(a_in.xhat) + (size_t)b0 * a_in.N * C::FW * 2;
        c.slots0 = a_in.slots0 + (size_t)b0 * a_in.K * D;
        c.slots_out = a_in.slots_out + (size_t)b0 * a_in.K * D;
        if (a_in.attn_out != nullptr) c.attn_out = a_in.attn_out + (size_t)b0 * a_in.N * a_in.K;
        c.prepared = 1;  // the weights were prepared above (or by the caller)
        const int rc = launch_umma<D, H, CL, NL, KB, NKS, NVS, NWB, NUS, DEFER, F, BT>(c, stream);
        if (rc != OCRL_OK) return rc;
      }
      return OCRL_OK;
    }
  }
  cfg.gridDim = dim3((unsigned)(ncl * CL));
  OCRL_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, a, tm_k, tm_v));
  ocrl::count_launch();
  return OCRL_OK;
}

}  // namespace umma

static int g_dev_variant = 0;
extern "C" void ocrl_dev_iter_variant(int v) { g_dev_variant = v; }  // development knob (scripts/quick_iter.py), not in the header

// Returns OCRL_E_SHAPE (without launching) for shapes this kernel does not cover.
int sa_iter_fwd_umma_dispatch(const IterFwdArgs& a, cudaStream_t s) {
  if (a.K > 16) {
    set_error("sa_iter_fwd(tcgen05): K <= 16");
    return OCRL_E_SHAPE;
  }
  if (a.xhat != nullptr) {  // factored pass (inference)
    if (a.D == 192 && a.H == 192 && a.F == 64) {
      // five images in flight per cluster, three update streams (640 threads); from 2048 tokens up (every CTA of the
      // cluster owns at least one full tile) 256-token steps.  Measured at B = 64, N = 4096, K = 6, T = 3 (graph replays,
      // us; lanes / streams): 3 / 2: 77.0, 5 / 3: 69.5, 5 / 4: 73.4; 5 / 3 with 256-token steps: 63.1
      const bool big = a.N >= 2048 && g_dev_variant != 5;
      if (a.K <= 6) {
        if (g_dev_variant == 2) return umma::launch_umma<192, 192, 8, 5, 6, 3, 3, 2, 2, 0, 64, 1>(a, s);
        if (g_dev_variant == 4) return umma::launch_umma<192, 192, 8, 5, 6, 2, 2, 2, 4, 0, 64, 1>(a, s);
        if (big) return umma::launch_umma<192, 192, 8, 5, 6, 3, 3, 2, 3, 0, 64, 1>(a, s);
        return umma::launch_umma<192, 192, 8, 5, 6, 8, 8, 2, 3, 0, 64>(a, s);
      }
      if (a.K <= 8) {
        if (big) return umma::launch_umma<192, 192, 8, 5, 8, 3, 3, 2, 3, 0, 64, 1>(a, s);
        return umma::launch_umma<192, 192, 8, 5, 8, 8, 8, 2, 3, 0, 64>(a, s);
      }
      if (g_dev_variant == 6 || !big) {
        if (a.K <= 12) return umma::launch_umma<192, 192, 8, 3, 12, 6, 6, 2, 2, 0, 64>(a, s);
        return umma::launch_umma<192, 192, 8, 3, 16, 6, 6, 2, 2, 0, 64>(a, s);
      }
      if (a.K <= 12) return umma::launch_umma<192, 192, 8, 3, 12, 2, 2, 2, 2, 0, 64, 1>(a, s);
      return umma::launch_umma<192, 192, 8, 3, 16, 2, 2, 2, 2, 0, 64, 1>(a, s);
    }
    set_error("sa_iter_fwd(tcgen05, factored): D=%d H=%d C_in=%d not instantiated", a.D, a.H, a.F);
    return OCRL_E_SHAPE;
  }
  if (a.D == 192 && a.H == 192) {
    // Lanes per cluster when the caller does not say: the count that needs fewer rounds of images per lane on the ~15
    // clusters of eight CTAs a B200 holds; on a tie two lanes (more clusters busy: B = 16 images of 16 384 tokens run
    // 119 us on eight clusters against 163 us on six)
    int lanes = a.lanes;
    if (lanes != 2 && lanes != 3) {
      auto rounds = [&](int nl) {
        const int ncl = min(15, (a.B + nl - 1) / nl);
        return (a.B + nl * ncl - 1) / (nl * ncl);
      };
      lanes = rounds(3) < rounds(2) ? 3 : 2;
    }
    // three lanes hide the slot update completely; two keep the k/v of the images in flight inside the L2 (compulsory
    // DRAM traffic only).  Measured at B = 64 (us, three lanes / two lanes): one update stream, 4 v slots 124 / 133; two
    // streams, 3 v slots 118 / 140; two streams + deferred drain 125 / 162
    if (a.K <= 6) {
      if (lanes == 2) return umma::launch_umma<192, 192, 8, 2, 6, 3, 4, 2, 1, 0>(a, s);
      if (g_dev_variant == 1) return umma::launch_umma<192, 192, 8, 3, 6, 3, 3, 2, 2, 1>(a, s);
      return umma::launch_umma<192, 192, 8, 3, 6, 3, 3, 2, 2, 0>(a, s);
    }
    if (a.K <= 8) {
      if (lanes == 2) return umma::launch_umma<192, 192, 8, 2, 8, 3, 3, 2, 2>(a, s);
      return umma::launch_umma<192, 192, 8, 3, 8, 3, 3, 2, 2>(a, s);
    }
    // 9 .. 16 slots: 16 slot columns in every operand and accumulator; the slot-sized buffers double, which leaves the
    // ring five half-tile slots (two k, three v) and one update stream
    if (a.K <= 12) return umma::launch_umma<192, 192, 8, 3, 12, 2, 3, 2, 1>(a, s);
    return umma::launch_umma<192, 192, 8, 3, 16, 2, 3, 2, 1>(a, s);
  }
  if (a.D == 64 && a.H == 128) {  // the "Slot-Attention (small)" configuration
    if (a.K <= 8) {
      if (a.B >= 48 && a.lanes != 3) return umma::launch_umma<64, 128, 4, 2, 8, 6, 6, 2, 1>(a, s);
      return umma::launch_umma<64, 128, 8, 3, 8, 6, 6, 2, 1>(a, s);
    }
    if (a.B >= 48 && a.lanes != 3) return umma::launch_umma<64, 128, 4, 2, 16, 6, 6, 2, 1>(a, s);
    return umma::launch_umma<64, 128, 8, 3, 16, 6, 6, 2, 1>(a, s);
  }
  set_error("sa_iter_fwd(tcgen05): D=%d H=%d not instantiated", a.D, a.H);
  return OCRL_E_SHAPE;
}

}  // namespace ocrl

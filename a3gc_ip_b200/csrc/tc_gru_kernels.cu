// tcgen05 / TMEM engine for the graph-GRU recurrent layers (G_GRU_cell, net_aagc.py:343-368, looped as in :547-568).
//
//   msg = P^T-mix (h Wg^T)                      the node mix acts on the node index, the GEMM on the feature index: they commute,
//   r = sig(Wri x + br + Wrh msg)               so  W_gh msg = (P^T-mix h) (W_gh Wg)^T  for g in {r, u, c}.  With the fused message
//   u = sig(Wui x + bu + Wuh msg)               weights W'_g = W_gh Wg (computed once per launch, fp64 accumulation) and the MIXED
//   c = tanh(Wci x + bc + r * (Wch msg))        state h~ = P^T-mix h in the operand image, a step is ONE recurrent GEMM
//   h' = u h + (1 - u) c                            D[128, 256] = x[128, F] * [Wri; Wui; Wci]^T   (N = 192, cols r | u | cx)
//   returns (h', h'): no output activation (:368)              + h~[128, H] * [W'r; W'u]^T        (N = 128, cols r | u)
//                                                              ,  h~[128, H] * W'c^T              (N = 64,  cols ch)
//   instead of two dependent GEMMs with a state exchange after each (h -> msg -> h').
//
// Same decomposition as the LSTM-family kernel (tc_kernels.cu): a CTA owns 8 sequences (128 accumulator rows) x 64
// hidden units of one direction for all T steps, a cluster of C = H/64 CTAs covers the hidden dimension, the operand image of
// the mixed state lives in shared memory and is all-gathered through DSMEM once per step.  The x part of step t+1 is
// issued behind the recurrent part of step t (ping-pong accumulator buffers), i.e. while the epilogue warps run the gate
// math of step t and the exchange is in flight.  The gate math runs in the TMEM-native layout (thread = row, 16 units); the
// new state h' stays in fp32 registers, and its node mix -- 16 x 16 per sequence on the warp-level tensor path, fp16 hi/lo
// split, exactly like the gate mix of the LSTM kernel -- produces the next operand.
#include "common.cuh"
#include "tc_ptx.cuh"
#include <cstdlib>
#include <cstring>

namespace a3gc {
namespace {

constexpr int kRows = 128;
constexpr int kSeqTile = 8;
constexpr int kEpiWarps = 16;
constexpr int kEpiThreads = 32 * kEpiWarps;
constexpr int kProducers = 3;           // warp 0 and the two warps behind the epilogue warps (see tc_kernels.cu)
constexpr int kThreadsTC = 64 + kEpiThreads + 32 * (kProducers - 1);
constexpr int kMaxStages = 8;
constexpr int kWstFloats = 8 * 32;

// optional per-phase timeline of CTA (0,0) (A3GC_TC_TRACE=gru): [role 0 = epilogue, 1 = mma][step < 16][slot < 16] clock64
__device__ unsigned long long g_gru_trace[2][16][16];
#define GRU_TRACE(role, slot)                                                                \
  do {                                                                                       \
    if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && t < 16) g_gru_trace[role][t][slot] = clock64(); \
  } while (0)

struct GruDir {
  const uint16_t* wx_img;   // [C][F/16][NP][2][192][8]   rows 64*g + unit of dense_{r,u,c}_in.weight
  const uint16_t* wm_img;   // [C][H/16][NP][2][192][8]   rows 64*g + unit of W'_g = dense_{r,u,c}_hid.weight @ gcn_kernel
  const float* P;           // [16][16] zero padded, msg = P u  (P[n][m] = adjacency[m][n], net_aagc.py:348)
  const float* bias3;       // [H][4]   (b_r, b_u, b_c, 0)
  const float* h0; float* hT;
  int reverse;
};
struct GruLayerParams {
  GruDir d[2];
  const uint16_t* x_img;    // [tiles][T][F/16][NP][2][128][8]
  float* y; int64_t syb, syt, yld;
  uint16_t* y_img; int y_kf;
  int B, T, F, H, C, S;
  int acoll;                // A-operand collector reuse (see ptx::umma_f16_coll)
  int nprod;                // bulk-copy producer threads (1..3)
  int trace;
  int xprefetch;            // L2 prefetch of the next step's x image
};

enum { BAR_FULL = 0, BAR_EMPTY = kMaxStages, BAR_ACC_FULL = 2 * kMaxStages, BAR_ACC_EMPTY = BAR_ACC_FULL + 2, BAR_H = BAR_ACC_EMPTY + 2,
       BAR_HFREE, BAR_COUNT };

// fp32-split mode: no separate transposition buffers (each warp borrows 2 x 512 bytes of the state image, see `wst`)
__host__ __device__ inline size_t gru_fixed_smem_bytes(bool split) {
  return (split ? (size_t)0 : (size_t)kEpiWarps * kWstFloats * 4) + 256 * 4 + 256 * 4 + 32 * 8 + 16;   // wst, bias, Pfrag, barriers, tmem slot
}

__device__ __forceinline__ uint32_t img_off(int k, int row) {
  return (uint32_t)(((k >> 3) * kRows + row) * 16 + (k & 7) * 2);
}

template <bool SPLIT>
__global__ void __launch_bounds__(kThreadsTC, 1)
tc_gru_layer_kernel(const GruLayerParams p) {
  constexpr int NP = SPLIT ? 2 : 1;
  constexpr uint32_t kXB = NP * 2 * 192 * 16;       // one K=16 block of [Wri; Wui; Wci] (all parts)
  constexpr uint32_t kXA = NP * 2 * 128 * 16;       // one K=16 block of x rows
  constexpr uint32_t kMB = NP * 2 * 192 * 16;       // one K=16 block of [W'r; W'u; W'c]
  constexpr uint32_t kStageBytes = 2 * kMB;          // ring slot: 2 recurrent blocks >= 1 x block (B + A)
  constexpr uint32_t kHBlock = 8 * kRows * 16;
  extern __shared__ __align__(1024) uint8_t smem[];

  const int C = p.C, S = p.S, H = p.H, F = p.F, T = p.T;
  const uint32_t c = C > 1 ? ptx::cluster_ctarank() : 0u;
  const int tile = blockIdx.x / C;
  const GruDir& d = p.d[blockIdx.y];
  const int KF = F / 16, KH = H / 16;
  const int warp = threadIdx.x >> 5;

  uint8_t* hbuf = smem;                                              // operand image of the mixed state h~: [NP][H/8][128][8]
  uint8_t* ring = hbuf + (size_t)NP * H * 256;
  float* staging = reinterpret_cast<float*>(ring + (size_t)S * kStageBytes);
  float* biasg = staging + (SPLIT ? 0 : kEpiWarps * kWstFloats);     // [3][64]  (+ 64 pad)
  uint32_t* Pfrag = reinterpret_cast<uint32_t*>(biasg + 256);        // [hi, lo][32 lanes][4 regs]
  uint64_t* bars = reinterpret_cast<uint64_t*>(Pfrag + 256);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 32);

  if (threadIdx.x == 0) {
    for (int i = 0; i < BAR_COUNT; ++i)
      ptx::mbar_init(&bars[i], i == BAR_HFREE ? (uint32_t)C : (i >= BAR_FULL && i < BAR_FULL + kMaxStages) ? 2u : 1u);
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, 512);
  {
    uint4* z = reinterpret_cast<uint4*>(hbuf);
    const int n16 = NP * H * 16;
    for (int i = threadIdx.x; i < n16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < 192; i += blockDim.x) biasg[i] = d.bias3[(size_t)c * 256 + (i & 63) * 4 + (i >> 6)];
    for (int i = threadIdx.x; i < 128; i += blockDim.x) {
      const int r = i & 3, l = i >> 2;
      const int mm = (l >> 2) + 8 * (r & 1), nn = 2 * (l & 3) + 8 * (r >> 1);
      uint32_t hi, lo;
      ptx::split_pair_f16(d.P[mm * 16 + nn], d.P[mm * 16 + nn + 1], hi, lo);
      Pfrag[(0 * 32 + l) * 4 + r] = hi;
      Pfrag[(1 * 32 + l) * 4 + r] = lo;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (C > 1) ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint16_t cta_mask = (uint16_t)((1u << C) - 1u);

  if (warp == 0 || warp >= 2 + kEpiWarps) {
    // ================================================================ producers (same scheme as tc_kernels.cu: the weight (B)
    // copy of stage i is issued by producer i % 2, the x-image (A) copy of every x stage by the third producer; every FULL
    // barrier takes two arrivals.  p.nprod = 1 / 2: one / two threads issue both copies of their stages.)
    // (whole warp, elected lane issues: see the MMA warp)
    const uint32_t my = __shfl_sync(0xffffffffu, warp == 0 ? 0u : (uint32_t)(warp - (1 + kEpiWarps)), 0);
    const uint32_t nprod = (uint32_t)p.nprod;
    const bool split_a = nprod == 3;
    const uint32_t nb = split_a ? 2u : nprod;
    if (my < nprod) {
      uint32_t st = 0, ph = 0, turn = 0;
      const bool tr_on = p.trace && blockIdx.x == 0 && blockIdx.y == 0;
      unsigned long long w_empty = 0, t_begin = tr_on ? clock64() : 0;      // trace: cycles this producer waited for a free slot
      auto load_stage = [&](const void* bsrc, uint32_t bbytes, uint32_t aoff, const void* asrc, uint32_t abytes) {
        uint8_t* dst = ring + st * kStageBytes;
        const bool b_side = my < nb && turn == my, a_side = split_a ? my == 2u : b_side;
        if (b_side || a_side) {
          const unsigned long long t0 = tr_on ? clock64() : 0;
          ptx::mbar_wait(&bars[BAR_EMPTY + st], ph ^ 1u);
          if (tr_on) w_empty += clock64() - t0;
        }
        if ((b_side || a_side) && ptx::elect_one()) {
          if (b_side) {
            ptx::mbar_arrive_expect_tx(&bars[BAR_FULL + st], bbytes);
            ptx::bulk_g2s(dst, bsrc, bbytes, &bars[BAR_FULL + st]);
          }
          if (a_side) {
            if (abytes) {
              ptx::mbar_arrive_expect_tx(&bars[BAR_FULL + st], abytes);
              ptx::bulk_g2s(dst + aoff, asrc, abytes, &bars[BAR_FULL + st]);
            } else {
              ptx::mbar_arrive(&bars[BAR_FULL + st]);
            }
          }
        }
        __syncwarp();
        if (++turn == nb) turn = 0;
        if (++st == (uint32_t)S) { st = 0; ph ^= 1u; }
      };
      const uint8_t* wx = reinterpret_cast<const uint8_t*>(d.wx_img) + (size_t)c * KF * kXB;
      const uint8_t* wm = reinterpret_cast<const uint8_t*>(d.wm_img) + (size_t)c * KH * kMB;
      const uint8_t* xi = reinterpret_cast<const uint8_t*>(p.x_img);
      auto xblocks = [&](int t, int kb0, int kb1) {
        const int ta = d.reverse ? T - 1 - t : t;
        const uint8_t* xs = xi + ((size_t)tile * T + ta) * KF * kXA;
        // the x operand image comes from HBM: this CTA's share of the NEXT step's image is pulled into L2 one step ahead
        if (p.xprefetch && my == 0 && t + 1 < T) {
          const int tn = d.reverse ? T - 2 - t : t + 1;
          const uint32_t share = (uint32_t)(KF / C) * kXA;
          if (ptx::elect_one()) ptx::bulk_prefetch_l2(xi + ((size_t)tile * T + tn) * KF * kXA + (size_t)c * share, share);
          __syncwarp();
        }
        for (int kb = kb0; kb < kb1; ++kb) load_stage(wx + (size_t)kb * kXB, kXB, kXB, xs + (size_t)kb * kXA, kXA);
      };
      xblocks(0, 0, KF);
      for (int t = 0; t < T; ++t) {
        for (int s2 = 0; s2 < KH / 2; ++s2) load_stage(wm + (size_t)s2 * 2 * kMB, 2 * kMB, 0, nullptr, 0);
        if (t + 1 < T) xblocks(t + 1, 0, KF);
      }
      if (tr_on && (threadIdx.x & 31) == 0) { g_gru_trace[0][15][2 * my] = w_empty; g_gru_trace[0][15][2 * my + 1] = clock64() - t_begin; }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer: the whole warp walks the stage sequence
    // (uniform control flow and operands), one elected lane issues the MMAs and commits of a stage (see ptx::elect_one)
    {
      const bool lane0 = (threadIdx.x & 31) == 0;
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
      const uint32_t idesc192 = ptx::make_idesc_f16(128, 192, !SPLIT);
      const uint32_t idesc128 = ptx::make_idesc_f16(128, 128, !SPLIT);
      const uint32_t idesc64 = ptx::make_idesc_f16(128, 64, !SPLIT);
      const uint32_t hbase = ptx::smem_u32(hbuf);
      const uint32_t hpart = (uint32_t)H * 256;
      uint32_t st = 0, ph = 0;
      uint32_t empty_k[2] = {0, 0};
      const uint32_t ring_addr = ptx::smem_u32(ring);
      const bool tr_on = p.trace && blockIdx.x == 0 && blockIdx.y == 0;
      unsigned long long w_full = 0, w_h = 0, t_begin = tr_on ? clock64() : 0;   // trace: cycles waited for a full slot / for BAR_H
      auto wait_stage = [&]() -> uint32_t {
        const unsigned long long t0 = tr_on ? clock64() : 0;
        ptx::mbar_wait(&bars[BAR_FULL + st], ph);
        if (tr_on) w_full += clock64() - t0;
        ptx::tc_fence_after();
        return ring_addr + st * kStageBytes;
      };
      auto next_stage = [&]() { if (++st == (uint32_t)S) { st = 0; ph ^= 1u; } };
      const uint64_t dA = ptx::make_smem_desc(0, kRows * 16, 128);
      const uint64_t dB192 = ptx::make_smem_desc(0, 192 * 16, 128);
      const bool acoll = SPLIT && p.acoll;
      // one K=16 block: D[:, dcol..dcol+N) (+)= A * B^T with the split passes hi*hi + hi*lo + lo*hi  (elected lane only)
      auto block_mma = [&](uint32_t dcol, uint32_t a0, uint32_t astride, uint32_t b0, uint32_t bstride, uint64_t dB, uint32_t idesc, bool first) {
        const uint64_t ah = dA + ((a0 & 0x3FFFFu) >> 4), bh = dB + ((b0 & 0x3FFFFu) >> 4);
        if (acoll) {
          ptx::umma_f16_coll(tmem_u + dcol, ah, bh, idesc, first ? 0u : 1u, 1);
          ptx::umma_f16_coll(tmem_u + dcol, ah, bh + (bstride >> 4), idesc, 1u, 3);
          ptx::umma_f16(tmem_u + dcol, ah + (astride >> 4), bh, idesc, 1u);
        } else {
          ptx::umma_f16(tmem_u + dcol, ah, bh, idesc, first ? 0u : 1u);
          if (SPLIT) {
            ptx::umma_f16(tmem_u + dcol, ah + (astride >> 4), bh, idesc, 1u);
            ptx::umma_f16(tmem_u + dcol, ah, bh + (bstride >> 4), idesc, 1u);
          }
        }
      };
      auto xblocks = [&](uint32_t dcol, int kb0, int kb1) {
        for (int kb = kb0; kb < kb1; ++kb) {
          const uint32_t sa = wait_stage();
          if (ptx::elect_one()) {
            block_mma(dcol, sa + kXB, kXA / NP, sa, kXB / NP, dB192, idesc192, kb == 0);
            ptx::umma_commit(&bars[BAR_EMPTY + st]);
          }
          __syncwarp();
          next_stage();
        }
      };
      xblocks(0, 0, KF);
      for (int t = 0; t < T; ++t) {
        const uint32_t b = t & 1, bo = b ^ 1u;
        const bool nx = t + 1 < T;
        // recurrent part: r | u accumulate onto the x part, ch starts fresh in columns [192,256)
        if (lane0) GRU_TRACE(1, 0);
        {
          const unsigned long long t0 = tr_on ? clock64() : 0;
          ptx::mbar_wait(&bars[BAR_H], t & 1);         // the mixed state of every chunk is in the local image
          if (tr_on) w_h += clock64() - t0;
        }
        ptx::tc_fence_after();
        if (lane0) GRU_TRACE(1, 1);
        for (int s2 = 0; s2 < KH / 2; ++s2) {
          const uint32_t sa = wait_stage();
          if (ptx::elect_one()) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const int kb = s2 * 2 + j;
              const uint32_t a0 = hbase + (uint32_t)kb * 2 * kRows * 16, b0 = sa + (uint32_t)j * kMB;
              if (acoll) {
                // one A tile (h~ hi, then h~ lo) against r|u (N = 128) and ch (N = 64), hi and lo parts of the weights: the A
                // operand is read from shared memory once per part and kept in the collector buffer for the other MMAs
                const uint64_t ah = dA + ((a0 & 0x3FFFFu) >> 4), al = ah + (hpart >> 4);
                const uint64_t bh = dB192 + ((b0 & 0x3FFFFu) >> 4), bl = bh + ((kMB / NP) >> 4);
                const uint64_t ch = bh + ((128u * 16u) >> 4), cl = bl + ((128u * 16u) >> 4);
                const uint32_t first = (s2 | j) == 0 ? 0u : 1u;
                ptx::umma_f16_coll(tmem_u + b * 256, ah, bh, idesc128, 1u, 1);
                ptx::umma_f16_coll(tmem_u + b * 256, ah, bl, idesc128, 1u, 2);
                ptx::umma_f16_coll(tmem_u + b * 256 + 192, ah, ch, idesc64, first, 2);
                ptx::umma_f16_coll(tmem_u + b * 256 + 192, ah, cl, idesc64, 1u, 3);
                ptx::umma_f16_coll(tmem_u + b * 256, al, bh, idesc128, 1u, 1);
                ptx::umma_f16_coll(tmem_u + b * 256 + 192, al, ch, idesc64, 1u, 3);
              } else {
                block_mma(b * 256, a0, hpart, b0, kMB / NP, dB192, idesc128, false);
                block_mma(b * 256 + 192, a0, hpart, b0 + 128 * 16, kMB / NP, dB192, idesc64, (s2 | j) == 0);
              }
            }
            ptx::umma_commit(&bars[BAR_EMPTY + st]);
            if (s2 == KH / 2 - 1) {
              ptx::umma_commit(&bars[BAR_ACC_FULL + b]);
              if (C > 1) ptx::umma_commit_multicast(&bars[BAR_HFREE], cta_mask); else ptx::umma_commit(&bars[BAR_HFREE]);
            }
          }
          __syncwarp();
          next_stage();
        }
        if (lane0) GRU_TRACE(1, 2);
        // x part of step t+1 -> columns [0,192) of the other buffer (free once the gate math of step t-1 has drained it)
        if (nx && t >= 1) {
          ptx::mbar_wait(&bars[BAR_ACC_EMPTY + bo], empty_k[bo] & 1u);
          empty_k[bo] += 1;
          ptx::tc_fence_after();
        }
        if (lane0) GRU_TRACE(1, 3);
        if (nx) xblocks(bo * 256, 0, KF);
        if (lane0) GRU_TRACE(1, 4);
      }
      if (tr_on && lane0) { g_gru_trace[1][15][0] = w_full; g_gru_trace[1][15][1] = w_h; g_gru_trace[1][15][2] = clock64() - t_begin; }
    }
  } else {
    // ================================================================ epilogue warps
    const int et = threadIdx.x - 64;
    const int ew = et >> 5, lane = et & 31;
    const int qd = warp & 3;                         // TMEM lane quarter: rows 32qd .. 32qd+31 (sequences 2qd, 2qd+1)
    const int ug = ew >> 2;
    const int ubase = 16 * ug;                       // this warp's 16 units of the CTA's 64
    const int tq = lane >> 2, tr = lane & 3;
    // Warp-private transposition buffer (8 columns x 32 rows fp32, as two 512-byte halves: columns 0..3 at wst, 4..7 at wst2).
    // bf16 mode: a separate allocation.  fp32-split mode: rows 32qd..32qd+31 of K chunk 2ug+1 of the CTA's own block of the
    // state image, hi and lo part -- the bytes that this warp itself rewrites last in mix_store (unit block 1), and that
    // nobody reads between BAR_HFREE (every MMA on h~_{t-1} has completed, so every outgoing copy has landed) and the next
    // publish.  The 16 KB this saves are a fourth ring slot at H = 256.
    float* wst = staging + ew * kWstFloats;
    float* wst2 = wst + 4 * 32;
    if (SPLIT) {
      wst = reinterpret_cast<float*>(hbuf + ((size_t)((int)c * 8 + 2 * ug + 1) * kRows + 32 * qd) * 16);
      wst2 = wst + (size_t)H * 64;
    }
    const float* wsp = (tq < 4 ? wst : wst2) + (tq & 3) * 32;        // column tq (B-fragment reads)
    const uint32_t tmem_row = tmem + ((uint32_t)(qd * 32) << 16);
    const int ycol = blockIdx.y * H + (int)c * 64;
    // gate-math ownership (TMEM-native): this thread = accumulator row `row`, units ubase .. ubase+15
    const int row = 32 * qd + lane;
    const int rseq = tile * kSeqTile + (row >> 4), rnode = row & 15;
    const bool rvalid = rseq < p.B && rnode < kNodes;
    float hreg[16];
#pragma unroll
    for (int i = 0; i < 16; ++i)
      hreg[i] = (rvalid && d.h0 != nullptr) ? d.h0[((size_t)rseq * kNodes + rnode) * H + c * 64 + ubase + i] : 0.f;

    auto publish_block = [&](int bar, int acc_empty, bool tmem_read) {
      ptx::fence_proxy_async();
      if (tmem_read) ptx::tc_fence_before();
      ptx::named_bar_sync(1, kEpiThreads);
      if (et == 0) {
        if (acc_empty >= 0) ptx::mbar_arrive(&bars[BAR_ACC_EMPTY + acc_empty]);
        // rotated send order (c+1, c+2, ...): every receiver is served by one sender at a time instead of all senders
        // pushing to CTA 0 first
        for (uint32_t i = 1; i < (uint32_t)C; ++i) {
          const uint32_t peer = (c + i) % (uint32_t)C;
          for (int part = 0; part < NP; ++part)
            ptx::bulk_s2remote(hbuf + (size_t)part * H * 256 + (size_t)c * kHBlock, kHBlock, &bars[bar], peer);
        }
        ptx::mbar_arrive_expect_tx(&bars[bar], (uint32_t)(C - 1) * NP * kHBlock);
      }
    };
    const size_t img_step = (size_t)p.y_kf * NP * 4096;
    const size_t img_base = ((size_t)tile * T * p.y_kf + ((ycol + ubase) >> 4)) * NP * 4096 + (size_t)row * 16;
    auto emit = [&](int ta, const float (&v)[16]) {
      if (p.y != nullptr && rvalid) {
        float4* yp = reinterpret_cast<float4*>(p.y + (size_t)rseq * p.syb + (size_t)ta * p.syt + (size_t)rnode * p.yld + ycol + ubase);
#pragma unroll
        for (int j = 0; j < 4; ++j) yp[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
      if (p.y_img != nullptr) {
        uint8_t* ip = reinterpret_cast<uint8_t*>(p.y_img) + img_base + (size_t)ta * img_step;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t hw[4], lw[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (SPLIT) {
              ptx::split_pair_f16(v[hf * 8 + 2 * j], v[hf * 8 + 2 * j + 1], hw[j], lw[j]);
            } else {
              const __nv_bfloat162 bb = __floats2bfloat162_rn(v[hf * 8 + 2 * j], v[hf * 8 + 2 * j + 1]);
              hw[j] = *reinterpret_cast<const uint32_t*>(&bb);
              lw[j] = 0;
            }
          }
          *reinterpret_cast<uint4*>(ip + hf * 2048) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          if (SPLIT) *reinterpret_cast<uint4*>(ip + hf * 2048 + 4096) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
      }
    };

    const uint4* Pfrag4 = reinterpret_cast<const uint4*>(Pfrag);
    const int swz = 8 * (tq & 3);
    // h~ = P h' (16 x 16 node mix per sequence) of this warp's 32 rows x 16 units -> local operand image.  The values go
    // registers (lane = row) -> warp-private transposition buffer -> B fragments of mma.m16n8k16 (k = node), fp16 hi/lo
    // split, 3 passes; the C fragment (nodes tq, tq+8 x units 2tr, 2tr+1) is written as packed pairs.
    auto mix_store = [&](const float (&h)[16]) {
      const uint4 ah4 = Pfrag4[0 * 32 + lane], al4 = Pfrag4[1 * 32 + lane];
      const uint32_t ah[4] = {ah4.x, ah4.y, ah4.z, ah4.w}, al[4] = {al4.x, al4.y, al4.z, al4.w};
#pragma unroll
      for (int ub = 0; ub < 2; ++ub) {
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) (j < 4 ? wst : wst2)[(j & 3) * 32 + (lane ^ (8 * (j & 3)))] = h[8 * ub + j];
        __syncwarp();
        const int k = (int)c * 64 + ubase + 8 * ub + 2 * tr;
        float2 u0[2], u1[2];
#pragma unroll
        for (int sq = 0; sq < 2; ++sq) {
          u0[sq] = *reinterpret_cast<const float2*>(wsp + ((16 * sq + 2 * tr) ^ swz));
          u1[sq] = *reinterpret_cast<const float2*>(wsp + ((16 * sq + 2 * tr + 8) ^ swz));
        }
        if (SPLIT && ub == 1) __syncwarp();          // the results of unit block 1 go to the bytes the buffer borrows
#pragma unroll
        for (int sq = 0; sq < 2; ++sq) {
          float z[4] = {0.f, 0.f, 0.f, 0.f};
          if (SPLIT) {
            uint32_t bh0, bl0, bh1, bl1;
            ptx::split_pair_f16(u0[sq].x, u0[sq].y, bh0, bl0);
            ptx::split_pair_f16(u1[sq].x, u1[sq].y, bh1, bl1);
            float z2[4] = {0.f, 0.f, 0.f, 0.f}, z3[4] = {0.f, 0.f, 0.f, 0.f};
            ptx::mma_16816_f16(z, ah, bh0, bh1);
            ptx::mma_16816_f16(z2, al, bh0, bh1);
            ptx::mma_16816_f16(z3, ah, bl0, bl1);
#pragma unroll
            for (int j = 0; j < 4; ++j) z[j] += z2[j] + z3[j];
          } else {
            const __half2 h0 = __floats2half2_rn(u0[sq].x, u0[sq].y), h1 = __floats2half2_rn(u1[sq].x, u1[sq].y);
            ptx::mma_16816_f16(z, ah, *reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
          }
#pragma unroll
          for (int up = 0; up < 2; ++up) {
            uint32_t hi, lo = 0;
            if (SPLIT) {
              ptx::split_pair_f16(z[2 * up], z[2 * up + 1], hi, lo);
            } else {
              const __nv_bfloat162 bb = __floats2bfloat162_rn(z[2 * up], z[2 * up + 1]);
              hi = *reinterpret_cast<const uint32_t*>(&bb);
            }
            const uint32_t off = img_off(k, 16 * (2 * qd + sq) + tq + 8 * up);
            *reinterpret_cast<uint32_t*>(hbuf + off) = hi;
            if (SPLIT) *reinterpret_cast<uint32_t*>(hbuf + (size_t)H * 256 + off) = lo;
          }
        }
      }
    };

    mix_store(hreg);                                 // h~_{-1} = P h0
    publish_block(BAR_H, -1, false);

    for (int t = 0; t < T; ++t) {
      const uint32_t b = t & 1;
      const int ta = d.reverse ? T - 1 - t : t;
      // ---------------------------------------------------------------- gates (TMEM-native layout: thread = row)
      if (et == 0) GRU_TRACE(0, 0);
      ptx::mbar_wait(&bars[BAR_ACC_FULL + b], (t >> 1) & 1);
      if (et == 0) GRU_TRACE(0, 1);
      ptx::mbar_wait(&bars[BAR_HFREE], t & 1);       // every CTA has finished reading h~_{t-1}: the image may be rewritten
      ptx::tc_fence_after();
      if (et == 0) GRU_TRACE(0, 2);
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        float zr[8], zu[8], zx[8], zh[8];
        const uint32_t col = tmem_row + b * 256 + ubase + 8 * hf;
        ptx::tmem_ld8(col, zr);
        ptx::tmem_ld8(col + 64, zu);
        ptx::tmem_ld8(col + 128, zx);
        ptx::tmem_ld8(col + 192, zh);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int un = ubase + 8 * hf + i;
          const float r = fast_sigmoid(zr[i] + biasg[un]);
          const float u = fast_sigmoid(zu[i] + biasg[64 + un]);
          const float cc = fast_tanh(fmaf(r, zh[i], zx[i] + biasg[128 + un]));
          const float hn = fmaf(u, hreg[hf * 8 + i] - cc, cc);            // u h + (1 - u) c   (net_aagc.py:364)
          hreg[hf * 8 + i] = rvalid ? hn : 0.f;
        }
      }
      if (et == 0) GRU_TRACE(0, 3);
      mix_store(hreg);
      if (et == 0) GRU_TRACE(0, 4);
      publish_block(BAR_H, (int)b, true);
      if (et == 0) GRU_TRACE(0, 5);
      emit(ta, hreg);
      if (et == 0) GRU_TRACE(0, 6);
    }
    if (rvalid && d.hT != nullptr) {
#pragma unroll
      for (int i = 0; i < 16; ++i) d.hT[((size_t)rseq * kNodes + rnode) * H + c * 64 + ubase + i] = hreg[i];
    }
    ptx::mbar_wait(&bars[BAR_H], T & 1);             // the last publish must have landed everywhere before any CTA exits
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (C > 1) ptx::cluster_sync_all();
  if (warp == 1) ptx::tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------ packing
__device__ __forceinline__ uint16_t part_bits(float v, int part, bool split) {
  if (!split) return __bfloat16_as_ushort(__float2bfloat16_rn(v));
  const __half hi = __float2half_rn(v);
  if (part == 0) return __half_as_ushort(hi);
  return __half_as_ushort(__float2half_rn(v - __half2float(hi)));
}

struct GruPackedTc { uint16_t* wx_img; uint16_t* wm_img; float* wfused; float* P; float* bias3; };

// fused message weights W'_g = dense_g_hid.weight @ gcn_kernel  (g = r, u, c; [H, H] each), fp64 accumulation:
// dense_g_hid(P^T-mix(h gcn_kernel^T)) = (P^T-mix h) W'_g^T   (net_aagc.py:347-348, :353-361)
__global__ void tc_gru_fuse_kernel(a3gc_cell_params cp, float* __restrict__ wfused, int H) {
  const int64_t n = (int64_t)3 * H * H;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % H), j = (int)((i / H) % H), g = (int)(i / ((int64_t)H * H));
    const float* wh = cp.dense_hid_w[g] + (size_t)j * H;
    double acc = 0.0;
    for (int m = 0; m < H; ++m) acc += (double)wh[m] * (double)cp.g_gcn_kernel[(size_t)m * H + k];
    wfused[i] = (float)acc;
  }
}

__global__ void tc_pack_gru_kernel(a3gc_cell_params cp, GruPackedTc out, int F, int H, int split) {
  const int NP = split ? 2 : 1;
  const int KF = F / 16, KH = H / 16, C = H / 64;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // 192-row images: index = ((((c*KB + kb)*NP + part)*2 + kc)*192 + r)*8 + e,  r = 64*g + unit
  for (int which = 0; which < 2; ++which) {
    const int KB = which == 0 ? KF : KH, K = which == 0 ? F : H;
    uint16_t* img = which == 0 ? out.wx_img : out.wm_img;
    const int64_t n = (int64_t)C * KB * NP * 2 * 192 * 8;
    for (int64_t i = tid; i < n; i += stride) {
      const int e = (int)(i & 7);
      int64_t rest = i >> 3;
      const int r = (int)(rest % 192); rest /= 192;
      const int kc = (int)(rest & 1); rest >>= 1;
      const int part = (int)(rest % NP); rest /= NP;
      const int kb = (int)(rest % KB); const int c = (int)(rest / KB);
      const int g = r >> 6, j = c * 64 + (r & 63), k = kb * 16 + kc * 8 + e;
      const float* w = which == 0 ? cp.dense_in_w[g] : out.wfused + (size_t)g * H * H;
      img[i] = part_bits(w[(size_t)j * K + k], part, split);
    }
  }
  for (int64_t i = tid; i < 256; i += stride) {
    const int n = (int)(i / 16), m = (int)(i % 16);
    out.P[i] = (m < kNodes && n < kNodes) ? cp.g_adjacency[m * kNodes + n] : 0.f;     // used transposed (net_aagc.py:348)
  }
  for (int64_t i = tid; i < (int64_t)H * 4; i += stride) {
    const int g = (int)(i & 3);
    out.bias3[i] = g < 3 ? cp.dense_in_b[g][i >> 2] : 0.f;
  }
}

size_t gru_dir_bytes(int F, int H, int NP) {
  size_t b = 0;
  b += align_up((size_t)F * 192 * (H / 64) * NP * 2, 256);
  b += align_up((size_t)H * 192 * (H / 64) * NP * 2, 256);
  b += align_up((size_t)3 * H * H * 4, 256);                 // fused message weights W' (fp32)
  b += align_up((size_t)(256 + 4 * H) * 4, 256);
  return b;
}

GruPackedTc gru_carve(char* base, int F, int H, int NP) {
  GruPackedTc p;
  p.wx_img = reinterpret_cast<uint16_t*>(base); base += align_up((size_t)F * 192 * (H / 64) * NP * 2, 256);
  p.wm_img = reinterpret_cast<uint16_t*>(base); base += align_up((size_t)H * 192 * (H / 64) * NP * 2, 256);
  p.wfused = reinterpret_cast<float*>(base); base += align_up((size_t)3 * H * H * 4, 256);
  float* f = reinterpret_cast<float*>(base);
  p.P = f; p.bias3 = f + 256;
  return p;
}

}  // namespace

size_t tc_gru_weights_bytes(int f_in, int hidden, int num_dirs, int precision) {
  return (size_t)num_dirs * gru_dir_bytes(f_in, hidden, precision == A3GC_PREC_FP32 ? 2 : 1);
}

// fused message weights + operand images + mix matrix + biases of every direction -> weights_ws (tc_gru_weights_bytes)
int tc_gru_pack_weights(int num_dirs, const a3gc_cell_params* cells, int f_in, int hidden, int precision, char* wbase, cudaStream_t stream) {
  const bool split = precision == A3GC_PREC_FP32;
  const int NP = split ? 2 : 1;
  const size_t dir_bytes = gru_dir_bytes(f_in, hidden, NP);
  for (int d = 0; d < num_dirs; ++d) {
    GruPackedTc pk = gru_carve(wbase + d * dir_bytes, f_in, hidden, NP);
    tc_gru_fuse_kernel<<<148 * 2, 256, 0, stream>>>(cells[d], pk.wfused, hidden);
    A3GC_LAUNCH_CHECK("tc_gru_fuse_kernel");
    tc_pack_gru_kernel<<<148, 256, 0, stream>>>(cells[d], pk, f_in, hidden, split ? 1 : 0);
    A3GC_LAUNCH_CHECK("tc_pack_gru_kernel");
  }
  return A3GC_OK;
}

// G-GRU layer on the tensor-core engine; a.x_img must hold the packed input (tc_layer_forward packs it if needed)
int tc_gru_layer_launch(const LayerArgs& a, const uint16_t* x_img, char* wbase, cudaStream_t stream) {
  const int F = a.f_in, H = a.hidden;
  const bool split = a.precision == A3GC_PREC_FP32;
  const int NP = split ? 2 : 1;
  const int C = H / 64;
  const int64_t tiles = (a.batch + kSeqTile - 1) / kSeqTile;
  GruLayerParams p;
  memset(&p, 0, sizeof(p));
  const size_t dir_bytes = gru_dir_bytes(F, H, NP);
  if (a.packed == nullptr) {
    int rc = tc_gru_pack_weights(a.num_dirs, a.cells, F, H, a.precision, wbase, stream);
    if (rc) return rc;
  } else {
    wbase = const_cast<char*>(static_cast<const char*>(a.packed));
  }
  for (int d = 0; d < a.num_dirs; ++d) {
    GruPackedTc pk = gru_carve(wbase + d * dir_bytes, F, H, NP);
    GruDir& g = p.d[d];
    g.wx_img = pk.wx_img; g.wm_img = pk.wm_img; g.P = pk.P; g.bias3 = pk.bias3;
    g.h0 = a.h0[d]; g.hT = a.hT[d]; g.reverse = a.reverse[d];
  }
  p.x_img = x_img;
  p.y = a.y; p.syb = a.y_stride_b; p.syt = a.y_stride_t; p.yld = a.y_ld;
  p.y_img = a.y_img; p.y_kf = a.y_img_f / 16;
  p.B = (int)a.batch; p.T = (int)a.steps; p.F = F; p.H = H; p.C = C;
  p.acoll = getenv("A3GC_TC_ACOLL") ? atoi(getenv("A3GC_TC_ACOLL")) : 1;
  p.nprod = getenv("A3GC_TC_NPROD") ? atoi(getenv("A3GC_TC_NPROD")) : kProducers;
  if (p.nprod < 1) p.nprod = 1;
  if (p.nprod > kProducers) p.nprod = kProducers;
  p.xprefetch = getenv("A3GC_TC_XPREFETCH") ? atoi(getenv("A3GC_TC_XPREFETCH")) : 1;
  p.trace = (getenv("A3GC_TC_TRACE") != nullptr && strcmp(getenv("A3GC_TC_TRACE"), "gru") == 0) ? 1 : 0;
  int dev = 0, smem_max = 0;
  A3GC_CUDA_TRY(cudaGetDevice(&dev));
  A3GC_CUDA_TRY(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  const size_t stage_bytes = (size_t)2 * NP * 2 * 192 * 16;
  const size_t fixed = (size_t)NP * H * 256 + gru_fixed_smem_bytes(split);
  int S = kMaxStages;
  if (const char* e = getenv("A3GC_TC_STAGES")) { const int v = atoi(e); if (v >= 2 && v <= kMaxStages) S = v; }
  while (S > 1 && fixed + (size_t)S * stage_bytes > (size_t)smem_max) --S;
  if (fixed + (size_t)S * stage_bytes > (size_t)smem_max || S < 2) {
    set_error("tc engine (G-GRU): shared memory budget exceeded (hidden=%d)", H);
    return A3GC_ERR_UNSUPPORTED;
  }
  p.S = S;
  const size_t smem = fixed + (size_t)S * stage_bytes;
  void (*kern)(const GruLayerParams) = split ? tc_gru_layer_kernel<true> : tc_gru_layer_kernel<false>;
  A3GC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(tiles * C), (unsigned)a.num_dirs, 1);
  cfg.blockDim = dim3(kThreadsTC, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  A3GC_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, p));
  A3GC_LAUNCH_CHECK("tc_gru_layer_kernel");
  return A3GC_OK;
}

// debug: the per-phase timeline of CTA (0,0) of the last traced G-GRU launch (see a3gc_debug_read_tc_trace)
int tc_gru_read_trace(unsigned long long* host_out) {
  A3GC_CUDA_TRY(cudaMemcpyFromSymbol(host_out, g_gru_trace, sizeof(unsigned long long) * 2 * 16 * 16));
  return A3GC_OK;
}

}  // namespace a3gc

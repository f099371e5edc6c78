// tcgen05 / TMEM engine for the LSTM-family recurrent layers (AAGC / A3GC / AGC cells,
// net_aagc.py:103-126, :178-217, :266-303 looped as in :435-456).
//
// Decomposition.  One CTA owns a batch tile of 8 sequences (8 x 16 padded nodes = the 128 rows of
// a UMMA M tile) and 64 hidden units (4 gates x 64 = 256 accumulator columns, col = 4*unit + gate)
// of ONE direction, for all T steps.  A thread-block cluster of C = H/64 CTAs covers the hidden
// dimension; after every step the CTAs all-gather their 64-unit slices of the new state through
// distributed shared memory (bulk smem->smem copies that complete on the peers' mbarriers).
//
// Per step and CTA (S = [x_t | h_{t-1}], rows = (sequence, node)):
//   gates  U[128,256] = S[128,F+H] * Wc[256,F+H]^T      tcgen05.mma, fp32 accumulate in TMEM;
//                                                        the x half for step t+1 is issued while the
//                                                        epilogue of step t runs (ping-pong TMEM buffers)
//   mix    z_g = P_g U_g  (15x15 adjacency per gate)     CUDA cores, after a TMEM->smem transposition,
//                                                        P_g rows held in registers
//   LSTM   c' = f c + i g,  hy = o tanh(c')              registers (c never leaves the register file)
//   attention (A3GC/AGC)  Wh*hy on the tensor core (N=64), q / Wq*q and the node reductions on CUDA cores
//   h' = hy (1 + a),  y_t = tanh(h')
// Weights never fit on chip (H=256: 3.9 MB per direction), so every step streams this CTA's slice
// from L2 through a shared-memory ring with 1-D bulk copies (TMA engine, SASS UBLKCP) that land in
// the exact UMMA operand image (K-major, no swizzle: [K/8][rows][8]) prepared once per launch.
//
// Precision.  A3GC_PREC_FP32: every operand is split into fp16 hi + lo (22 significand bits) and each
// product is accumulated as hi*hi + lo*hi + hi*lo (3 tensor passes, fp32 accumulate): measured error
// vs the CPU reference ~1e-6.  A3GC_PREC_BF16: one bf16 pass.
#include "common.cuh"
#include "tc_ptx.cuh"
#include <cstdlib>

namespace a3gc {
namespace {

constexpr int kRows = 128;             // 8 sequences x 16 node slots
constexpr int kSeqTile = 8;
constexpr int kUnits = 64;             // hidden units per CTA
constexpr int kEpiWarps = 16;          // warps 2..17: warp (qd = warp%4, ug = (warp-2)/4) owns rows 32qd..32qd+31 x units 16ug..16ug+15
constexpr int kEpiThreads = 32 * kEpiWarps;
constexpr int kProducers = 3;           // warp 0 and the two warps behind the epilogue warps (18, 19)
constexpr int kXchgWarp = 2 + kEpiWarps + (kProducers - 2);   // warp 19: the third producer, or the state exchange through L2 (p.xchg;
                                                               // a 21st warp would cut the register budget from 96 to 80 per thread)
constexpr int kThreadsTC = 64 + kEpiThreads + 32 * (kProducers - 1);
constexpr int kMaxStages = 8;
constexpr int kWstFloats = 8 * 32;     // warp-private transposition buffer: 8 accumulator columns x 32 rows

// optional per-phase timeline of CTA (0,0) (A3GC_TC_TRACE=1): [role 0 = epilogue, 1 = mma][step < 16][slot < 16] clock64
__device__ unsigned long long g_tc_trace[2][16][16];
#define TC_TRACE(role, slot)                                                                 \
  do {                                                                                       \
    if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && t < 16) g_tc_trace[role][t][slot] = clock64(); \
  } while (0)
// wall-clock nanoseconds next to the cycle counter: their ratio is the SM clock the kernel actually ran at
#define TC_TRACE_NS(role, slot)                                                              \
  do {                                                                                       \
    if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && t < 16) {                           \
      unsigned long long ns_;                                                                \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns_));                                \
      g_tc_trace[role][t][slot] = ns_;                                                       \
    }                                                                                        \
  } while (0)

struct TcDir {
  const uint16_t* wg_img;   // [C][(F+H)/16][NP][2][256][8]   gate weights, rows = 64*gate + unit
  const uint16_t* a1_img;   // [C][H/16][NP][2][128][8]       rows 0..63 attention_wh, 64..127 attention_w of this chunk
  const uint16_t* a2_img;   // [C][H/16][NP][2][64][8]        attention_wq rows of this chunk
  const float* P;           // [4][16][16]  zero padded mixing matrices  z = P_g u
  const float* bias4;       // [H][4]
  const float* bs;          // [H]
  const float* u;           // [H]
  const float* bu;          // [16]
  const float* h0; const float* c0; float* hT; float* cT;
  int reverse;
};
struct TcLayerParams {
  TcDir d[2];
  const uint16_t* x_img;    // [tiles][T][F/16][NP][2][128][8]
  float* y; int64_t syb, syt, yld;   // fp32 output (may be null)
  uint16_t* y_img; int y_kf;         // next layer's operand image (may be null) and its K/16
  int B, T, F, H, out_act, C, S, trace;   // S = slots of the weight ring
  int SX;                            // slots of the x-image ring
  int chunk;                         // bytes per bulk copy of the weight / x stream
  int xsplit;                        // x part as two N=128 MMAs per K block
  int xdefer;                        // first x segment only after the gate phase of the step
  int xprefetch;                     // pull the next step's x operand image into L2 one step ahead
  int earlypub;                      // 4-CTA clusters: first half of a CTA's block is sent while the second half is computed
                                     // (measured -2..4 % per step at H = 256, +1 % at H = 128, where it stays off)
  int puborder;                      // order in which a CTA serves its peers in the state exchange
  int rescale;                       // fp32 inference: peers' h' blocks are rescaled locally instead of a second all-gather
  int rescale_bf16;                  // the same in bf16 mode (costs a second bf16 rounding of the peers' blocks; experiment knob)
  int pubbytes;                      // TIMING DIAGNOSTIC ONLY (results are wrong if < kHBlock): bytes per state-exchange copy
  int xchg;                          // state exchange through L2: bulk store of the CTA's block, then ONE multicast bulk load into the peers
  uint8_t* xchg_buf;                 // [direction][tile][C][NP][16 KB] scratch of that exchange (caller's workspace)
  int acoll;                         // A-operand collector reuse between the hi*hi and hi*lo passes of a K block
  int nprod;                         // producer threads (one per warp) that issue the bulk copies of the weight / x stream, stage i by thread i % nprod
  int n1, n2;                        // x-part blocks of step t+1 issued before A1 / between A1 and A2 of step t (rest after A2)
  // training mode (TRAIN): tape of per-step intermediates (include/a3gc_b200.h, a3gc_tape) and optional recurrent-dropout mask
  a3gc_tape tape;
  const float* hmask;                // [D][B][T][15][H] or null
};

// barrier slots in shared memory
// BAR_H / BAR_HHAT / BAR_Q are ARRAYS of 4 barriers, one per source CTA of the cluster: the MMA warp consumes the K range
// of a chunk as soon as THAT chunk has arrived, its own chunk (available without any DSMEM transfer) first
// BAR_FULL / BAR_EMPTY: slots of the weight ring (every stage); BAR_XFULL / BAR_XEMPTY: slots of the x-image ring (x stages only)
enum { BAR_FULL = 0, BAR_EMPTY = kMaxStages, BAR_XFULL = 2 * kMaxStages, BAR_XEMPTY = 3 * kMaxStages,
       BAR_ACC_FULL = 4 * kMaxStages, BAR_ATT_FULL = BAR_ACC_FULL + 2, BAR_ATT2_FULL,
       BAR_ACC_EMPTY, BAR_A = BAR_ACC_EMPTY + 2, BAR_HFREE, BAR_A1FREE, BAR_H, BAR_HHAT = BAR_H + 4, BAR_Q = BAR_HHAT + 4,
       BAR_XRDY = BAR_Q + 4,     // [2] halves of this CTA's block are in its image: the exchange warp may store them (p.xchg)
       BAR_COUNT = BAR_XRDY + 2 };
constexpr int kBarSlots = 56;

// fp32-split mode: the warp-private transposition buffers are not a separate allocation -- each warp uses the 2 x 512 bytes of
// the state operand image it will itself overwrite at the end of the gate phase (see `wst` in the epilogue)
__host__ __device__ inline size_t tc_fixed_smem_bytes(int C, bool split) {
  return (split ? (size_t)0 : (size_t)kEpiWarps * kWstFloats * 4)   // wst
         + (size_t)C * 128 * 4                // apart
         + 4 * 128 * 4                        // ahalf
         + 256 * 4 + 64 * 4 * 2 + 16 * 4      // biasg, bs, u, bu
         + 1024 * 4                           // Pfrag
         + kBarSlots * 8 + 16;                // barriers, tmem slot
}

// byte offset of element (row, k) inside one part of an operand image [K/8][128][8]
__device__ __forceinline__ uint32_t img_off(int k, int row) {
  return (uint32_t)(((k >> 3) * kRows + row) * 16 + (k & 7) * 2);
}

template <bool SPLIT>
__device__ __forceinline__ void split_bits(float v, uint16_t& hi, uint16_t& lo) {
  if (SPLIT) {
    const __half h = __float2half_rn(v);
    hi = __half_as_ushort(h);
    lo = __half_as_ushort(__float2half_rn(v - __half2float(h)));
  } else {
    hi = __bfloat16_as_ushort(__float2bfloat16_rn(v));
    lo = 0;
  }
}

template <bool SPLIT, bool ATT, bool TRAIN>
__global__ void __launch_bounds__(kThreadsTC, 1)
tc_lstm_layer_kernel(const TcLayerParams p) {
  constexpr int NP = SPLIT ? 2 : 1;
  constexpr uint32_t kBBytes = NP * 2 * 256 * 16;     // one K=16 block of gate weights (all parts)
  constexpr uint32_t kABytes = NP * 2 * 128 * 16;     // one K=16 block of x rows
  constexpr uint32_t kA1Block = NP * 2 * 128 * 16;    // one K=16 block of [Wh ; Wa] rows
  constexpr uint32_t kA2Block = NP * 2 * 64 * 16;     // one K=16 block of Wq rows
  constexpr uint32_t kHBlock = 8 * kRows * 16;         // this CTA's 64 units of one operand part: 8 K-chunks
  extern __shared__ __align__(1024) uint8_t smem[];

  const int C = p.C, S = p.S, H = p.H, F = p.F, T = p.T;
  const uint32_t c = C > 1 ? ptx::cluster_ctarank() : 0u;
  const int tile = blockIdx.x / C;
  const TcDir& d = p.d[blockIdx.y];
  const int KF = F / 16, KH = H / 16;
  const int warp = threadIdx.x >> 5;
  const int n1 = p.n1, n2 = p.n2;

  // Two rings instead of one ring of (weights + x) slots: every stage takes a 16 KB (fp32-split) weight slot, only the x
  // stages take an 8 KB x-image slot.  The ring is latency-bound (a slot turns around in ~2 k cycles: copy ~1 k, its MMAs,
  // the commit -> producer hand-off), so its rate is (bytes in flight) / turn: at H = 256 the 128 KB state image left room
  // for three 24 KB slots; with the transposition buffers aliased into the state image (below) the same shared memory holds
  // four weight slots + three x slots, and the recurrent / attention stages no longer occupy x space they do not use.
  // bf16 mode keeps ONE ring of (weights + x) slots: its stages are a single 129-cycle MMA, and the second barrier wait and
  // second commit per x stage of the two-ring form cost more than the decoupling gains (AAGC bf16: -9 %, measured).
  constexpr bool kTwoRings = SPLIT;
  constexpr uint32_t kSlotStride = kTwoRings ? kBBytes : kBBytes + kABytes;
  const int SX = p.SX;
  uint8_t* hbuf = smem;
  uint8_t* ring = hbuf + (size_t)NP * H * 256;                // weight slots (bf16: weight + x slots)
  uint8_t* xring = ring + (size_t)S * kBBytes;                // x-image slots (fp32 mode)
  float* staging = reinterpret_cast<float*>(kTwoRings ? xring + (size_t)SX * kABytes : ring + (size_t)S * kSlotStride);   // [16 warps][8][32]  (bf16 mode only)
  float* apart = staging + (SPLIT ? 0 : kEpiWarps * kWstFloats);   // [C][128]
  float* ahalf = apart + C * 128;            // [4][128]
  float* biasg = ahalf + 4 * 128;            // [4 gates][64 units]
  float* bss = biasg + 256;                  // [64]
  float* us = bss + 64;                      // [64]
  float* bus = us + 64;                      // [16]
  uint32_t* Pfrag = reinterpret_cast<uint32_t*>(bus + 16);   // [4 gates][hi, lo][32 lanes][4 regs] A fragments of P_g
  uint64_t* bars = reinterpret_cast<uint64_t*>(Pfrag + 1024);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kBarSlots);

  // ------------------------------------------------------------------ setup
  if (threadIdx.x == 0) {
    for (int i = 0; i < BAR_COUNT; ++i)
      ptx::mbar_init(&bars[i], (i == BAR_HFREE || i == BAR_A1FREE) ? (uint32_t)C
                                   : (!kTwoRings && i >= BAR_FULL && i < BAR_FULL + kMaxStages) ? 2u : 1u);   // one ring: B side + A side
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, 512);
  {
    uint4* z = reinterpret_cast<uint4*>(hbuf);
    const int n16 = NP * H * 16;
    for (int i = threadIdx.x; i < n16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < 256; i += blockDim.x) biasg[i] = d.bias4[(size_t)c * 256 + (i & 63) * 4 + (i >> 6)];
    // A fragments (mma.m16n8k16, row-major A = P_g): reg r of lane l holds P_g[l/4 + 8*(r&1)][2*(l%4) + 8*(r>>1) + {0,1}]
    for (int i = threadIdx.x; i < 512; i += blockDim.x) {
      const int r = i & 3, l = (i >> 2) & 31, g = i >> 7;
      const int mm = (l >> 2) + 8 * (r & 1), nn = 2 * (l & 3) + 8 * (r >> 1);
      uint32_t hi, lo;
      ptx::split_pair_f16(d.P[(g * 16 + mm) * 16 + nn], d.P[(g * 16 + mm) * 16 + nn + 1], hi, lo);
      Pfrag[((g * 2 + 0) * 32 + l) * 4 + r] = hi;
      Pfrag[((g * 2 + 1) * 32 + l) * 4 + r] = lo;
    }
    if (ATT) {
      for (int i = threadIdx.x; i < 64; i += blockDim.x) { bss[i] = d.bs[c * 64 + i]; us[i] = d.u[c * 64 + i]; }
      if (threadIdx.x < 16) bus[threadIdx.x] = d.bu[threadIdx.x];
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (C > 1) ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint16_t cta_mask = (uint16_t)((1u << C) - 1u);

  const bool xchg_warp_on = p.xchg && !TRAIN && C > 2 && p.earlypub;   // (the host then runs two producers: p.nprod <= 2)
  if (warp == kXchgWarp && xchg_warp_on) {
    // ================================================================ exchange warp (p.xchg, 4-CTA clusters, inference): the
    // CTA's 64-unit block of hy goes to the peers through L2 -- a bulk store of each half as soon as the epilogue warps have
    // written it (BAR_XRDY), then ONE multicast bulk load per part drops it into every peer and completes on the peers'
    // barriers of source c.  The DSMEM copies it replaces move ~13 B/cycle per SM (the three peer blocks land 4.4 / 5.8 / 7.2 k
    // cycles after the hand-off) and read the block three times; this path needs ~2.5 k cycles per half and its wait for the
    // store blocks only this warp.
    {
      constexpr uint32_t kHalf = kHBlock / 2;
      uint8_t* mine = p.xchg_buf + ((((size_t)blockIdx.y * (gridDim.x / C) + tile) * C + c) * NP) * kHBlock;
      const int bar = ATT ? BAR_HHAT : BAR_H;
      const uint16_t peers = (uint16_t)(cta_mask & ~(1u << c));
      for (int t = 0; t < T; ++t) {
        for (int half = 0; half < 2; ++half) {
          ptx::mbar_wait(&bars[BAR_XRDY + half], t & 1);
          if (ptx::elect_one()) {
            for (int part = 0; part < NP; ++part)
              ptx::bulk_s2g(mine + (size_t)part * kHBlock + (size_t)half * kHalf,
                            hbuf + (size_t)part * H * 256 + (size_t)c * kHBlock + (size_t)half * kHalf, kHalf);
            ptx::bulk_commit_group();
            ptx::bulk_wait_group0();
            for (int part = 0; part < NP; ++part)
              ptx::bulk_g2s_multicast(hbuf + (size_t)part * H * 256 + (size_t)c * kHBlock + (size_t)half * kHalf,
                                      mine + (size_t)part * kHBlock + (size_t)half * kHalf, kHalf, &bars[bar + c], peers);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 0 || warp >= 2 + kEpiWarps) {
    // ================================================================ producers: weights / x -> ring
    // One thread issues a bulk copy every ~440 cycles whatever its size (tests/diag_stream_rate.py: 444 cycles per copy for
    // 4 KB .. 48 KB; 222 / 131 cycles with 2 / 4 issuing warps), and an H = 256, F = 512 step needs 92 copies (two per x
    // stage): 41 k cycles of ONE thread's time for a 44 k-cycle step.  So up to three warps share the stream, all walking the
    // same static stage sequence:  nprod = 2: stage i is issued by producer i % 2;  nprod = 3: the weight (B) copy of stage i
    // by producer i % 2 and the x-image (A) copy of every x stage by the third, which follows its own ring and therefore runs
    // ahead of the weight stream across the recurrent / attention stages.
    // Like the MMA warp, a producer warp walks the sequence on all 32 lanes (uniform operands) and one elected lane issues.
    const uint32_t my = __shfl_sync(0xffffffffu, warp == 0 ? 0u : (uint32_t)(warp - (1 + kEpiWarps)), 0);
    const uint32_t nprod = (uint32_t)p.nprod;
    const bool split_a = nprod == 3;                 // producer 2 = the A side
    const uint32_t nb = split_a ? 2u : nprod;        // producers that take turns on the stages
    // The order of the stages is the static issue order of the MMA warp: h-part of step t, then the x-part of step
    // t+1 cut into three segments placed around the two attention GEMMs of step t, so that the tensor pipe has work
    // while the epilogue warps are busy and the attention GEMMs are never queued behind a long x-part.
    if (my < nprod) {
      uint32_t ws = 0, wph = 0;           // weight-ring slot and the parity of its current fill (no div / mod in the loop)
      uint32_t xs = 0, xph = 0;           // the same for the x-image ring
      uint32_t turn = 0;                  // whose stage this is
      const uint32_t chunk = (uint32_t)p.chunk;
      auto load_stage = [&](const void* bsrc, uint32_t bbytes, const void* asrc, uint32_t abytes) {
        if (kTwoRings) {
          const bool b_side = my < nb && turn == my, a_side = abytes != 0 && (split_a ? my == 2u : b_side);
          if (b_side) {
            uint8_t* dst = ring + ws * kBBytes;
            ptx::mbar_wait(&bars[BAR_EMPTY + ws], wph ^ 1u);
            if (ptx::elect_one()) {
              ptx::mbar_arrive_expect_tx(&bars[BAR_FULL + ws], bbytes);
              // p.chunk (tuning knob) can cut the copies into pieces; measured slower than one copy per operand
              for (uint32_t o = 0; o < bbytes; o += chunk)
                ptx::bulk_g2s(dst + o, static_cast<const uint8_t*>(bsrc) + o, min(chunk, bbytes - o), &bars[BAR_FULL + ws]);
            }
            __syncwarp();
          }
          if (a_side) {
            uint8_t* dst = xring + xs * kABytes;
            ptx::mbar_wait(&bars[BAR_XEMPTY + xs], xph ^ 1u);
            if (ptx::elect_one()) {
              ptx::mbar_arrive_expect_tx(&bars[BAR_XFULL + xs], abytes);
              for (uint32_t o = 0; o < abytes; o += chunk)
                ptx::bulk_g2s(dst + o, static_cast<const uint8_t*>(asrc) + o, min(chunk, abytes - o), &bars[BAR_XFULL + xs]);
            }
            __syncwarp();
          }
        } else {
          // one ring: every FULL barrier takes two arrivals (B side, A side)
          uint8_t* dst = ring + ws * kSlotStride;
          const bool b_side = my < nb && turn == my, a_side = split_a ? my == 2u : b_side;
          if (b_side || a_side) ptx::mbar_wait(&bars[BAR_EMPTY + ws], wph ^ 1u);
          if ((b_side || a_side) && ptx::elect_one()) {
            if (b_side) {
              ptx::mbar_arrive_expect_tx(&bars[BAR_FULL + ws], bbytes);
              for (uint32_t o = 0; o < bbytes; o += chunk)
                ptx::bulk_g2s(dst + o, static_cast<const uint8_t*>(bsrc) + o, min(chunk, bbytes - o), &bars[BAR_FULL + ws]);
            }
            if (a_side) {
              if (abytes) {
                ptx::mbar_arrive_expect_tx(&bars[BAR_FULL + ws], abytes);
                for (uint32_t o = 0; o < abytes; o += chunk)
                  ptx::bulk_g2s(dst + kBBytes + o, static_cast<const uint8_t*>(asrc) + o, min(chunk, abytes - o), &bars[BAR_FULL + ws]);
              } else {
                ptx::mbar_arrive(&bars[BAR_FULL + ws]);
              }
            }
          }
          __syncwarp();
        }
        if (++turn == nb) turn = 0;
        if (++ws == (uint32_t)S) { ws = 0; wph ^= 1u; }
        if (abytes != 0 && ++xs == (uint32_t)SX) { xs = 0; xph ^= 1u; }
      };
      const uint8_t* wg = reinterpret_cast<const uint8_t*>(d.wg_img) + (size_t)c * (KF + KH) * kBBytes;
      const uint8_t* a1 = reinterpret_cast<const uint8_t*>(d.a1_img) + (size_t)c * KH * kA1Block;
      const uint8_t* a2 = reinterpret_cast<const uint8_t*>(d.a2_img) + (size_t)c * KH * kA2Block;
      const uint8_t* xi = reinterpret_cast<const uint8_t*>(p.x_img);
      auto xblocks = [&](int t, int kb0, int kb1) {
        const int ta = d.reverse ? T - 1 - t : t;
        const uint8_t* xs = xi + ((size_t)tile * T + ta) * KF * kABytes;
        // the x operand image is read once, from HBM; with three ring slots at H = 256 the stage latency bounds the
        // x-part stream, so this CTA's share of the NEXT step's image is pulled into L2 one step ahead
        if (kb0 == 0 && p.xprefetch && my == 0 && t + 1 < T) {
          const int tn = d.reverse ? T - 2 - t : t + 1;
          const uint32_t share = (uint32_t)(KF / C) * kABytes;
          if (ptx::elect_one()) ptx::bulk_prefetch_l2(xi + ((size_t)tile * T + tn) * KF * kABytes + (size_t)c * share, share);
          __syncwarp();
        }
        for (int kb = kb0; kb < kb1; ++kb) load_stage(wg + (size_t)kb * kBBytes, kBBytes, xs + (size_t)kb * kABytes, kABytes);
      };
      xblocks(0, 0, KF);
      for (int t = 0; t < T; ++t) {
        const bool nx = t + 1 < T;
        // the K range of the state operand is walked chunk by chunk starting with this CTA's own chunk (rotated order)
        for (int i = 0; i < C; ++i) {
          const int src = ((int)c + i) % C;
          for (int kb = 4 * src; kb < 4 * src + 4; ++kb) load_stage(wg + (size_t)(KF + kb) * kBBytes, kBBytes, nullptr, 0);
        }
        if (!ATT) {
          if (nx) xblocks(t + 1, 0, KF);
          continue;
        }
        if (nx) xblocks(t + 1, 0, n1);
        for (int i = 0; i < C; ++i) {
          const int src = ((int)c + i) % C;
          for (int s2 = 2 * src; s2 < 2 * src + 2; ++s2) load_stage(a1 + (size_t)s2 * 2 * kA1Block, 2 * kA1Block, nullptr, 0);
        }
        if (nx) xblocks(t + 1, n1, n1 + n2);
        for (int i = 0; i < C; ++i) {
          const int src = ((int)c + i) % C;
          load_stage(a2 + (size_t)src * 4 * kA2Block, 4 * kA2Block, nullptr, 0);
        }
        if (nx) xblocks(t + 1, n1 + n2, KF);
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    // The whole warp walks the static stage sequence -- slot indices, phases and descriptors are the same in every lane, so
    // they live in uniform registers -- and ONE elected lane issues the MMAs and commits of a stage, back to back
    // (ptx::elect_one).  Under `if (lane == 0)` every tcgen05 instruction was wrapped in an ELECT / R2UR / BRA.U.ANY waterfall and
    // a stage cost ~400-600 cycles of issue latency on that one thread -- more than its three MMAs take to execute, and the
    // real reason why ring depth, producer count and the restructurings of the chain changed so little.
    {
      const bool lane0 = (threadIdx.x & 31) == 0;
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
      const uint32_t idesc256 = ptx::make_idesc_f16(128, 256, !SPLIT);
      const uint32_t idesc128 = ptx::make_idesc_f16(128, 128, !SPLIT);
      const uint32_t idesc64 = ptx::make_idesc_f16(128, 64, !SPLIT);
      const uint32_t hbase = ptx::smem_u32(hbuf);
      const uint32_t hpart = (uint32_t)H * 256;
      uint32_t ws = 0, wph = 0, xs = 0, xph = 0;     // ring slots and the parities of their current fills (no div / mod)
      uint32_t empty_k[2] = {0, 0};     // completions of BAR_ACC_EMPTY[b] consumed so far (no-attention variant)
      const uint32_t ring_addr = ptx::smem_u32(ring), xring_addr = ptx::smem_u32(xring);
      // trace: cycles this warp waited for a full weight slot / x slot / anything else (state blocks, q, accumulators)
      const bool tr_on = p.trace && blockIdx.x == 0 && blockIdx.y == 0;
      unsigned long long w_full = 0, w_xfull = 0, w_other = 0, t_begin = tr_on ? clock64() : 0;
      auto wait_stage = [&]() -> uint32_t {          // next weight slot
        const unsigned long long t0 = tr_on ? clock64() : 0;
        ptx::mbar_wait(&bars[BAR_FULL + ws], wph);
        if (tr_on) w_full += clock64() - t0;
        ptx::tc_fence_after();
        return ring_addr + ws * kSlotStride;
      };
      auto wait_xstage = [&]() -> uint32_t {         // next x-image slot
        const unsigned long long t0 = tr_on ? clock64() : 0;
        ptx::mbar_wait(&bars[BAR_XFULL + xs], xph);
        if (tr_on) w_xfull += clock64() - t0;
        ptx::tc_fence_after();
        return xring_addr + xs * kABytes;
      };
      auto wait_other = [&](uint64_t* bar, uint32_t parity) {
        const unsigned long long t0 = tr_on ? clock64() : 0;
        ptx::mbar_wait(bar, parity);
        if (tr_on) w_other += clock64() - t0;
      };
      auto next_stage = [&]() { if (++ws == (uint32_t)S) { ws = 0; wph ^= 1u; } };
      auto next_xstage = [&]() { if (++xs == (uint32_t)SX) { xs = 0; xph ^= 1u; } };
      const uint64_t dA = ptx::make_smem_desc(0, kRows * 16, 128);
      const uint64_t dB256 = ptx::make_smem_desc(0, 256 * 16, 128), dB128 = ptx::make_smem_desc(0, 128 * 16, 128),
                     dB64 = ptx::make_smem_desc(0, 64 * 16, 128);
      const bool acoll = p.acoll != 0;
      // one K=16 block: D[:, dcol..dcol+N) (+)= A * B^T with the split passes hi*hi + lo*hi + hi*lo   (elected lane only)
      auto block_mma = [&](uint32_t dcol, uint32_t a0, uint32_t astride, uint32_t b0, uint32_t bstride, uint64_t dB,
                           uint32_t idesc, bool first) {
        // (shared-window addresses of non-zero cluster ranks carry the rank above bit 18: keep the 18-bit offset only)
        const uint64_t ah = dA + ((a0 & 0x3FFFFu) >> 4), bh = dB + ((b0 & 0x3FFFFu) >> 4);
        if (SPLIT && acoll) {
          // A_hi is multiplied by B_hi and by B_lo back to back: the second MMA takes it from the collector buffer
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
      // p.xsplit: issue the (off-critical-path) x part as two N = 128 halves instead of one N = 256 MMA.  The warp-level
      // HMMAs of the gate mix share the tensor pipe with the UMMAs and only get a slot between two of them; shorter
      // UMMAs halve that wait at the cost of reading the A operand twice.
      const bool xsplit = p.xsplit != 0;
      auto xblocks = [&](uint32_t dcol, int kb0, int kb1) {
        for (int kb = kb0; kb < kb1; ++kb) {
          const uint32_t sa = wait_stage();
          const uint32_t xa = kTwoRings ? wait_xstage() : sa + kBBytes;
          if (ptx::elect_one()) {
            if (xsplit) {
              block_mma(dcol, xa, kABytes / NP, sa, kBBytes / NP, dB256, idesc128, kb == 0);
              block_mma(dcol + 128, xa, kABytes / NP, sa + 128 * 16, kBBytes / NP, dB256, idesc128, kb == 0);
            } else {
              block_mma(dcol, xa, kABytes / NP, sa, kBBytes / NP, dB256, idesc256, kb == 0);
            }
            ptx::umma_commit(&bars[BAR_EMPTY + ws]);
            if (kTwoRings) ptx::umma_commit(&bars[BAR_XEMPTY + xs]);
          }
          __syncwarp();
          next_stage();
          if (kTwoRings) next_xstage();
        }
      };
      xblocks(0, 0, KF);
      for (int t = 0; t < T; ++t) {
        const uint32_t b = t & 1, bo = b ^ 1u, dcol = b * 256;
        const bool nx = t + 1 < T;
        // h-part of step t (needs h'_{t-1} of every chunk in local shared memory)
        if (lane0) TC_TRACE(1, 0);
        for (int i = 0; i < C; ++i) {
          const int src = ((int)c + i) % C;
          wait_other(&bars[BAR_H + src], t & 1);
          ptx::tc_fence_after();
          if (i == 0 && lane0) TC_TRACE(1, 1);
          for (int kb = 4 * src; kb < 4 * src + 4; ++kb) {
            const uint32_t sa = wait_stage();
            if (ptx::elect_one()) {
              block_mma(dcol, hbase + (uint32_t)kb * 2 * kRows * 16, hpart, sa, kBBytes / NP, dB256, idesc256, false);
              ptx::umma_commit(&bars[BAR_EMPTY + ws]);
            }
            __syncwarp();
            next_stage();
          }
        }
        if (ptx::elect_one()) {
          ptx::umma_commit(&bars[BAR_ACC_FULL + b]);
          if (C > 1) ptx::umma_commit_multicast(&bars[BAR_HFREE], cta_mask); else ptx::umma_commit(&bars[BAR_HFREE]);
        }
        __syncwarp();
        if (lane0) TC_TRACE(1, 2);
        // x-part of step t+1 goes into the other buffer (free once the epilogue of step t-1 has drained it)
        if (nx && t >= 1) {
          if (ATT) wait_other(&bars[BAR_ACC_EMPTY + bo], 1u);
          else { wait_other(&bars[BAR_ACC_EMPTY + bo], empty_k[bo] & 1u); empty_k[bo] += 1; }
          ptx::tc_fence_after();
        }
        if (!ATT) {
          if (nx) xblocks(bo * 256, 0, KF);
          if (lane0) TC_TRACE(1, 3);
          continue;
        }
        // p.xdefer (tuning knob, default off): hold the first x segment back until the gate phase of this step is over.  The
        // warp-level HMMAs of the mix share the tensor pipe with the UMMAs; deferring shortens the gate phase from ~12 k to
        // ~7 k cycles at H=256 but the x part then delays the attention GEMM by as much: no net gain (measured).
        if (p.xdefer) { ptx::mbar_wait(&bars[BAR_ACC_EMPTY + b], 0u); ptx::tc_fence_after(); }
        if (nx) xblocks(bo * 256, 0, n1);
        if (lane0) TC_TRACE(1, 3);
        // A1: [Wh hy | Wa (hy, node-sum in row 15)] -> columns [0,128) of the drained buffer
        wait_other(&bars[BAR_ACC_EMPTY + b], 0u);
        for (int i = 0; i < C; ++i) {
          const int src = ((int)c + i) % C;
          wait_other(&bars[BAR_HHAT + src], t & 1);
          ptx::tc_fence_after();
          if (lane0) {
            if (i == 0) TC_TRACE(1, 4);
            if (i == 1) TC_TRACE(1, 9);
            if (i == 2) TC_TRACE(1, 10);
            if (i == 3) TC_TRACE(1, 11);
          }
          for (int s2 = 2 * src; s2 < 2 * src + 2; ++s2) {
            const uint32_t sa = wait_stage();
            if (ptx::elect_one()) {
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                const int kb = s2 * 2 + j;
                block_mma(dcol, hbase + (uint32_t)kb * 2 * kRows * 16, hpart, sa + (uint32_t)j * kA1Block, kA1Block / NP, dB128, idesc128,
                          i == 0 && s2 == 2 * src && j == 0);
              }
              ptx::umma_commit(&bars[BAR_EMPTY + ws]);
            }
            __syncwarp();
            next_stage();
          }
        }
        if (ptx::elect_one()) {
          ptx::umma_commit(&bars[BAR_ATT_FULL]);
          if (C > 1) ptx::umma_commit_multicast(&bars[BAR_A1FREE], cta_mask); else ptx::umma_commit(&bars[BAR_A1FREE]);
        }
        __syncwarp();
        if (lane0) TC_TRACE(1, 5);
        if (nx) xblocks(bo * 256, n1, n1 + n2);
        // A2: Wq q (q sits in row 15 of every sequence) -> columns [128,192)
        for (int i = 0; i < C; ++i) {
          const int src = ((int)c + i) % C;           // q of chunk src sits in row 15 of K blocks 4 src .. 4 src + 3
          { const unsigned long long t0 = tr_on ? clock64() : 0; ptx::mbar_wait_cluster(&bars[BAR_Q + src], t & 1); if (tr_on) w_other += clock64() - t0; }
          ptx::fence_proxy_async();
          ptx::tc_fence_after();
          if (i == 0 && lane0) TC_TRACE(1, 6);
          const uint32_t sa = wait_stage();
          if (ptx::elect_one()) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int kb = src * 4 + j;
              block_mma(dcol + 128, hbase + (uint32_t)kb * 2 * kRows * 16, hpart, sa + (uint32_t)j * kA2Block, kA2Block / NP, dB64, idesc64,
                        i == 0 && j == 0);
            }
            ptx::umma_commit(&bars[BAR_EMPTY + ws]);
          }
          __syncwarp();
          next_stage();
        }
        if (ptx::elect_one()) ptx::umma_commit(&bars[BAR_ATT2_FULL]);
        __syncwarp();
        if (lane0) TC_TRACE(1, 7);
        if (nx) xblocks(bo * 256, n1 + n2, KF);
        if (lane0) TC_TRACE(1, 8);
      }
      if (tr_on && lane0) {
        g_tc_trace[1][15][0] = w_full; g_tc_trace[1][15][1] = w_xfull; g_tc_trace[1][15][2] = w_other; g_tc_trace[1][15][3] = clock64() - t_begin;
      }
    }
  } else {
    // ================================================================ epilogue warps (512 threads)
    // Warp (qd, ug) owns accumulator rows 32qd..32qd+31 (sequences 2qd, 2qd+1 of the tile; a warp may only read the
    // TMEM lane quarter warp%4) x hidden units 16ug..16ug+15 of this CTA's 64.  Gate accumulators go TMEM -> registers
    // (lane = row) -> a WARP-PRIVATE transposition buffer -> B fragments of mma.m16n8k16 (k = node), so the whole
    // gate phase needs no CTA-wide barrier.  Pointwise ownership follows the C fragment: lane (tq = lane/4,
    // tr = lane%4) owns, for sequence sq and 8-unit block ub, nodes {tq, tq+8} x units {2tr, 2tr+1};
    // node 15 (tq == 7, upper half) is the pad slot.
    const int et = threadIdx.x - 64;
    const int ew = et >> 5, lane = et & 31;
    const int qd = warp & 3;
    const int ug = ew >> 2;
    const int tq = lane >> 2, tr = lane & 3;
    const bool pad_hi = tq == 7;
    const uint32_t tmem_row = tmem + ((uint32_t)(qd * 32) << 16);
    const int ycol = blockIdx.y * H + (int)c * 64;
    const int ubase = 16 * ug;                       // attention outputs: first unit (within the CTA's 64) of this warp
    // gate / state units of this warp: gbase + 32 ub + [0, 8).  Unit block ub = 0 of all warps is the first half of the
    // CTA's 64 units (K chunks 0..3 of its block of the operand image), so that half can be sent to the peers while the
    // second unit block is still being computed.
    // (The training variant keeps the contiguous mapping gbase = 16 ug, unit block stride 8: it sends whole blocks, and
    // its tape stores are measurably faster with 16 contiguous units per warp: 34 vs 40 ms per stage-1 step.)
    constexpr int kUbs = TRAIN ? 8 : 32;             // unit distance between the two unit blocks of a warp
    const int gbase = TRAIN ? 16 * ug : 8 * ug;
    // Warp-private transposition buffer: 8 accumulator columns x 32 rows fp32 = 1 KB, as two 512-byte halves (columns 0..3 at
    // wst, columns 4..7 at wst2).  bf16 mode: a separate allocation.  fp32-split mode: the rows 32qd..32qd+31 of the K chunk of
    // this warp's SECOND unit block in the CTA's own block of the state image, hi part and lo part -- 2 x 512 bytes that only
    // this warp writes, and only at the end of its gate phase (store_units of unit block 1, after the last use of the buffer,
    // rewrites every byte).  The region is dead during the gate phase: the phase starts after BAR_HFREE (every MMA of the
    // cluster that reads h'_{t-1} has completed, and therefore every outgoing copy of the block has landed), peers never
    // write into a CTA's own block, and this CTA's q rows are written after the gate phase.
    // the first epilogue warp issues the hand-offs (bulk copies to the peers, remote barrier arrives): warp-uniform flag, so that
    // the operands of those instructions are uniform and one elected lane issues them without a per-instruction waterfall
    const bool warp0 = __shfl_sync(0xffffffffu, ew, 0) == 0;
    float* wst = staging + ew * kWstFloats;
    float* wst2 = wst + 4 * 32;
    if (SPLIT) {
      const int ch1 = (gbase + kUbs) >> 3;
      wst = reinterpret_cast<float*>(hbuf + ((size_t)((int)c * 8 + ch1) * kRows + 32 * qd) * 16);
      wst2 = wst + (size_t)H * 64;                   // the lo part of the image is H * 256 bytes further on
    }
    const float* wsp = (tq < 4 ? wst : wst2) + (tq & 3) * 32;      // column tq of the buffer (B-fragment reads)
    int bseq[2]; bool valid[2];
#pragma unroll
    for (int sq = 0; sq < 2; ++sq) { bseq[sq] = tile * kSeqTile + 2 * qd + sq; valid[sq] = bseq[sq] < p.B; }

    float creg[2][2][4], hreg[2][2][4];              // [sequence sq][unit block ub][element j]
#pragma unroll
    for (int sq = 0; sq < 2; ++sq)
#pragma unroll
      for (int ub = 0; ub < 2; ++ub)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int node = tq + 8 * (j >> 1), unit = gbase + kUbs * ub + 2 * tr + (j & 1);
          const bool ok = valid[sq] && node < kNodes;
          const size_t gi = ((size_t)bseq[sq] * kNodes + node) * H + c * 64 + unit;
          creg[sq][ub][j] = (ok && d.c0 != nullptr) ? d.c0[gi] : 0.f;
          hreg[sq][ub][j] = (ok && d.h0 != nullptr) ? d.h0[gi] : 0.f;
        }

    // write this thread's 16 values into the local operand image (rows of sequences 2qd, 2qd+1)
    // tm >= 0 (training with recurrent dropout): the image receives v * hmask[., tm, ., .], the mask of the step that will
    // consume it (net_aagc.py:181 drops the h that enters the gates only; the carried state stays unmasked)
    auto store_units = [&](const float (&v)[2][2][4], int tm = -1, int only_ub = -1) {
#pragma unroll
      for (int sq = 0; sq < 2; ++sq)
#pragma unroll
        for (int ub = 0; ub < 2; ++ub) {
          if (only_ub >= 0 && ub != only_ub) continue;
          const int k = (int)c * 64 + gbase + kUbs * ub + 2 * tr;
#pragma unroll
          for (int up = 0; up < 2; ++up) {
            uint32_t hi, lo;
            float v0 = v[sq][ub][2 * up], v1 = v[sq][ub][2 * up + 1];
            if (TRAIN && tm >= 0 && valid[sq] && tq + 8 * up < kNodes) {
              const float2 mk = *reinterpret_cast<const float2*>(
                  p.hmask + ((((size_t)blockIdx.y * p.B + bseq[sq]) * T + tm) * kNodes + tq + 8 * up) * H + k);
              v0 *= mk.x; v1 *= mk.y;
            }
            if (SPLIT) {
              ptx::split_pair_f16(v0, v1, hi, lo);
            } else {
              const __nv_bfloat162 bb = __floats2bfloat162_rn(v0, v1);
              hi = *reinterpret_cast<const uint32_t*>(&bb);
              lo = 0;
            }
            const uint32_t off = img_off(k, 16 * (2 * qd + sq) + tq + 8 * up);
            *reinterpret_cast<uint32_t*>(hbuf + off) = hi;
            if (SPLIT) *reinterpret_cast<uint32_t*>(hbuf + (size_t)H * 256 + off) = lo;
          }
        }
    };
    // all epilogue threads have written their part of the local operand image: make it visible to the async proxy,
    // then send this CTA's 64-unit block to every peer; `bar` completes in each CTA when all C blocks have landed.
    // `acc_empty` >= 0: the accumulator buffer has been drained as well (tcgen05 loads fenced before the barrier).
    // this CTA's slot of the L2 exchange scratch: [direction][tile][chunk][part][16 KB]
    uint8_t* xchg_mine = p.xchg_buf == nullptr ? nullptr
        : p.xchg_buf + ((((size_t)blockIdx.y * (gridDim.x / C) + tile) * C + c) * NP) * kHBlock;
    auto publish_block = [&](int bar, int acc_empty) {
      ptx::fence_proxy_async();
      if (acc_empty >= 0) ptx::tc_fence_before();
      ptx::named_bar_sync(1, kEpiThreads);
      if (warp0 && ptx::elect_one()) {
        if (acc_empty >= 0) ptx::mbar_arrive(&bars[BAR_ACC_EMPTY + acc_empty]);
        ptx::mbar_arrive(&bars[bar + c]);                               // own chunk: usable at once
        // Send order: CTA c serves c-1 first, then c-2, ...: a receiver walks the sources own, r+1, r+2, ... (rotated K order),
        // so the block it needs first is the one every sender pushes first (p.puborder 0 = fixed order 0, 1, 2, ...; 2 = c+1 first)
        if (p.xchg && C > 1) {
          // through L2: store the block once, then one multicast load per part drops it into every peer (and completes on the
          // peers' barriers of source c) -- the DSMEM copies move ~13 B/cycle per SM and read the block three times
          for (int part = 0; part < NP; ++part)
            ptx::bulk_s2g(xchg_mine + (size_t)part * kHBlock, hbuf + (size_t)part * H * 256 + (size_t)c * kHBlock, kHBlock);
          ptx::bulk_commit_group();
          ptx::bulk_wait_group0();
          for (int part = 0; part < NP; ++part)
            ptx::bulk_g2s_multicast(hbuf + (size_t)part * H * 256 + (size_t)c * kHBlock, xchg_mine + (size_t)part * kHBlock, kHBlock, &bars[bar + c],
                                    (uint16_t)(cta_mask & ~(1u << c)));
        } else
        for (uint32_t i = 1; i < (uint32_t)C; ++i) {
          const uint32_t peer = p.puborder == 1 ? (c + (uint32_t)C - i) % (uint32_t)C
                              : p.puborder == 2 ? (c + i) % (uint32_t)C
                              : (i - 1 < c ? i - 1 : i);
          for (int part = 0; part < NP; ++part)                           // lands on the peer's barrier of source c
            ptx::bulk_s2remote(hbuf + (size_t)part * H * 256 + (size_t)c * kHBlock, (uint32_t)p.pubbytes, &bars[bar + c], peer);
        }
        for (uint32_t peer = 0; peer < (uint32_t)C; ++peer)
          if (peer != c) ptx::mbar_arrive_expect_tx(&bars[bar + peer], (uint32_t)NP * (uint32_t)p.pubbytes);   // arm the barriers of the peers' blocks
      }
    };
    // The same hand-off in two halves: half 0 = K chunks 0..3 of this CTA's block (unit block ub = 0 of every warp), sent
    // while the second unit block is still being computed; half 1 completes the block, arms the barriers and releases the
    // accumulator buffer.  (A complete_tx that lands before the receiver has armed the phase is fine: the phase cannot
    // complete before its arming arrive.)
    auto publish_half = [&](int bar, int half, int acc_empty) {
      ptx::fence_proxy_async();
      if (acc_empty >= 0) ptx::tc_fence_before();
      ptx::named_bar_sync(1, kEpiThreads);
      if (warp0 && ptx::elect_one()) {
        constexpr uint32_t kHalf = kHBlock / 2;
        if (p.xchg) {
          ptx::mbar_arrive(&bars[BAR_XRDY + half]);                      // the exchange warp stores / multicasts this half
        } else
        for (uint32_t i = 1; i < (uint32_t)C; ++i) {
          const uint32_t peer = (c + (uint32_t)C - i) % (uint32_t)C;
          for (int part = 0; part < NP; ++part)
            ptx::bulk_s2remote(hbuf + (size_t)part * H * 256 + (size_t)c * kHBlock + (size_t)half * kHalf, kHalf, &bars[bar + c], peer);
        }
        if (half == 1) {
          if (acc_empty >= 0) ptx::mbar_arrive(&bars[BAR_ACC_EMPTY + acc_empty]);
          ptx::mbar_arrive(&bars[bar + c]);
          for (uint32_t peer = 0; peer < (uint32_t)C; ++peer)
            if (peer != c) ptx::mbar_arrive_expect_tx(&bars[bar + peer], (uint32_t)NP * kHBlock);
        }
      }
    };
    // y_t = act(h'_t).  All addressing that does not depend on t is folded into per-thread bases; the values of
    // pad slots (node 15, sequences beyond the batch) are exact zeros already (masked in the gate phase), so the
    // operand-image stores need no predicate: the next layer multiplies those rows by zero weights of P.
    const int col0 = ycol + gbase + 2 * tr;                        // feature of (ub = 0, unit 2tr); ub = 1 is 32 features on
    const int64_t ybase = (int64_t)bseq[0] * p.syb + (int64_t)tq * p.yld + col0;
    const int ysq = (int)p.syb, yup = 8 * (int)p.yld;               // element offsets of (sq = 1) and (node + 8)
    const size_t img_step = (size_t)p.y_kf * NP * 4096;             // bytes of one (tile, t) slab of the output image
    const size_t img_base = ((size_t)tile * T * p.y_kf + (col0 >> 4)) * NP * 4096 + (TRAIN ? (size_t)0 : (size_t)(ug & 1) * 2048) + (size_t)(32 * qd + tq) * 16 + 4 * tr;
    auto emit = [&](int ta, const float (&v)[2][2][4]) {
      const bool th = p.out_act == A3GC_ACT_TANH;
      float o[2][2][4];
#pragma unroll
      for (int sq = 0; sq < 2; ++sq)
#pragma unroll
        for (int ub = 0; ub < 2; ++ub)
#pragma unroll
          for (int j = 0; j < 4; ++j) o[sq][ub][j] = th ? fast_tanh(v[sq][ub][j]) : v[sq][ub][j];
      if (p.y != nullptr) {
        float* yp = p.y + ybase + (int64_t)ta * p.syt;
#pragma unroll
        for (int sq = 0; sq < 2; ++sq)
#pragma unroll
          for (int up = 0; up < 2; ++up) {
            if (!valid[sq] || (up == 1 && pad_hi)) continue;
#pragma unroll
            for (int ub = 0; ub < 2; ++ub)
              *reinterpret_cast<float2*>(yp + sq * ysq + up * yup + kUbs * ub) = make_float2(o[sq][ub][2 * up], o[sq][ub][2 * up + 1]);
          }
      }
      if (p.y_img != nullptr) {
        // element (row 16s+node, feature k) of image [tile][ta][k/16][part][(k/8)%2][128][8]
        uint8_t* ip = reinterpret_cast<uint8_t*>(p.y_img) + img_base + (size_t)ta * img_step;
#pragma unroll
        for (int sq = 0; sq < 2; ++sq)
#pragma unroll
          for (int ub = 0; ub < 2; ++ub)
#pragma unroll
            for (int up = 0; up < 2; ++up) {
              uint32_t hi, lo = 0;
              if (SPLIT) {
                ptx::split_pair_f16(o[sq][ub][2 * up], o[sq][ub][2 * up + 1], hi, lo);
              } else {
                const __nv_bfloat162 bb = __floats2bfloat162_rn(o[sq][ub][2 * up], o[sq][ub][2 * up + 1]);
                hi = *reinterpret_cast<const uint32_t*>(&bb);
              }
              uint8_t* q = ip + (TRAIN ? (size_t)ub * 2048 : (size_t)ub * 2 * NP * 4096) + sq * 256 + up * 128;   // ub = 1: 8 / 32 features on
              *reinterpret_cast<uint32_t*>(q) = hi;
              if (SPLIT) *reinterpret_cast<uint32_t*>(q + 4096) = lo;
            }
      }
    };

    store_units(hreg, (TRAIN && p.hmask != nullptr) ? (d.reverse ? T - 1 : 0) : -1);   // h_{-1} = h0  (completion #0 of BAR_H)
    publish_block(BAR_H, -1);

    // tape.hp [D][B][T][15][H] (node-major like x): h'_t of this thread's elements
    auto tape_hp = [&](int ta) {
      float* hp0 = p.tape.hp + ((((size_t)blockIdx.y * p.B + bseq[0]) * T + ta) * kNodes + tq) * H + (int)c * 64 + gbase + 2 * tr;
      const size_t seq_stride = (size_t)T * kNodes * H;
#pragma unroll
      for (int sq = 0; sq < 2; ++sq) {
        if (!valid[sq]) continue;
#pragma unroll
        for (int up = 0; up < 2; ++up) {
          if (tq + 8 * up >= kNodes) continue;
          float* hp = hp0 + sq * seq_stride + (size_t)up * 8 * H;
#pragma unroll
          for (int ub = 0; ub < 2; ++ub) *reinterpret_cast<float2*>(hp + kUbs * ub) = make_float2(hreg[sq][ub][2 * up], hreg[sq][ub][2 * up + 1]);
        }
      }
    };
    const uint4* Pfrag4 = reinterpret_cast<const uint4*>(Pfrag);
    const int swz = 8 * (tq & 3);                 // XOR swizzle of the transposition buffer column tq

    for (int t = 0; t < T; ++t) {
      const uint32_t b = t & 1;
      const int ta = d.reverse ? T - 1 - t : t;
      // time of the next step of this direction (the step that consumes h'_t), -1: none / no recurrent dropout
      const int tnext = (TRAIN && p.hmask != nullptr && t + 1 < T) ? (d.reverse ? ta - 1 : ta + 1) : -1;
      // ---------------------------------------------------------------- gates -> c', hy
      if (et == 0) TC_TRACE(0, 0);
      ptx::mbar_wait(&bars[BAR_ACC_FULL + b], (t >> 1) & 1);
      // every CTA of the cluster has finished reading h'_{t-1}: the operand images may be overwritten
      ptx::mbar_wait(&bars[BAR_HFREE], t & 1);
      ptx::tc_fence_after();
      if (et == 0) TC_TRACE(0, 1);
      if (et == 0) TC_TRACE_NS(0, 15);
      // training: per-step bases of the tape rows this thread writes in the gate phase, so that the 128 stores of the phase
      // take an immediate offset each instead of a 64-bit index computation (they were a third of the phase's instructions)
      const size_t H16 = (size_t)H * 16;
      float* tape_g0 = nullptr; float* tape_u0 = nullptr;
      if (TRAIN) {
        const size_t rec_t = (size_t)blockIdx.y * T + ta;
        tape_g0 = p.tape.gates + (rec_t * p.B + bseq[0]) * 4 * H16 + (size_t)((int)c * 64 + gbase + 2 * tr) * 16 + tq;
        if (p.tape.u != nullptr)
          tape_u0 = p.tape.u + (rec_t * p.B + (tile * kSeqTile + 2 * qd + (lane >> 4))) * 4 * H16 + (size_t)((int)c * 64 + gbase) * 16 + (lane & 15);
      }
#pragma unroll
      for (int ub = 0; ub < 2; ++ub) {
        float e1[2][4], e2[2][4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          {   // 8 accumulator columns (gate g, units gbase+32ub .. +7) TMEM -> wst[col][row ^ swizzle(col)]
            float v[8];
            ptx::tmem_ld8(tmem_row + b * 256 + g * 64 + gbase + kUbs * ub, v);
            if (TRAIN) {
              // tape.u [rec][gate][unit][16]: this lane = accumulator row = (sequence, node); node slot 15 holds an exact 0
              const int rs = tile * kSeqTile + 2 * qd + (lane >> 4);
              if (tape_u0 != nullptr && rs < p.B) {
                float* up = tape_u0 + (size_t)g * H16 + kUbs * ub * 16;
#pragma unroll
                for (int j = 0; j < 8; ++j) up[j * 16] = v[j];
              }
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j) (j < 4 ? wst : wst2)[(j & 3) * 32 + (lane ^ (8 * (j & 3)))] = v[j];
            __syncwarp();
          }
          // z[16 nodes x 8 units] = P_g (16x16) . U (16 nodes x 8 units) per sequence on the warp-level tensor path,
          // fp16 hi/lo split (3 passes) so the mix stays fp32-accurate
          const uint4 ah4 = Pfrag4[(g * 2 + 0) * 32 + lane], al4 = Pfrag4[(g * 2 + 1) * 32 + lane];
          const uint32_t ah[4] = {ah4.x, ah4.y, ah4.z, ah4.w}, al[4] = {al4.x, al4.y, al4.z, al4.w};
          const float2 bias = *reinterpret_cast<const float2*>(biasg + g * 64 + gbase + kUbs * ub + 2 * tr);
#pragma unroll
          for (int sq = 0; sq < 2; ++sq) {
            const float2 u0 = *reinterpret_cast<const float2*>(wsp + ((16 * sq + 2 * tr) ^ swz));        // nodes 2tr, 2tr+1 of unit column tq
            const float2 u1 = *reinterpret_cast<const float2*>(wsp + ((16 * sq + 2 * tr + 8) ^ swz));    // nodes 2tr+8, 2tr+9
            float z[4] = {bias.x, bias.y, bias.x, bias.y};
            if (SPLIT) {
              uint32_t bh0, bl0, bh1, bl1;
              ptx::split_pair_f16(u0.x, u0.y, bh0, bl0);
              ptx::split_pair_f16(u1.x, u1.y, bh1, bl1);
              // three INDEPENDENT accumulators: a warp-level HMMA queues behind the UMMAs that share the tensor pipe, so a
              // dependent chain of three would pay that latency three times
              float z2[4] = {0.f, 0.f, 0.f, 0.f}, z3[4] = {0.f, 0.f, 0.f, 0.f};
              ptx::mma_16816_f16(z, ah, bh0, bh1);
              ptx::mma_16816_f16(z2, al, bh0, bh1);
              ptx::mma_16816_f16(z3, ah, bl0, bl1);
#pragma unroll
              for (int j = 0; j < 4; ++j) z[j] += z2[j] + z3[j];
            } else {
              // bf16 path (stated bound 5e-3): one fp16 pass for the mix (11-bit operands, fp32 accumulate)
              const __half2 h0 = __floats2half2_rn(u0.x, u0.y), h1 = __floats2half2_rn(u1.x, u1.y);
              ptx::mma_16816_f16(z, ah, *reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
            }
            float* tape_gp = TRAIN ? tape_g0 + (size_t)(4 * sq + g) * H16 + kUbs * ub * 16 : nullptr;   // + (j & 1) * 16 + 8 * (j >> 1)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const bool ok = valid[sq] && !(pad_hi && j >= 2);
              // c' = sig(zf) c + sig(zi) tanh(zc), hy = sig(zo) tanh(c') with shared reciprocals: 5 ex2 + 2 rcp per element
              float gv = 0.f;                                                                    // activated gate (tape only)
              if (!SPLIT) {
                // bf16 path: single-MUFU activations (5 per element)
                if (g == 0) e1[sq][j] = sigmoid_approx(z[j]);
                else if (g == 1) e2[sq][j] = sigmoid_approx(z[j]);
                else if (g == 2) {
                  float cn = fmaf(e2[sq][j], creg[sq][ub][j], e1[sq][j] * tanh_approx(z[j]));
                  cn = ok ? cn : 0.f;
                  creg[sq][ub][j] = cn;
                  e1[sq][j] = tanh_approx(cn);
                } else {
                  hreg[sq][ub][j] = ok ? sigmoid_approx(z[j]) * e1[sq][j] : 0.f;
                }
              } else if (g == 0) {
                e1[sq][j] = one_plus_exp_neg(z[j]);                                              // 1 + e^-zi
                if (TRAIN) gv = rcp_ftz(e1[sq][j]);
              } else if (g == 1) {
                e2[sq][j] = one_plus_exp_neg(z[j]);                                              // 1 + e^-zf
                if (TRAIN) gv = rcp_ftz(e2[sq][j]);
              } else if (g == 2) {
                const float eb = exp_neg2(z[j]);                                                 // tanh(zc) = (1 - eb) / (1 + eb)
                if (TRAIN) gv = (1.0f - eb) * rcp_ftz(1.0f + eb);
                const float dab = e1[sq][j] * (1.0f + eb);
                float cn = fmaf(creg[sq][ub][j], dab, (1.0f - eb) * e2[sq][j]) * rcp_ftz(e2[sq][j] * dab);
                cn = ok ? cn : 0.f;
                creg[sq][ub][j] = cn;
                const float ec = exp_neg2(cn);
                e1[sq][j] = 1.0f - ec;                                                           // tanh(c') = (1 - ec) / (1 + ec)
                e2[sq][j] = 1.0f + ec;
              } else {
                const float eo = one_plus_exp_neg(z[j]);
                const float hy = e1[sq][j] * rcp_ftz(e2[sq][j] * eo);
                if (TRAIN) gv = rcp_ftz(eo);
                hreg[sq][ub][j] = ok ? hy : 0.f;
              }
              if (TRAIN && valid[sq]) tape_gp[(j & 1) * 16 + 8 * (j >> 1)] = ok ? gv : 0.f;   // [rec][gate][unit][node]
            }
          }
        }
        // ---- unit block ub is final: write it into the local operand image (+ the node sums for the attention GEMM) and,
        // for ub = 0, start sending that half of the block to the peers while the second unit block is computed.  (The
        // training variant, already short of registers because of the tape, does this once after both unit blocks.)
        if (TRAIN) {
        } else if (ATT) {
          // node sum of hy (q_t = relu((sum_n hy_n) W_a^T), net_aagc.py:200) goes to row 15 of the sequence, which the
          // attention GEMM reads as a 16th "node": the lanes owning the pad slot store it there
          float keep[2][2];
#pragma unroll
          for (int sq = 0; sq < 2; ++sq)
#pragma unroll
            for (int u2 = 0; u2 < 2; ++u2) {
              float sum = hreg[sq][ub][u2] + hreg[sq][ub][2 + u2];
              sum += __shfl_xor_sync(0xffffffffu, sum, 4);
              sum += __shfl_xor_sync(0xffffffffu, sum, 8);
              sum += __shfl_xor_sync(0xffffffffu, sum, 16);
              keep[sq][u2] = hreg[sq][ub][2 + u2];
              if (pad_hi) hreg[sq][ub][2 + u2] = sum;
              if (TRAIN && pad_hi && valid[sq])
                p.tape.s[(((size_t)blockIdx.y * T + ta) * p.B + bseq[sq]) * H + (int)c * 64 + gbase + kUbs * ub + 2 * tr + u2] = sum;
            }
          store_units(hreg, -1, ub);
#pragma unroll
          for (int sq = 0; sq < 2; ++sq)
#pragma unroll
            for (int u2 = 0; u2 < 2; ++u2) hreg[sq][ub][2 + u2] = keep[sq][u2];
        } else {
          store_units(hreg, tnext, ub);
        }
        if (!TRAIN && ub == 0 && C > 2 && p.earlypub) publish_half(ATT ? BAR_HHAT : BAR_H, 0, -1);
      }
      if (TRAIN) {
#pragma unroll
        for (int sq = 0; sq < 2; ++sq) {
          if (!valid[sq]) continue;
          const size_t rec = ((size_t)blockIdx.y * T + ta) * p.B + bseq[sq];
#pragma unroll
          for (int ub = 0; ub < 2; ++ub)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const size_t o = (rec * H + (int)c * 64 + gbase + kUbs * ub + 2 * tr + (j & 1)) * 16 + tq + 8 * (j >> 1);
              p.tape.c[o] = creg[sq][ub][j];          // pad slots hold exact zeros
              p.tape.hh[o] = hreg[sq][ub][j];
            }
        }
      }
      if (TRAIN && ATT) {
        float keep[2][2][2];
#pragma unroll
        for (int sq = 0; sq < 2; ++sq)
#pragma unroll
          for (int ub = 0; ub < 2; ++ub)
#pragma unroll
            for (int u2 = 0; u2 < 2; ++u2) {
              float sum = hreg[sq][ub][u2] + hreg[sq][ub][2 + u2];
              sum += __shfl_xor_sync(0xffffffffu, sum, 4);
              sum += __shfl_xor_sync(0xffffffffu, sum, 8);
              sum += __shfl_xor_sync(0xffffffffu, sum, 16);
              keep[sq][ub][u2] = hreg[sq][ub][2 + u2];
              if (pad_hi) hreg[sq][ub][2 + u2] = sum;
              if (pad_hi && valid[sq])
                p.tape.s[(((size_t)blockIdx.y * T + ta) * p.B + bseq[sq]) * H + (int)c * 64 + gbase + kUbs * ub + 2 * tr + u2] = sum;
            }
        store_units(hreg);
#pragma unroll
        for (int sq = 0; sq < 2; ++sq)
#pragma unroll
          for (int ub = 0; ub < 2; ++ub)
#pragma unroll
            for (int u2 = 0; u2 < 2; ++u2) hreg[sq][ub][2 + u2] = keep[sq][ub][u2];
      } else if (TRAIN) {
        tape_hp(ta);
        store_units(hreg, tnext);
      }
      if (et == 0) TC_TRACE(0, 2);

      if (!ATT) {
        if (!TRAIN && C > 2 && p.earlypub) publish_half(BAR_H, 1, (int)b); else publish_block(BAR_H, (int)b);
        emit(ta, hreg);                 // global stores of y_t after the hand-off: off the recurrence's critical path
        if (et == 0) TC_TRACE(0, 11);
        continue;
      }
      if (!TRAIN && C > 2 && p.earlypub) publish_half(BAR_HHAT, 1, (int)b); else publish_block(BAR_HHAT, (int)b);
      if (et == 0) TC_TRACE(0, 4);
      // ---- q = relu(Wa . sum_n hy): rows 15 of the A1 accumulator, columns [64,128)
      ptx::mbar_wait(&bars[BAR_ATT_FULL], t & 1);
      // rows 15 are about to be overwritten with q in EVERY CTA: all A1 GEMMs of the cluster must have read them
      ptx::mbar_wait(&bars[BAR_A1FREE], t & 1);
      ptx::tc_fence_after();
      if (et == 0) TC_TRACE(0, 5);
      {
        float v[16];
        ptx::tmem_ld16(tmem_row + b * 256 + 64 + ubase, v);
        if ((lane & 15) == 15) {
          const int qrow = qd * 32 + lane;                  // = 16*seq + 15
          if (TRAIN) {
            const int rs = tile * kSeqTile + 2 * qd + (lane >> 4);
            if (rs < p.B) {
              float* qp = p.tape.q + (((size_t)blockIdx.y * T + ta) * p.B + rs) * H + (int)c * 64 + ubase;
#pragma unroll
              for (int i = 0; i < 16; ++i) qp[i] = fmaxf(v[i], 0.f);
            }
          }
#pragma unroll
          for (int g8 = 0; g8 < 2; ++g8) {                  // 8 consecutive units = one 16-byte K chunk of the row
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint16_t h0, l0, h1, l1;
              split_bits<SPLIT>(fmaxf(v[g8 * 8 + 2 * j], 0.f), h0, l0);
              split_bits<SPLIT>(fmaxf(v[g8 * 8 + 2 * j + 1], 0.f), h1, l1);
              hw[j] = (uint32_t)h0 | ((uint32_t)h1 << 16);
              lw[j] = (uint32_t)l0 | ((uint32_t)l1 << 16);
            }
            const uint4 hq = make_uint4(hw[0], hw[1], hw[2], hw[3]), lq = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            const uint32_t off = img_off((int)c * 64 + ubase + g8 * 8, qrow);
            *reinterpret_cast<uint4*>(hbuf + off) = hq;
            if (SPLIT) *reinterpret_cast<uint4*>(hbuf + (size_t)H * 256 + off) = lq;
#pragma unroll 1
            for (uint32_t peer = 0; peer < (uint32_t)C; ++peer) {
              if (peer == c) continue;
              ptx::st_remote_v4(hbuf + off, peer, hq);
              if (SPLIT) ptx::st_remote_v4(hbuf + (size_t)H * 256 + off, peer, lq);
            }
          }
        }
      }
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      ptx::named_bar_sync(1, kEpiThreads);
      if (warp0 && ptx::elect_one()) {
        // one cluster-scope release fence, then relaxed arrives (a release arrive per peer costs a MEMBAR.GPU each)
        if (C > 1) ptx::fence_acq_rel_cluster();
        for (uint32_t peer = 0; peer < (uint32_t)C; ++peer) ptx::mbar_arrive_remote_relaxed(&bars[BAR_Q + c], peer);
        TC_TRACE(0, 6);
      }
      // ---- e = tanh(Wh hy + Wq q + bs),  partial a = e . u over this warp's 16 units (lane = row)
      ptx::mbar_wait(&bars[BAR_ATT2_FULL], t & 1);
      ptx::tc_fence_after();
      if (et == 0) TC_TRACE(0, 7);
      {
        float part = 0.f;
        float vh[16], vq[16];
        ptx::tmem_ld16(tmem_row + b * 256 + ubase, vh);
        ptx::tmem_ld16(tmem_row + b * 256 + 128 + ubase, vq);
        const int rs3 = tile * kSeqTile + 2 * qd + (lane >> 4);
        float* ep = TRAIN ? p.tape.e + ((((size_t)blockIdx.y * T + ta) * p.B + rs3) * H + (int)c * 64 + ubase) * 16 + (lane & 15) : nullptr;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float wq = __shfl_sync(0xffffffffu, vq[i], (lane & 16) | 15);
          const float ev = fast_tanh(vh[i] + wq + bss[ubase + i]);
          if (TRAIN && rs3 < p.B) ep[i * 16] = (lane & 15) < kNodes ? ev : 0.f;
          part = fmaf(ev, us[ubase + i], part);
        }
        ahalf[ug * 128 + qd * 32 + lane] = part;
      }
      ptx::tc_fence_before();
      ptx::named_bar_sync(1, kEpiThreads);
      if (et == 0) ptx::mbar_arrive(&bars[BAR_ACC_EMPTY + b]);
      if (et < 128) apart[c * 128 + et] = (ahalf[et] + ahalf[128 + et]) + (ahalf[256 + et] + ahalf[384 + et]);
      ptx::fence_proxy_async();
      ptx::named_bar_sync(1, kEpiThreads);
      if (warp0 && ptx::elect_one()) {
        for (uint32_t peer = 0; peer < (uint32_t)C; ++peer)
          if (peer != c) ptx::bulk_s2remote(apart + c * 128, 512, &bars[BAR_A], peer);
        ptx::mbar_arrive_expect_tx(&bars[BAR_A], (uint32_t)(C - 1) * 512);
      }
      if (et == 0) TC_TRACE(0, 8);
      ptx::mbar_wait(&bars[BAR_A], t & 1);
      if (et == 0) TC_TRACE(0, 9);
      // a[row] = 1 + sigmoid(sum over chunks + bu[node])   (hy + hy * a_t, net_aagc.py:212-213); one thread per row
      if (et < 128) {
        float a = bus[et & 15];
        for (int src = 0; src < C; ++src) a += apart[src * 128 + et];
        const float sg = fast_sigmoid(a);
        ahalf[et] = 1.0f + sg;
        if (TRAIN && c == 0) {
          const int rs = tile * kSeqTile + (et >> 4);
          if (rs < p.B) p.tape.a[(((size_t)blockIdx.y * T + ta) * p.B + rs) * 16 + (et & 15)] = (et & 15) < kNodes ? sg : 0.f;
        }
      }
      ptx::named_bar_sync(1, kEpiThreads);
      // ---- h' = hy (1 + a): next step's operand and y_t = act(h')
#pragma unroll
      for (int sq = 0; sq < 2; ++sq) {
        const int r0 = 16 * (2 * qd + sq) + tq;
        const float alo = ahalf[r0], ahi = pad_hi ? 0.f : ahalf[r0 + 8];
#pragma unroll
        for (int ub = 0; ub < 2; ++ub)
#pragma unroll
          for (int j = 0; j < 4; ++j) hreg[sq][ub][j] *= (j & 2) ? ahi : alo;
      }
      if (TRAIN) tape_hp(ta);
      store_units(hreg, tnext);
      if (et == 0) TC_TRACE(0, 10);
      if (!TRAIN && C > 1 && (SPLIT ? p.rescale : p.rescale_bf16)) {
        // h' = hy (1 + a) differs from hy by a factor per ROW, and every CTA already holds the hy blocks of its peers (they
        // were exchanged for the attention GEMM) and a[row]: instead of a second all-gather over DSMEM -- measured 8-10 % of
        // the step at H = 256, the copies also delay the weight stream queued behind them -- each CTA rescales the peers'
        // blocks in its own operand image.  (hi + lo) is exact in fp32, so the operand differs from split(hy (1 + a)) by
        // one fp32 rounding (2^-24).  Blocks become ready in the rotated order the MMA warp walks them.
        ptx::fence_proxy_async();
        ptx::named_bar_sync(1, kEpiThreads);
        if (et == 0) ptx::mbar_arrive(&bars[BAR_H + c]);                // own block (from the fp32 registers): usable at once
        const int rr = et & 127, cg = et >> 7;                           // row, and which 2 of a block's 8 K-chunks
        const float ar = ahalf[rr];
        for (int i = 1; i < C; ++i) {
          const int src = ((int)c + i) % C;
          ptx::mbar_wait(&bars[BAR_HHAT + src], t & 1);                  // the block has landed (already consumed by A1)
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            uint8_t* ph = hbuf + (size_t)((src * 8 + cg * 2 + h2) * kRows + rr) * 16;
            const uint4 hv = *reinterpret_cast<const uint4*>(ph);
            const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
            uint32_t ho[4];
            if constexpr (SPLIT) {
              uint8_t* pl = ph + (size_t)H * 256;
              const uint4 lv = *reinterpret_cast<const uint4*>(pl);
              const uint32_t lw[4] = {lv.x, lv.y, lv.z, lv.w};
              uint32_t lo[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw[j]));
                const float2 lf = __half22float2(*reinterpret_cast<const __half2*>(&lw[j]));
                ptx::split_pair_f16((hf.x + lf.x) * ar, (hf.y + lf.y) * ar, ho[j], lo[j]);
              }
              *reinterpret_cast<uint4*>(pl) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            } else {                                                     // bf16 (opt-in): one extra bf16 rounding per element
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 hf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hw[j]));
                const __nv_bfloat162 bb = __floats2bfloat162_rn(hf.x * ar, hf.y * ar);
                ho[j] = *reinterpret_cast<const uint32_t*>(&bb);
              }
            }
            *reinterpret_cast<uint4*>(ph) = make_uint4(ho[0], ho[1], ho[2], ho[3]);
          }
          ptx::fence_proxy_async();
          ptx::named_bar_sync(1, kEpiThreads);
          if (et == 0) ptx::mbar_arrive(&bars[BAR_H + src]);
        }
      } else {
        publish_block(BAR_H, -1);
      }
      emit(ta, hreg);                   // global stores of y_t after the hand-off: off the recurrence's critical path
      if (et == 0) TC_TRACE(0, 11);
    }
    // final state: h' and c' after the last step (registers)
#pragma unroll
    for (int sq = 0; sq < 2; ++sq) {
      if (!valid[sq]) continue;
#pragma unroll
      for (int ub = 0; ub < 2; ++ub)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int node = tq + 8 * (j >> 1), unit = gbase + kUbs * ub + 2 * tr + (j & 1);
          if (node >= kNodes) continue;
          const size_t gi = ((size_t)bseq[sq] * kNodes + node) * H + c * 64 + unit;
          if (d.cT != nullptr) d.cT[gi] = creg[sq][ub][j];
          if (d.hT != nullptr) d.hT[gi] = hreg[sq][ub][j];
        }
    }
    // the last publish must have landed everywhere before any CTA may exit
    for (int src = 0; src < C; ++src) ptx::mbar_wait(&bars[BAR_H + src], T & 1);
  }

  // ------------------------------------------------------------------ teardown
  ptx::tc_fence_before();
  __syncthreads();
  if (C > 1) ptx::cluster_sync_all();
  if (warp == 1) ptx::tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------
// operand-image packing
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint16_t part_bits(float v, int part, bool split) {
  if (!split) return __bfloat16_as_ushort(__float2bfloat16_rn(v));
  const __half hi = __float2half_rn(v);
  if (part == 0) return __half_as_ushort(hi);
  return __half_as_ushort(__float2half_rn(v - __half2float(hi)));
}

struct TcPacked {
  uint16_t* wg_img; uint16_t* a1_img; uint16_t* a2_img;
  float* P; float* bias4; float* bs; float* u; float* bu;
};

__global__ void tc_pack_weights_kernel(a3gc_cell_params cp, TcPacked out, int F, int H, int variant, int split) {
  const int NP = split ? 2 : 1;
  const int K = F + H, KB = K / 16, KH = H / 16, C = H / 64;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // gate weights: index = ((((c*KB + kb)*NP + part)*2 + kc)*256 + r)*8 + e
  const int64_t n_wg = (int64_t)C * KB * NP * 2 * 256 * 8;
  for (int64_t i = tid; i < n_wg; i += stride) {
    const int e = (int)(i & 7), r = (int)((i >> 3) & 255), kc = (int)((i >> 11) & 1);
    int64_t rest = i >> 12;
    const int part = (int)(rest % NP); rest /= NP;
    const int kb = (int)(rest % KB); const int c = (int)(rest / KB);
    const int gate = r >> 6, j = c * 64 + (r & 63), k = kb * 16 + kc * 8 + e;
    out.wg_img[i] = part_bits(cp.gcn_kernel[gate][(size_t)j * K + k], part, split);
  }
  for (int64_t i = tid; i < (int64_t)H * 4; i += stride) out.bias4[i] = cp.gcn_bias[i & 3][i >> 2];
  for (int64_t i = tid; i < 4 * 256; i += stride) {
    const int gg = (int)(i / 256), mm = (int)((i % 256) / 16), nn = (int)(i % 16);
    float v = 0.f;
    if (mm < kNodes && nn < kNodes)
      v = (variant == A3GC_VARIANT_AGC) ? cp.adjacency[0][nn * kNodes + mm] : cp.adjacency[gg][mm * kNodes + nn];
    out.P[i] = v;
  }
  if (cp.attention_w != nullptr) {
    // A1 image: index = ((((c*KH + kb)*NP + part)*2 + kc)*128 + r)*8 + e ; r < 64: attention_wh row, else attention_w row
    const int64_t n_a1 = (int64_t)C * KH * NP * 2 * 128 * 8;
    for (int64_t i = tid; i < n_a1; i += stride) {
      const int e = (int)(i & 7), r = (int)((i >> 3) & 127), kc = (int)((i >> 10) & 1);
      int64_t rest = i >> 11;
      const int part = (int)(rest % NP); rest /= NP;
      const int kb = (int)(rest % KH); const int c = (int)(rest / KH);
      const float* w = r < 64 ? cp.attention_wh : cp.attention_w;
      out.a1_img[i] = part_bits(w[(size_t)(c * 64 + (r & 63)) * H + kb * 16 + kc * 8 + e], part, split);
    }
    // A2 image (attention_wq): index = ((((c*KH + kb)*NP + part)*2 + kc)*64 + ul)*8 + e
    const int64_t n_a2 = (int64_t)C * KH * NP * 2 * 64 * 8;
    for (int64_t i = tid; i < n_a2; i += stride) {
      const int e = (int)(i & 7), ul = (int)((i >> 3) & 63), kc = (int)((i >> 9) & 1);
      int64_t rest = i >> 10;
      const int part = (int)(rest % NP); rest /= NP;
      const int kb = (int)(rest % KH); const int c = (int)(rest / KH);
      out.a2_img[i] = part_bits(cp.attention_wq[(size_t)(c * 64 + ul) * H + kb * 16 + kc * 8 + e], part, split);
    }
    for (int64_t i = tid; i < H; i += stride) { out.bs[i] = cp.attention_bs[i]; out.u[i] = cp.attention_u[i]; }
    for (int64_t i = tid; i < 16; i += stride) out.bu[i] = i < kNodes ? cp.attention_bu[i] : 0.f;
  }
}

// x [B,T,15,F] fp32 (strided) -> x_img [tiles][T][F/16][NP][2][128][8]
__global__ void tc_pack_x_kernel(const float* __restrict__ x, int64_t sxb, int64_t sxt, uint16_t* __restrict__ img,
                                 int B, int T, int F, int split, int64_t total) {
  const int NP = split ? 2 : 1;
  const int KF = F / 16;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int e = (int)(i & 7), row = (int)((i >> 3) & 127), kc = (int)((i >> 10) & 1);
    int64_t rest = i >> 11;
    const int part = (int)(rest % NP); rest /= NP;
    const int kb = (int)(rest % KF); rest /= KF;
    const int t = (int)(rest % T); const int tile = (int)(rest / T);
    const int b = tile * kSeqTile + (row >> 4), n = row & 15, k = kb * 16 + kc * 8 + e;
    float v = 0.f;
    if (b < B && n < kNodes) v = x[(size_t)b * sxb + (size_t)t * sxt + (size_t)n * F + k];
    img[i] = part_bits(v, part, split);
  }
}

size_t tc_dir_bytes(int F, int H, int NP) {
  size_t b = 0;
  b += align_up((size_t)(F + H) * 4 * H * NP * 2, 256);     // wg_img
  b += align_up((size_t)2 * H * H * NP * 2, 256);            // a1_img
  b += align_up((size_t)H * H * NP * 2, 256);                // a2_img
  b += align_up((size_t)(1024 + 4 * H + 2 * H + 16) * 4, 256);
  return b;
}

TcPacked tc_carve(char* base, int F, int H, int NP) {
  TcPacked p;
  p.wg_img = reinterpret_cast<uint16_t*>(base); base += align_up((size_t)(F + H) * 4 * H * NP * 2, 256);
  p.a1_img = reinterpret_cast<uint16_t*>(base); base += align_up((size_t)2 * H * H * NP * 2, 256);
  p.a2_img = reinterpret_cast<uint16_t*>(base); base += align_up((size_t)H * H * NP * 2, 256);
  float* f = reinterpret_cast<float*>(base);
  p.P = f; f += 1024;
  p.bias4 = f; f += 4 * H;
  p.bs = f; f += H;
  p.u = f; f += H;
  p.bu = f;
  return p;
}

}  // namespace

bool tc_layer_supported(int variant, int f_in, int hidden, int precision) {
  if (variant < A3GC_VARIANT_AAGC || variant > A3GC_VARIANT_GGRU) return false;
  if (hidden != 64 && hidden != 128 && hidden != 256) return false;
  if (f_in <= 0 || f_in % 16 != 0) return false;
  return precision == A3GC_PREC_FP32 || precision == A3GC_PREC_BF16;
}

size_t tc_image_bytes(int64_t batch, int64_t steps, int features, int precision) {
  const int NP = precision == A3GC_PREC_FP32 ? 2 : 1;
  const int64_t tiles = (batch + kSeqTile - 1) / kSeqTile;
  return align_up((size_t)tiles * steps * features * kRows * NP * 2, 256);
}

size_t tc_layer_workspace_bytes(int variant, int64_t batch, int64_t steps, int f_in, int hidden, int num_dirs, int precision) {
  const int NP = precision == A3GC_PREC_FP32 ? 2 : 1;
  size_t b = variant == A3GC_VARIANT_GGRU ? tc_gru_weights_bytes(f_in, hidden, num_dirs, precision)
                                          : (size_t)num_dirs * tc_dir_bytes(f_in, hidden, NP);
  b += tc_image_bytes(batch, steps, f_in, precision);   // x image (unused when the caller hands one in)
  if (variant != A3GC_VARIANT_GGRU) {                    // scratch of the state exchange through L2 (LSTM family, clusters only)
    const int64_t tiles = (batch + kSeqTile - 1) / kSeqTile;
    b += (size_t)num_dirs * tiles * (hidden / 64) * NP * (8 * kRows * 16);
  }
  return b + 256;
}

size_t tc_packed_weights_bytes(int variant, int f_in, int hidden, int num_dirs, int precision) {
  const int NP = precision == A3GC_PREC_FP32 ? 2 : 1;
  return variant == A3GC_VARIANT_GGRU ? tc_gru_weights_bytes(f_in, hidden, num_dirs, precision) : (size_t)num_dirs * tc_dir_bytes(f_in, hidden, NP);
}

// operand images, mix matrices and biases of every direction -> packed (tc_packed_weights_bytes)
int tc_pack_weights(int variant, int num_dirs, const a3gc_cell_params* cells, int f_in, int hidden, int precision, void* packed,
                    cudaStream_t stream) {
  if (variant == A3GC_VARIANT_GGRU) return tc_gru_pack_weights(num_dirs, cells, f_in, hidden, precision, static_cast<char*>(packed), stream);
  const bool split = precision == A3GC_PREC_FP32;
  const int NP = split ? 2 : 1;
  const size_t dir_bytes = tc_dir_bytes(f_in, hidden, NP);
  for (int d = 0; d < num_dirs; ++d) {
    TcPacked pk = tc_carve(static_cast<char*>(packed) + d * dir_bytes, f_in, hidden, NP);
    tc_pack_weights_kernel<<<148, 256, 0, stream>>>(cells[d], pk, f_in, hidden, variant, split ? 1 : 0);
    A3GC_LAUNCH_CHECK("tc_pack_weights_kernel");
  }
  return A3GC_OK;
}

int tc_layer_forward(const LayerArgs& a, void* ws, size_t ws_bytes, cudaStream_t stream) {
  const int F = a.f_in, H = a.hidden;
  const bool split = a.precision == A3GC_PREC_FP32;
  const int NP = split ? 2 : 1;
  const size_t need = tc_layer_workspace_bytes(a.variant, a.batch, a.steps, F, H, a.num_dirs, a.precision);
  if (ws == nullptr || ws_bytes < need) {
    set_error("a3gc_layer_forward (tc): workspace too small (%zu < %zu bytes)", ws_bytes, need);
    return A3GC_ERR_WORKSPACE;
  }
  if (a.steps == 0) return A3GC_OK;
  if (a.variant == A3GC_VARIANT_GGRU) {
    char* wb = static_cast<char*>(ws);
    const uint16_t* ximg = a.x_img;
    if (ximg == nullptr) {
      uint16_t* own = reinterpret_cast<uint16_t*>(wb + tc_gru_weights_bytes(F, H, a.num_dirs, a.precision));
      ximg = own;
      const int64_t tl = (a.batch + kSeqTile - 1) / kSeqTile;
      const int64_t total = tl * a.steps * (F / 16) * NP * 2 * kRows * 8;
      int64_t blocks = (total + 255) / 256;
      if (blocks > 148 * 32) blocks = 148 * 32;
      tc_pack_x_kernel<<<(unsigned)blocks, 256, 0, stream>>>(a.x, a.x_stride_b, a.x_stride_t, own, (int)a.batch, (int)a.steps, F, split ? 1 : 0, total);
      A3GC_LAUNCH_CHECK("tc_pack_x_kernel");
    }
    return tc_gru_layer_launch(a, ximg, wb, stream);
  }
  const bool att = a.variant != A3GC_VARIANT_AAGC;
  const int C = H / 64;
  const int64_t tiles = (a.batch + kSeqTile - 1) / kSeqTile;
  char* base = static_cast<char*>(ws);
  TcLayerParams p;
  memset(&p, 0, sizeof(p));
  const size_t dir_bytes = tc_dir_bytes(F, H, NP);
  char* wbase = base;
  if (a.packed == nullptr) {
    int rc = tc_pack_weights(a.variant, a.num_dirs, a.cells, F, H, a.precision, base, stream);
    if (rc) return rc;
  } else {
    wbase = const_cast<char*>(static_cast<const char*>(a.packed));
  }
  for (int d = 0; d < a.num_dirs; ++d) {
    TcPacked pk = tc_carve(wbase + d * dir_bytes, F, H, NP);
    TcDir& td = p.d[d];
    td.wg_img = pk.wg_img; td.a1_img = pk.a1_img; td.a2_img = pk.a2_img; td.P = pk.P; td.bias4 = pk.bias4;
    td.bs = pk.bs; td.u = pk.u; td.bu = pk.bu;
    td.h0 = a.h0[d]; td.c0 = a.c0[d]; td.hT = a.hT[d]; td.cT = a.cT[d]; td.reverse = a.reverse[d];
  }
  const uint16_t* x_img = a.x_img;
  if (x_img == nullptr) {
    uint16_t* own = reinterpret_cast<uint16_t*>(base + a.num_dirs * dir_bytes);
    x_img = own;
    const int64_t total = tiles * a.steps * (F / 16) * NP * 2 * kRows * 8;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    tc_pack_x_kernel<<<(unsigned)blocks, 256, 0, stream>>>(a.x, a.x_stride_b, a.x_stride_t, own, (int)a.batch, (int)a.steps, F, split ? 1 : 0, total);
    A3GC_LAUNCH_CHECK("tc_pack_x_kernel");
  }
  p.x_img = x_img;
  p.y = a.y; p.syb = a.y_stride_b; p.syt = a.y_stride_t; p.yld = a.y_ld;
  p.y_img = a.y_img; p.y_kf = a.y_img_f / 16;
  p.B = (int)a.batch; p.T = (int)a.steps; p.F = F; p.H = H; p.out_act = a.out_act; p.C = C;
  p.trace = getenv("A3GC_TC_TRACE") != nullptr ? 1 : 0;
  {
    // static placement of the x-part of step t+1 around the attention GEMMs of step t: ~45 % of its K blocks while the
    // epilogue warps run the gate phase, ~35 % during the q hand-off, the rest during e.u / h' / the state exchange
    const int KF = F / 16;
    int n1 = (KF * 45 + 99) / 100, n2 = KF * 35 / 100;   // measured best all-round split (45 % / 35 % / 20 %)
    // per-shape optimum of the sweep with the lean MMA issue loop (profiles/r02_knob_sweep_lean_issue.log; the curve is flat:
    // 1-3 % between the best and 45/35)
    if (H == 256 && F < 512) { n1 = KF * 35 / 100; n2 = KF * 45 / 100; }
    else if (H == 128 && F < 256) { n1 = KF * 20 / 100; n2 = KF * 40 / 100; }
    else if (H == 64) { n1 = KF * 10 / 100; n2 = KF * 30 / 100; }
    if (const char* e = getenv("A3GC_TC_SPLIT")) {
      int a = 0, b = 0;
      if (sscanf(e, "%d,%d", &a, &b) == 2) { n1 = KF * a / 100; n2 = KF * b / 100; }
    }
    if (n1 > KF) n1 = KF;
    if (n1 + n2 > KF) n2 = KF - n1;
    p.n1 = n1; p.n2 = n2;
    p.xsplit = getenv("A3GC_TC_XSPLIT") ? atoi(getenv("A3GC_TC_XSPLIT")) : 0;
    p.xdefer = getenv("A3GC_TC_XDEFER") ? atoi(getenv("A3GC_TC_XDEFER")) : 0;
    // L2 prefetch of the next step's x image: no effect while the issue thread was the bottleneck, -3.7 % at F512:H256 and
    // -1.6 % at F256:H128 since (neutral elsewhere)
    p.xprefetch = getenv("A3GC_TC_XPREFETCH") ? atoi(getenv("A3GC_TC_XPREFETCH")) : 1;
    p.acoll = getenv("A3GC_TC_ACOLL") ? atoi(getenv("A3GC_TC_ACOLL")) : 1;
    p.nprod = getenv("A3GC_TC_NPROD") ? atoi(getenv("A3GC_TC_NPROD")) : kProducers;
    if (p.nprod < 1) p.nprod = 1;
    if (p.nprod > kProducers) p.nprod = kProducers;
    p.earlypub = getenv("A3GC_TC_EARLYPUB") ? atoi(getenv("A3GC_TC_EARLYPUB")) : 1;
    p.puborder = getenv("A3GC_TC_PUBORDER") ? atoi(getenv("A3GC_TC_PUBORDER")) : 1;
    p.rescale = getenv("A3GC_TC_RESCALE") ? atoi(getenv("A3GC_TC_RESCALE")) : 1;
    p.rescale_bf16 = getenv("A3GC_TC_RESCALE_BF16") ? atoi(getenv("A3GC_TC_RESCALE_BF16")) : 0;
    // state exchange through L2: on for 4-CTA clusters.  Inference: the third producer warp is the exchange warp (-1 % at H = 256
    // for A3GC / AGC, -3..5 % for AAGC); training forward (two all-gathers per step, issued in line by the first epilogue warp):
    // -8 % per H = 256 step.  Neutral for 2-CTA clusters, which keep the DSMEM copies.
    p.xchg = getenv("A3GC_TC_XCHG") ? atoi(getenv("A3GC_TC_XCHG")) : (C > 2 ? 1 : 0);
    p.xchg_buf = reinterpret_cast<uint8_t*>(base + a.num_dirs * dir_bytes + tc_image_bytes(a.batch, a.steps, F, a.precision));
    if (p.xchg && a.tape == nullptr && C > 2 && p.earlypub && p.nprod > 2) p.nprod = 2;   // the third producer warp runs the exchange
    p.pubbytes = 16384;  // = kHBlock; A3GC_TC_PUBBYTES < 16384 is a timing diagnostic (truncated state exchange, wrong results)
    if (const char* e = getenv("A3GC_TC_PUBBYTES")) { const int v = atoi(e); if (v >= 16 && v <= 16384 && v % 16 == 0) p.pubbytes = v; }
    p.chunk = 1 << 20;   // measured: one bulk copy per operand is fastest (4 KB pieces: -7 %, 2 KB pieces: -30 %)
    if (const char* e = getenv("A3GC_TC_CHUNK")) { const int v = atoi(e); if (v >= 1024 && v % 16 == 0) p.chunk = v; }
  }

  int dev = 0, smem_max = 0;
  A3GC_CUDA_TRY(cudaGetDevice(&dev));
  A3GC_CUDA_TRY(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  // two rings (weights: every stage; x image: x stages only): as many slot PAIRS as fit, then the remainder goes to single
  // slots (H = 256, fp32: 4 weight slots of 16 KB + 3 x slots of 8 KB next to the 128 KB state image)
  const size_t wslot = (size_t)NP * 2 * 256 * 16, xslot = (size_t)NP * 2 * 128 * 16;
  const size_t fixed = (size_t)NP * H * 256 + tc_fixed_smem_bytes(C, split);
  int S = kMaxStages;
  if (const char* e = getenv("A3GC_TC_STAGES")) { const int v = atoi(e); if (v >= 2 && v <= kMaxStages) S = v; }
  while (S > 1 && fixed + (size_t)S * (wslot + xslot) > (size_t)smem_max) --S;
  if (fixed + (size_t)S * (wslot + xslot) > (size_t)smem_max || S < 2) {
    set_error("tc engine: shared memory budget exceeded (hidden=%d)", H);
    return A3GC_ERR_UNSUPPORTED;
  }
  int SW = S, SX = S;
  if (split && getenv("A3GC_TC_STAGES") == nullptr) {   // bf16 mode: one ring of S (weights + x) slots
    while (SW < kMaxStages && fixed + (size_t)(SW + 1) * wslot + (size_t)SX * xslot <= (size_t)smem_max) ++SW;
    while (SX < kMaxStages && fixed + (size_t)SW * wslot + (size_t)(SX + 1) * xslot <= (size_t)smem_max) ++SX;
  }
  if (split) {
    if (const char* e = getenv("A3GC_TC_WSTAGES")) { const int v = atoi(e); if (v >= 2 && v <= kMaxStages) SW = v; }
    if (const char* e = getenv("A3GC_TC_XSTAGES")) { const int v = atoi(e); if (v >= 2 && v <= kMaxStages) SX = v; }
  }
  if (fixed + (size_t)SW * wslot + (size_t)SX * xslot > (size_t)smem_max) {
    set_error("tc engine: A3GC_TC_WSTAGES / A3GC_TC_XSTAGES exceed the shared memory budget (hidden=%d)", H);
    return A3GC_ERR_INVALID_ARG;
  }
  p.S = SW; p.SX = SX;
  const size_t smem = fixed + (size_t)SW * wslot + (size_t)SX * xslot;

  const bool train = a.tape != nullptr;
  if (train) { p.tape = *a.tape; p.hmask = a.hmask; }
  void (*kern)(const TcLayerParams);
  if (train) kern = att ? tc_lstm_layer_kernel<true, true, true> : tc_lstm_layer_kernel<true, false, true>;   // training: fp32-parity path only
  else kern = split ? (att ? tc_lstm_layer_kernel<true, true, false> : tc_lstm_layer_kernel<true, false, false>)
                    : (att ? tc_lstm_layer_kernel<false, true, false> : tc_lstm_layer_kernel<false, false, false>);
  A3GC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(tiles * C), (unsigned)a.num_dirs, 1);
  cfg.blockDim = dim3(kThreadsTC, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  A3GC_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, p));
  A3GC_LAUNCH_CHECK("tc_lstm_layer_kernel");
  return A3GC_OK;
}

}  // namespace a3gc

// debug: copy the per-phase clock64 timeline of CTA (0,0) of the last traced launch (2*16*16 uint64)
extern "C" int a3gc_debug_read_tc_trace(unsigned long long* host_out) {
  using namespace a3gc;
  if (!host_out) return A3GC_ERR_INVALID_ARG;
  if (const char* e = getenv("A3GC_TC_TRACE")) if (strcmp(e, "gru") == 0) return tc_gru_read_trace(host_out);
  A3GC_CUDA_TRY(cudaMemcpyFromSymbol(host_out, g_tc_trace, sizeof(unsigned long long) * 2 * 16 * 16));
  return A3GC_OK;
}

// debug / tuning: how many clusters of `cluster_size` CTAs of the fp32-split attention kernel (with `smem_bytes` of
// dynamic shared memory each) the device can hold at once
extern "C" int a3gc_debug_max_active_clusters(int cluster_size, int smem_bytes) {
  using namespace a3gc;
  void (*kern)(const TcLayerParams) = tc_lstm_layer_kernel<true, true, false>;
  A3GC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  if (cluster_size > 8) A3GC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(cluster_size * 64), 2, 1);
  cfg.blockDim = dim3(kThreadsTC, 1, 1);
  cfg.dynamicSmemBytes = (size_t)smem_bytes;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cluster_size; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int n = 0;
  A3GC_CUDA_TRY(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
  return n;
}

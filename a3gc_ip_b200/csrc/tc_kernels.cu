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

namespace a3gc {
namespace {

constexpr int kRows = 128;             // 8 sequences x 16 node slots
constexpr int kSeqTile = 8;
constexpr int kUnits = 64;             // hidden units per CTA
constexpr int kSub = 32;               // staged accumulator columns per sub-chunk (8 units x 4 gates)
constexpr int kPitch = 132;            // floats per staged column
constexpr int kEpiThreads = 256;       // warps 2..9
constexpr int kThreadsTC = 64 + kEpiThreads;
constexpr int kMaxStages = 4;

struct TcDir {
  const uint16_t* wg_img;   // [C][(F+H)/16][NP][2][256][8]   gate weights, rows = 4*unit + gate
  const uint16_t* wh_img;   // [C][H/16][NP][2][64][8]        attention_wh rows of this chunk
  const float* P;           // [4][16][16]  zero padded mixing matrices  z = P_g u
  const float* bias4;       // [H][4]
  const float* wa_t;        // [H][H]  attention_w  transposed (k-major)
  const float* wq_t;        // [H][H]  attention_wq transposed
  const float* bs;          // [H]
  const float* u;           // [H]
  const float* bu;          // [16]
  const float* h0; const float* c0; float* hT; float* cT;
  int reverse;
};
struct TcLayerParams {
  TcDir d[2];
  const uint16_t* x_img;    // [tiles][T][F/16][NP][2][128][8]
  float* y; int64_t syb, syt, yld;
  int B, T, F, H, out_act, C, S;
};

// barrier slots in shared memory
enum { BAR_FULL = 0, BAR_EMPTY = kMaxStages, BAR_ACC_FULL = 2 * kMaxStages, BAR_ATT_FULL = BAR_ACC_FULL + 2,
       BAR_ACC_EMPTY, BAR_H = BAR_ACC_EMPTY + 2, BAR_HHAT, BAR_Q, BAR_A, BAR_HFREE, BAR_COUNT };

__host__ __device__ inline size_t tc_fixed_smem_bytes(int C) {
  return (size_t)kSub * kPitch * 4      // staging
         + (size_t)C * 512 * 4 * 2      // sbuf, qbuf
         + (size_t)C * 128 * 4          // apart
         + 512 * 4 + 256 * 4            // wqbuf, ahalf
         + 256 * 4 + 64 * 4 * 2 + 16 * 4  // bias4, bs, u, bu
         + 32 * 8 + 16;                 // barriers, tmem slot
}

__device__ __forceinline__ void split_store(uint8_t* hbuf, int H, bool split, int k, int row, float v) {
  // element (row, k) of the operand image [part][K/8][128][8]
  const size_t off = ((size_t)(k >> 3) * kRows + row) * 16 + (size_t)(k & 7) * 2;
  if (split) {
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
    *reinterpret_cast<__half*>(hbuf + off) = hi;
    *reinterpret_cast<__half*>(hbuf + (size_t)H * 256 + off) = lo;
  } else {
    *reinterpret_cast<__nv_bfloat16*>(hbuf + off) = __float2bfloat16_rn(v);
  }
}

template <bool SPLIT, bool ATT>
__global__ void __launch_bounds__(kThreadsTC, 1)
tc_lstm_layer_kernel(const TcLayerParams p) {
  constexpr int NP = SPLIT ? 2 : 1;
  constexpr uint32_t kBBytes = NP * 2 * 256 * 16;     // one K=16 block of gate weights (all parts)
  constexpr uint32_t kABytes = NP * 2 * 128 * 16;     // one K=16 block of x rows
  constexpr uint32_t kStageBytes = kBBytes + kABytes;
  constexpr uint32_t kHBlock = 8 * kRows * 16;         // this CTA's 64 units of one operand part: 8 K-chunks
  extern __shared__ __align__(1024) uint8_t smem[];

  const int C = p.C, S = p.S, H = p.H, F = p.F, T = p.T;
  const uint32_t c = C > 1 ? ptx::cluster_ctarank() : 0u;
  const int tile = blockIdx.x / C;
  const TcDir& d = p.d[blockIdx.y];
  const int KF = F / 16, KH = H / 16;
  const int warp = threadIdx.x >> 5;

  uint8_t* hbuf = smem;
  uint8_t* ring = hbuf + (size_t)NP * H * 256;
  float* staging = reinterpret_cast<float*>(ring + (size_t)S * kStageBytes);
  float* sbuf = staging + kSub * kPitch;     // [C][8][64]
  float* qbuf = sbuf + C * 512;              // [C][8][64]
  float* apart = qbuf + C * 512;             // [C][128]
  float* wqbuf = apart + C * 128;            // [8][64]
  float* ahalf = wqbuf + 512;                // [2][128]
  float* bias4s = ahalf + 256;               // [64][4]
  float* bss = bias4s + 256;                 // [64]
  float* us = bss + 64;                      // [64]
  float* bus = us + 64;                      // [16]
  uint64_t* bars = reinterpret_cast<uint64_t*>(bus + 16);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 32);

  // ------------------------------------------------------------------ setup
  if (threadIdx.x == 0) {
    for (int i = 0; i < BAR_COUNT; ++i) ptx::mbar_init(&bars[i], i == BAR_HFREE ? (uint32_t)C : 1u);
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, 512);
  {
    uint4* z = reinterpret_cast<uint4*>(hbuf);
    const int n16 = NP * H * 16;
    for (int i = threadIdx.x; i < n16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < 256; i += blockDim.x) bias4s[i] = d.bias4[(size_t)c * 256 + i];
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

  if (warp == 0) {
    // ================================================================ producer: weights / x -> ring
    if ((threadIdx.x & 31) == 0) {
      uint32_t it = 0;
      auto load_stage = [&](const void* bsrc, uint32_t bbytes, const void* asrc, uint32_t abytes) {
        const uint32_t st = it % S, ph = (it / S) & 1u;
        ptx::mbar_wait(&bars[BAR_EMPTY + st], ph ^ 1u);
        ptx::mbar_arrive_expect_tx(&bars[BAR_FULL + st], bbytes + abytes);
        uint8_t* dst = ring + (size_t)st * kStageBytes;
        ptx::bulk_g2s(dst, bsrc, bbytes, &bars[BAR_FULL + st]);
        if (abytes) ptx::bulk_g2s(dst + kBBytes, asrc, abytes, &bars[BAR_FULL + st]);
        ++it;
      };
      const uint8_t* wg = reinterpret_cast<const uint8_t*>(d.wg_img) + (size_t)c * (KF + KH) * kBBytes;
      const uint8_t* wh = reinterpret_cast<const uint8_t*>(d.wh_img) + (size_t)c * KH * (NP * 2 * 64 * 16);
      const uint8_t* xi = reinterpret_cast<const uint8_t*>(p.x_img);
      auto xpart = [&](int t) {
        const int ta = d.reverse ? T - 1 - t : t;
        const uint8_t* xs = xi + ((size_t)tile * T + ta) * KF * kABytes;
        for (int kb = 0; kb < KF; ++kb) load_stage(wg + (size_t)kb * kBBytes, kBBytes, xs + (size_t)kb * kABytes, kABytes);
      };
      xpart(0);
      for (int t = 0; t < T; ++t) {
        for (int kb = 0; kb < KH; ++kb) load_stage(wg + (size_t)(KF + kb) * kBBytes, kBBytes, nullptr, 0);
        if (t + 1 < T) xpart(t + 1);
        if (ATT)
          for (int s4 = 0; s4 < KH / 4; ++s4) load_stage(wh + (size_t)s4 * 4 * (NP * 2 * 64 * 16), 4 * NP * 2 * 64 * 16, nullptr, 0);
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer (one thread)
    if ((threadIdx.x & 31) == 0) {
      const uint32_t idesc256 = ptx::make_idesc_f16(128, 256, !SPLIT);
      const uint32_t idesc64 = ptx::make_idesc_f16(128, 64, !SPLIT);
      const uint32_t hbase = ptx::smem_u32(hbuf);
      const uint32_t hpart = (uint32_t)H * 256;
      uint32_t it = 0;
      uint32_t empty_k[2] = {0, 0};     // completions of BAR_ACC_EMPTY[b] consumed so far
      auto wait_stage = [&]() -> uint32_t {
        const uint32_t st = it % S, ph = (it / S) & 1u;
        ptx::mbar_wait(&bars[BAR_FULL + st], ph);
        ptx::tc_fence_after();
        return ptx::smem_u32(ring + (size_t)st * kStageBytes);
      };
      auto release_stage = [&]() { ptx::umma_commit(&bars[BAR_EMPTY + it % S]); ++it; };
      // one K=16 block of the gate GEMM: A (parts at a0, a0+astride), B parts at b0, b0 + kBBytes/NP
      auto gate_block = [&](uint32_t dcol, uint32_t a0, uint32_t astride, uint32_t b0, bool first) {
        const uint64_t ah = ptx::make_smem_desc(a0, kRows * 16, 128);
        const uint64_t bh = ptx::make_smem_desc(b0, 256 * 16, 128);
        ptx::umma_f16(tmem + dcol, ah, bh, idesc256, first ? 0u : 1u);
        if (SPLIT) {
          const uint64_t al = ptx::make_smem_desc(a0 + astride, kRows * 16, 128);
          const uint64_t bl = ptx::make_smem_desc(b0 + kBBytes / 2, 256 * 16, 128);
          ptx::umma_f16(tmem + dcol, al, bh, idesc256, 1u);
          ptx::umma_f16(tmem + dcol, ah, bl, idesc256, 1u);
        }
      };
      auto xpart = [&](uint32_t dcol) {
        for (int kb = 0; kb < KF; ++kb) {
          const uint32_t st = wait_stage();
          gate_block(dcol, st + kBBytes, kABytes / NP, st, kb == 0);
          release_stage();
        }
      };
      xpart(0);
      for (int t = 0; t < T; ++t) {
        const uint32_t b = t & 1, dcol = b * 256;
        // h-part of step t (needs h'_{t-1} of every chunk in local shared memory)
        ptx::mbar_wait_cluster(&bars[BAR_H], t & 1);
        ptx::tc_fence_after();
        for (int kb = 0; kb < KH; ++kb) {
          const uint32_t st = wait_stage();
          gate_block(dcol, hbase + (uint32_t)kb * 2 * kRows * 16, hpart, st, false);
          release_stage();
        }
        ptx::umma_commit(&bars[BAR_ACC_FULL + b]);
        if (C > 1) ptx::umma_commit_multicast(&bars[BAR_HFREE], cta_mask); else ptx::umma_commit(&bars[BAR_HFREE]);
        // x-part of step t+1 into the other buffer (free once the epilogue of step t-1 has drained it)
        if (t + 1 < T) {
          if (t >= 1) {
            const uint32_t bo = b ^ 1u;
            if (ATT) { ptx::mbar_wait(&bars[BAR_ACC_EMPTY + bo], 1u); empty_k[bo] += 1; }
            else { ptx::mbar_wait(&bars[BAR_ACC_EMPTY + bo], empty_k[bo] & 1u); empty_k[bo] += 1; }
            ptx::tc_fence_after();
          }
          xpart((b ^ 1u) * 256);
        }
        if (ATT) {
          // attention GEMM  E[128,64] = hy[128,H] * Wh_c[64,H]^T into columns [0,64) of the drained buffer
          ptx::mbar_wait_cluster(&bars[BAR_HHAT], t & 1);
          ptx::mbar_wait(&bars[BAR_ACC_EMPTY + b], 0u);
          ptx::tc_fence_after();
          for (int s4 = 0; s4 < KH / 4; ++s4) {
            const uint32_t st = wait_stage();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int kb = s4 * 4 + j;
              const uint32_t a0 = hbase + (uint32_t)kb * 2 * kRows * 16;
              const uint32_t b0 = st + (uint32_t)j * (NP * 2 * 64 * 16);
              const uint64_t ah = ptx::make_smem_desc(a0, kRows * 16, 128);
              const uint64_t bh = ptx::make_smem_desc(b0, 64 * 16, 128);
              ptx::umma_f16(tmem + dcol, ah, bh, idesc64, (s4 | j) ? 1u : 0u);
              if (SPLIT) {
                const uint64_t al = ptx::make_smem_desc(a0 + hpart, kRows * 16, 128);
                const uint64_t bl = ptx::make_smem_desc(b0 + 2 * 64 * 16, 64 * 16, 128);
                ptx::umma_f16(tmem + dcol, al, bh, idesc64, 1u);
                ptx::umma_f16(tmem + dcol, ah, bl, idesc64, 1u);
              }
            }
            release_stage();
          }
          ptx::umma_commit(&bars[BAR_ATT_FULL]);
        }
      }
    }
  } else {
    // ================================================================ epilogue warps (256 threads)
    const int et = threadIdx.x - 64;
    const int ew = et >> 5, lane = et & 31;
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may read
    const int chalf = ew >> 2;                    // which half of the columns this warp stages
    const int g = lane & 3, hf = (lane >> 2) & 3, uh = lane >> 4;
    const int m = 4 * hf + g;                     // node owned for the pointwise update
    const int s = ew;                             // sequence within the tile
    const int bseq = tile * kSeqTile + s;
    const bool valid = bseq < p.B && m < kNodes;
    const int row = 16 * s + m;
    const int ycol = blockIdx.y * H + (int)c * 64;

    // mixing rows held in registers: Pr[j][n] = P_g[4*hf + j][n]
    float Pr[4][15];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int n = 0; n < 15; ++n) Pr[j][n] = d.P[(g * 16 + 4 * hf + j) * 16 + n];

    float creg[32], hreg[32];                     // index q*4+i  <->  unit 8q + 4uh + i
#pragma unroll
    for (int qi = 0; qi < 32; ++qi) {
      const int ul = 8 * (qi >> 2) + 4 * uh + (qi & 3);
      const size_t gi = ((size_t)bseq * kNodes + m) * H + c * 64 + ul;
      creg[qi] = (valid && d.c0 != nullptr) ? d.c0[gi] : 0.f;
      hreg[qi] = (valid && d.h0 != nullptr) ? d.h0[gi] : 0.f;
    }

    // write this thread's 32 state values into the local operand image, then send this CTA's 64-unit
    // block to every peer; `bar` completes in each CTA when all C blocks have landed
    auto publish_h = [&](int bar, const float* extra, uint32_t extra_bytes) {
      if (m < kNodes) {
#pragma unroll
        for (int qi = 0; qi < 32; ++qi) {
          const int ul = 8 * (qi >> 2) + 4 * uh + (qi & 3);
          split_store(hbuf, H, SPLIT, (int)c * 64 + ul, row, hreg[qi]);
        }
      }
      ptx::fence_proxy_async();
      ptx::named_bar_sync(1, kEpiThreads);
      if (et == 0) {
        for (uint32_t peer = 0; peer < (uint32_t)C; ++peer) {
          if (peer == c) continue;
          for (int part = 0; part < NP; ++part)
            ptx::bulk_s2remote(hbuf + (size_t)part * H * 256 + (size_t)c * kHBlock, kHBlock, &bars[bar], peer);
          if (extra_bytes) ptx::bulk_s2remote(const_cast<float*>(extra), extra_bytes, &bars[bar], peer);
        }
        ptx::mbar_arrive_expect_tx(&bars[bar], (uint32_t)(C - 1) * (NP * kHBlock + extra_bytes));
      }
    };
    auto publish_small = [&](int bar, float* buf, uint32_t bytes) {
      ptx::fence_proxy_async();
      ptx::named_bar_sync(1, kEpiThreads);
      if (et == 0) {
        for (uint32_t peer = 0; peer < (uint32_t)C; ++peer)
          if (peer != c) ptx::bulk_s2remote(buf, bytes, &bars[bar], peer);
        ptx::mbar_arrive_expect_tx(&bars[bar], (uint32_t)(C - 1) * bytes);
      }
    };

    publish_h(BAR_H, nullptr, 0);                 // completion #0 of BAR_H: h_{-1} = h0

    for (int t = 0; t < T; ++t) {
      const uint32_t b = t & 1;
      const int ta = d.reverse ? T - 1 - t : t;
      // ---------------------------------------------------------------- gates -> c', hy
      ptx::mbar_wait(&bars[BAR_ACC_FULL + b], (t >> 1) & 1);
      ptx::tc_fence_after();
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        {   // phase a: TMEM -> staging[col][row]
          float v[16];
          ptx::tmem_ld16(tmem + ((uint32_t)(quarter * 32) << 16) + b * 256 + q * kSub + chalf * 16, v);
          float* dst = staging + (chalf * 16) * kPitch + quarter * 32 + lane;
#pragma unroll
          for (int i = 0; i < 16; ++i) dst[i * kPitch] = v[i];
        }
        ptx::named_bar_sync(1, kEpiThreads);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int ul8 = 4 * uh + i;                       // unit within the sub-chunk
          const float4* sp = reinterpret_cast<const float4*>(staging + (4 * ul8 + g) * kPitch + 16 * s);
          float un[16];
#pragma unroll
          for (int r4 = 0; r4 < 4; ++r4) {
            const float4 v4 = sp[r4];
            un[4 * r4] = v4.x; un[4 * r4 + 1] = v4.y; un[4 * r4 + 2] = v4.z; un[4 * r4 + 3] = v4.w;
          }
          float z[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float a = 0.f;
#pragma unroll
            for (int n = 0; n < 15; ++n) a = fmaf(Pr[j][n], un[n], a);
            z[j] = a;
          }
          // 4x4 transpose over the lanes of one gate group: afterwards zz[gate] is for node m = 4hf + g
          float zz[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int jsend = (g - r) & 3;
            const float send = jsend == 0 ? z[0] : (jsend == 1 ? z[1] : (jsend == 2 ? z[2] : z[3]));
            const float got = __shfl_sync(0xffffffffu, send, (lane & ~3) | ((g + r) & 3));
            const int gate = (g + r) & 3;
            if (gate == 0) zz[0] = got; else if (gate == 1) zz[1] = got; else if (gate == 2) zz[2] = got; else zz[3] = got;
          }
          const float4 bias = reinterpret_cast<const float4*>(bias4s)[8 * q + ul8];
          const float ig = sigmoidf_(zz[0] + bias.x);
          const float fg = sigmoidf_(zz[1] + bias.y);
          const float cg = tanhf_(zz[2] + bias.z);
          const float og = sigmoidf_(zz[3] + bias.w);
          const float cn = valid ? fmaf(fg, creg[q * 4 + i], ig * cg) : 0.f;
          creg[q * 4 + i] = cn;
          hreg[q * 4 + i] = valid ? og * tanhf_(cn) : 0.f;
        }
        if (q == 7) ptx::tc_fence_before();
        ptx::named_bar_sync(1, kEpiThreads);
      }
      if (et == 0) ptx::mbar_arrive(&bars[BAR_ACC_EMPTY + b]);
      // every CTA of the cluster has finished reading h'_{t-1}: the operand images may be overwritten
      ptx::mbar_wait_cluster(&bars[BAR_HFREE], t & 1);

      if (ATT) {
        // node sums of hy for the attention query (net_aagc.py:200)
#pragma unroll
        for (int qi = 0; qi < 32; ++qi) {
          float v = hreg[qi];
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          if ((lane & 15) == 0) sbuf[c * 512 + s * 64 + 8 * (qi >> 2) + 4 * uh + (qi & 3)] = v;
        }
        publish_h(BAR_HHAT, sbuf + c * 512, 2048);
        ptx::mbar_wait_cluster(&bars[BAR_HHAT], t & 1);
        // q = relu(sum_n(hy) W_a^T)   this CTA's 64 units, two per thread
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int ul = lane + 32 * r;
          const float* wcol = d.wa_t + c * 64 + ul;
          float a0 = 0.f, a1 = 0.f;
          for (int k = 0; k < H; k += 2) {
            a0 = fmaf(sbuf[(k >> 6) * 512 + s * 64 + (k & 63)], __ldg(wcol + (size_t)k * H), a0);
            a1 = fmaf(sbuf[((k + 1) >> 6) * 512 + s * 64 + ((k + 1) & 63)], __ldg(wcol + (size_t)(k + 1) * H), a1);
          }
          qbuf[c * 512 + s * 64 + ul] = fmaxf(a0 + a1, 0.f);
        }
        publish_small(BAR_Q, qbuf + c * 512, 2048);
        ptx::mbar_wait_cluster(&bars[BAR_Q], t & 1);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int ul = lane + 32 * r;
          const float* wcol = d.wq_t + c * 64 + ul;
          float a0 = bss[ul], a1 = 0.f;
          for (int k = 0; k < H; k += 2) {
            a0 = fmaf(qbuf[(k >> 6) * 512 + s * 64 + (k & 63)], __ldg(wcol + (size_t)k * H), a0);
            a1 = fmaf(qbuf[((k + 1) >> 6) * 512 + s * 64 + ((k + 1) & 63)], __ldg(wcol + (size_t)(k + 1) * H), a1);
          }
          wqbuf[s * 64 + ul] = a0 + a1;
        }
        ptx::named_bar_sync(1, kEpiThreads);
        // e = tanh(Wh hy + Wq q + bs),  partial a = e . u over this CTA's 64 units (lane = row)
        ptx::mbar_wait(&bars[BAR_ATT_FULL], t & 1);
        ptx::tc_fence_after();
        {
          float v[32];
          ptx::tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + b * 256 + chalf * 32, v);
          const int r_ = quarter * 32 + lane;
          const float* wq = wqbuf + (r_ >> 4) * 64 + chalf * 32;
          float part = 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) part = fmaf(tanhf_(v[i] + wq[i]), us[chalf * 32 + i], part);
          ahalf[chalf * 128 + r_] = part;
        }
        ptx::tc_fence_before();
        ptx::named_bar_sync(1, kEpiThreads);
        if (et == 0) ptx::mbar_arrive(&bars[BAR_ACC_EMPTY + b]);
        if (et < 128) apart[c * 128 + et] = ahalf[et] + ahalf[128 + et];
        publish_small(BAR_A, apart + c * 128, 512);
        ptx::mbar_wait_cluster(&bars[BAR_A], t & 1);
        float a = 0.f;
        if (m < kNodes) {
          a = bus[m];
          for (int src = 0; src < C; ++src) a += apart[src * 128 + row];
          a = 1.0f + sigmoidf_(a);                          // hy + hy * a_t  (net_aagc.py:212-213)
        }
#pragma unroll
        for (int qi = 0; qi < 32; ++qi) hreg[qi] *= a;
      }
      // ---------------------------------------------------------------- outputs and next-step operand
      if (valid) {
        float* yp = p.y + (size_t)bseq * p.syb + (size_t)ta * p.syt + (size_t)m * p.yld + ycol;
#pragma unroll
        for (int qi = 0; qi < 32; ++qi)
          yp[8 * (qi >> 2) + 4 * uh + (qi & 3)] = apply_act(hreg[qi], p.out_act);
      }
      publish_h(BAR_H, nullptr, 0);
    }
    if (valid) {
#pragma unroll
      for (int qi = 0; qi < 32; ++qi) {
        const int ul = 8 * (qi >> 2) + 4 * uh + (qi & 3);
        const size_t gi = ((size_t)bseq * kNodes + m) * H + c * 64 + ul;
        if (d.hT != nullptr) d.hT[gi] = hreg[qi];
        if (d.cT != nullptr) d.cT[gi] = creg[qi];
      }
    }
    // the last publish must have landed everywhere before any CTA may exit
    ptx::mbar_wait_cluster(&bars[BAR_H], T & 1);
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
  uint16_t* wg_img; uint16_t* wh_img;
  float* P; float* bias4; float* wa_t; float* wq_t; float* bs; float* u; float* bu;
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
    const int gate = r & 3, j = c * 64 + (r >> 2), k = kb * 16 + kc * 8 + e;
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
    // attention_wh: index = ((((c*KH + kb)*NP + part)*2 + kc)*64 + ul)*8 + e
    const int64_t n_wh = (int64_t)C * KH * NP * 2 * 64 * 8;
    for (int64_t i = tid; i < n_wh; i += stride) {
      const int e = (int)(i & 7), ul = (int)((i >> 3) & 63), kc = (int)((i >> 9) & 1);
      int64_t rest = i >> 10;
      const int part = (int)(rest % NP); rest /= NP;
      const int kb = (int)(rest % KH); const int c = (int)(rest / KH);
      out.wh_img[i] = part_bits(cp.attention_wh[(size_t)(c * 64 + ul) * H + kb * 16 + kc * 8 + e], part, split);
    }
    for (int64_t i = tid; i < (int64_t)H * H; i += stride) {
      const int k = (int)(i / H), j = (int)(i % H);
      out.wa_t[i] = cp.attention_w[(size_t)j * H + k];
      out.wq_t[i] = cp.attention_wq[(size_t)j * H + k];
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
  b += align_up((size_t)H * H * NP * 2, 256);                // wh_img
  b += align_up((size_t)(1024 + 4 * H + 2 * (size_t)H * H + 2 * H + 16) * 4, 256);
  return b;
}

TcPacked tc_carve(char* base, int F, int H, int NP) {
  TcPacked p;
  p.wg_img = reinterpret_cast<uint16_t*>(base); base += align_up((size_t)(F + H) * 4 * H * NP * 2, 256);
  p.wh_img = reinterpret_cast<uint16_t*>(base); base += align_up((size_t)H * H * NP * 2, 256);
  float* f = reinterpret_cast<float*>(base);
  p.P = f; f += 1024;
  p.bias4 = f; f += 4 * H;
  p.wa_t = f; f += (size_t)H * H;
  p.wq_t = f; f += (size_t)H * H;
  p.bs = f; f += H;
  p.u = f; f += H;
  p.bu = f;
  return p;
}

}  // namespace

bool tc_layer_supported(int variant, int f_in, int hidden, int precision) {
  if (variant != A3GC_VARIANT_AAGC && variant != A3GC_VARIANT_A3GC && variant != A3GC_VARIANT_AGC) return false;
  if (hidden != 64 && hidden != 128 && hidden != 256) return false;
  if (f_in <= 0 || f_in % 16 != 0) return false;
  return precision == A3GC_PREC_FP32 || precision == A3GC_PREC_BF16;
}

size_t tc_layer_workspace_bytes(int variant, int64_t batch, int64_t steps, int f_in, int hidden, int num_dirs, int precision) {
  (void)variant;
  const int NP = precision == A3GC_PREC_FP32 ? 2 : 1;
  const int64_t tiles = (batch + kSeqTile - 1) / kSeqTile;
  size_t b = (size_t)num_dirs * tc_dir_bytes(f_in, hidden, NP);
  b += align_up((size_t)tiles * steps * f_in * kRows * NP * 2, 256);   // x image
  return b + 256;
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
  const bool att = a.variant != A3GC_VARIANT_AAGC;
  const int C = H / 64;
  const int64_t tiles = (a.batch + kSeqTile - 1) / kSeqTile;
  char* base = static_cast<char*>(ws);
  TcLayerParams p;
  memset(&p, 0, sizeof(p));
  const size_t dir_bytes = tc_dir_bytes(F, H, NP);
  for (int d = 0; d < a.num_dirs; ++d) {
    TcPacked pk = tc_carve(base + d * dir_bytes, F, H, NP);
    tc_pack_weights_kernel<<<148, 256, 0, stream>>>(a.cells[d], pk, F, H, a.variant, split ? 1 : 0);
    A3GC_LAUNCH_CHECK("tc_pack_weights_kernel");
    TcDir& td = p.d[d];
    td.wg_img = pk.wg_img; td.wh_img = pk.wh_img; td.P = pk.P; td.bias4 = pk.bias4; td.wa_t = pk.wa_t; td.wq_t = pk.wq_t;
    td.bs = pk.bs; td.u = pk.u; td.bu = pk.bu;
    td.h0 = a.h0[d]; td.c0 = a.c0[d]; td.hT = a.hT[d]; td.cT = a.cT[d]; td.reverse = a.reverse[d];
  }
  uint16_t* x_img = reinterpret_cast<uint16_t*>(base + a.num_dirs * dir_bytes);
  {
    const int64_t total = tiles * a.steps * (F / 16) * NP * 2 * kRows * 8;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    tc_pack_x_kernel<<<(unsigned)blocks, 256, 0, stream>>>(a.x, a.x_stride_b, a.x_stride_t, x_img, (int)a.batch, (int)a.steps, F, split ? 1 : 0, total);
    A3GC_LAUNCH_CHECK("tc_pack_x_kernel");
  }
  p.x_img = x_img;
  p.y = a.y; p.syb = a.y_stride_b; p.syt = a.y_stride_t; p.yld = a.y_ld;
  p.B = (int)a.batch; p.T = (int)a.steps; p.F = F; p.H = H; p.out_act = a.out_act; p.C = C;

  int dev = 0, smem_max = 0;
  A3GC_CUDA_TRY(cudaGetDevice(&dev));
  A3GC_CUDA_TRY(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  const size_t stage_bytes = (size_t)NP * (2 * 256 * 16 + 2 * 128 * 16);
  const size_t fixed = (size_t)NP * H * 256 + tc_fixed_smem_bytes(C) + 1024;
  int S = kMaxStages;
  while (S > 1 && fixed + (size_t)S * stage_bytes > (size_t)smem_max) --S;
  if (fixed + (size_t)S * stage_bytes > (size_t)smem_max || S < 2) {
    set_error("tc engine: shared memory budget exceeded (hidden=%d)", H);
    return A3GC_ERR_UNSUPPORTED;
  }
  p.S = S;
  const size_t smem = fixed + (size_t)S * stage_bytes;

  void (*kern)(const TcLayerParams) =
      split ? (att ? tc_lstm_layer_kernel<true, true> : tc_lstm_layer_kernel<true, false>)
            : (att ? tc_lstm_layer_kernel<false, true> : tc_lstm_layer_kernel<false, false>);
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

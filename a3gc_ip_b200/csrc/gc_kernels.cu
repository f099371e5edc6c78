// AAGC graph convolution y = act((adj @ x) @ W^T + b)  (net_aagc.py:61-66), the non-recurrent layers
// around the recurrent ones (linear_in: 12/15 -> H, linear_out: 2H -> 3/9).  Both are HBM-bound:
//   gc_in   small f_in (<= 32): one CTA per 8 frames; writes either fp32 [frames,15,O] (coalesced float4)
//           or directly the UMMA operand image of the first recurrent layer (fp16 hi/lo or bf16), so the
//           activation never makes an fp32 round trip through HBM on the tensor-core path;
//   gc_out  large f_in, f_out <= 16: 8 rows per warp, K split over the lanes with 128-bit coalesced loads,
//           W read from shared memory once per 8 rows, shuffle reduction, then the 15x15 mix per frame.
#include "common.cuh"
#include <cuda_fp16.h>
#include <cuda_bf16.h>

namespace a3gc {
namespace {

constexpr int kGcThreads = 256;

__device__ __forceinline__ void split16(float v, bool split, uint16_t& hi, uint16_t& lo) {
  if (split) {
    const __half h = __float2half_rn(v);
    hi = __half_as_ushort(h);
    lo = __half_as_ushort(__float2half_rn(v - __half2float(h)));
  } else {
    hi = __bfloat16_as_ushort(__float2bfloat16_rn(v));
    lo = 0;
  }
}

// ------------------------------------------------------------------------------------------
// gc_in: f_in <= 32.  Block = 8 frames.  IMG = false: frames f0..f0+7 -> y fp32.
// RAW = true: the input is not a materialised [frames,15,K] tensor but the raw IMU frame (GcRawInput, common.cuh):
// prepare_input (evaluate_a3gc_tp.py:64-94: normalise, drop IMU 6, scatter onto nodes [3,4,13,14,10]) and the stage
// concatenation cat(x, pos) (:168, :170) happen in the load, 288 (+180) bytes per frame instead of 720 / 900; the ten
// zero nodes are skipped in the adjacency mix of the 12 IMU features (adding their exact zeros changes nothing).
// IMG = true: the 8 frames are (tile, t): sequences 8*tile..8*tile+7 at time t -> operand image
// [tiles][T][O/16][NP][2][128][8] (row = 16*seq + node, row 15 of every sequence zero).
// ------------------------------------------------------------------------------------------
// KC > 0: f_in known at compile time (12 = raw IMU frame, 15 = + the previous stage's positions): the projection loops unroll
// fully and the index arithmetic divides by constants; KC = 0: any f_in <= 32.  Same operations in the same order either way.
template <bool IMG, bool RAW, int KC>
__global__ void __launch_bounds__(kGcThreads, 4)
gc_in_kernel(a3gc_gc_params p, const float* __restrict__ x, GcRawInput raw, float* __restrict__ y, uint16_t* __restrict__ img,
             int64_t frames, int B, int T, int Krt, int O, int act, int split) {
  const int K = KC > 0 ? KC : Krt;
  extern __shared__ __align__(16) float smem[];
  float* adj = smem;                 // [16][16]
  float* wt = adj + 256;             // [K][O]   W transposed
  float* bias = wt + (size_t)K * O;  // [O]
  float* xs = bias + O;              // [8][15][K]
  const int KP = K | 1;              // odd row stride: consecutive rows of xm fall into different banks
  float* xm = xs + 8 * kNodes * K;   // [128][KP]  (adj @ x), row 15 of each frame zero
  float* nrm = xm + 128 * KP;        // RAW: [72] mean, [72] std in the order of the raw frame (18 acc, 54 ori)
  if (RAW) {
    for (int i = threadIdx.x; i < 72; i += blockDim.x) {
      if (i < 18) {
        nrm[i] = raw.acc_mean != nullptr ? raw.acc_mean[i] : 0.f;
        nrm[72 + i] = raw.acc_std != nullptr ? raw.acc_std[i] : 1.f;
      } else {
        nrm[i] = raw.ori_mean != nullptr ? raw.ori_mean[i - 18] : 0.f;
        nrm[72 + i] = raw.ori_std != nullptr ? raw.ori_std[i - 18] : 1.f;
      }
    }
  }
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    const int m = i >> 4, n = i & 15;
    adj[i] = (m < kNodes && n < kNodes) ? p.adj[m * kNodes + n] : 0.f;
  }
  for (int i = threadIdx.x; i < K * O; i += blockDim.x) { const int k = i / O, o = i % O; wt[i] = p.gcn_kernel[(size_t)o * K + k]; }
  for (int i = threadIdx.x; i < O; i += blockDim.x) bias[i] = p.gcn_bias[i];
  const int per_frame = kNodes * K;
  const int NP = split ? 2 : 1;
  const int64_t groups = IMG ? (int64_t)((B + 7) / 8) * T : (frames + 7) / 8;
  for (int64_t grp = blockIdx.x; grp < groups; grp += gridDim.x) {
    __syncthreads();
    // ---- load the 8 frames of this group
    int64_t tile = 0; int t = 0;
    if (IMG) { tile = grp / T; t = (int)(grp % T); }
    if (RAW) {
      // coalesced: the 72 raw floats of each frame (18 acc + 54 ori) are read once, normalised from the shared-memory copy of
      // the statistics and scattered to (node, feature); IMU 6 is dropped.  The ten zero nodes' IMU features are never
      // written: the mix below reads nodes 3, 4, 10, 13, 14 only for k < 12.  Three independent loads per thread.
#pragma unroll
      for (int it = 0; it < 3; ++it) {
        const int i = (int)threadIdx.x + it * kGcThreads;
        if (i < 8 * 72) {
          const int fr = i / 72, c = i % 72;
          int64_t f = -1;
          if (IMG) {
            const int64_t b = tile * 8 + fr;
            if (b < B) f = b * T + t;
          } else {
            if (grp * 8 + fr < frames) f = grp * 8 + fr;
          }
          float v = 0.f;
          int imu, k;
          if (c < 18) {
            imu = c / 3; k = c % 3;
            if (f >= 0) {
              v = __ldg(raw.acc + (size_t)f * 18 + c);
              if (raw.acc_mean != nullptr) v = (v - nrm[c]) / nrm[72 + c];
            }
          } else {
            const int ch = c - 18;
            imu = ch / 9; k = 3 + ch % 9;
            if (f >= 0) {
              v = __ldg(raw.ori + (size_t)f * 54 + ch);
              if (raw.ori_mean != nullptr) v = (v - nrm[c]) / nrm[72 + c];
            }
          }
          if (imu < 5 && k < K) {
            const int node = imu == 0 ? 3 : imu == 1 ? 4 : imu == 2 ? 13 : imu == 3 ? 14 : 10;   // input_joints (evaluate_a3gc_tp.py:65)
            xs[fr * per_frame + node * K + k] = v;
          }
        }
      }
      const int PK = K > 12 ? K - 12 : 1;                 // features taken from the previous stage's output (3; none at K = 12)
      if (K > 12) {
        const int per = kNodes * PK;
        for (int i = threadIdx.x; i < 8 * per; i += blockDim.x) {
          const int fr = i / per, e = i % per, n = e / PK, j = e % PK;
          int64_t f = -1;
          if (IMG) {
            const int64_t b = tile * 8 + fr;
            if (b < B) f = b * T + t;
          } else {
            if (grp * 8 + fr < frames) f = grp * 8 + fr;
          }
          xs[fr * per_frame + n * K + 12 + j] = f >= 0 ? __ldg(raw.pos + (size_t)f * (kNodes * 3) + n * 3 + j) : 0.f;
        }
      }
    } else {
      for (int i = threadIdx.x; i < 8 * per_frame; i += blockDim.x) {
        const int fr = i / per_frame, r = i % per_frame;
        int64_t f = -1;                                   // frame index in [B*T) order, -1: beyond the batch
        if (IMG) {
          const int64_t b = tile * 8 + fr;
          if (b < B) f = b * T + t;
        } else {
          if (grp * 8 + fr < frames) f = grp * 8 + fr;
        }
        xs[i] = f >= 0 ? __ldg(x + (size_t)f * per_frame + r) : 0.f;
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 128 * K; i += blockDim.x) {
      const int k = i % K, row = i / K, fr = row >> 4, m = row & 15;
      float s = 0.f;
      if (m < kNodes) {
        const float* col = xs + (size_t)fr * per_frame + k;
        if (RAW && k < 12) {
          // only nodes 3, 4, 10, 13, 14 carry IMU data; same ascending order as the full loop
          s = fmaf(adj[m * 16 + 3], col[3 * K], s);
          s = fmaf(adj[m * 16 + 4], col[4 * K], s);
          s = fmaf(adj[m * 16 + 10], col[10 * K], s);
          s = fmaf(adj[m * 16 + 13], col[13 * K], s);
          s = fmaf(adj[m * 16 + 14], col[14 * K], s);
        } else {
#pragma unroll
          for (int n = 0; n < kNodes; ++n) s = fmaf(adj[m * 16 + n], col[n * K], s);
        }
      }
      xm[(size_t)row * KP + k] = s;
    }
    __syncthreads();
    if (IMG) {
      // item = (8 consecutive outputs, rows r and r + 64): the weights of a k are read once for QR = 2 rows (2 shared-memory
      // wavefronts per row and k instead of 3; QR = 4 spills under the 64-register cap); per row one 16-byte chunk of the image per part
      constexpr int QR = 2, RS = 128 / QR;
      const int chunks = O / 8;
      uint8_t* base = reinterpret_cast<uint8_t*>(img) + ((size_t)tile * T + t) * (size_t)O * 128 * NP * 2;
      for (int i = threadIdx.x; i < chunks * RS; i += blockDim.x) {
        const int r = i % RS, ch = i / RS;
        const float* a = xm + (size_t)r * KP;
        float acc[QR][8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float bj = bias[ch * 8 + j];
#pragma unroll
          for (int q = 0; q < QR; ++q) acc[q][j] = bj;
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const float4 w0 = *reinterpret_cast<const float4*>(wt + (size_t)k * O + ch * 8);
          const float4 w1 = *reinterpret_cast<const float4*>(wt + (size_t)k * O + ch * 8 + 4);
#pragma unroll
          for (int q = 0; q < QR; ++q) {
            const float av = a[(size_t)q * RS * KP + k];
            acc[q][0] = fmaf(av, w0.x, acc[q][0]); acc[q][1] = fmaf(av, w0.y, acc[q][1]); acc[q][2] = fmaf(av, w0.z, acc[q][2]); acc[q][3] = fmaf(av, w0.w, acc[q][3]);
            acc[q][4] = fmaf(av, w1.x, acc[q][4]); acc[q][5] = fmaf(av, w1.y, acc[q][5]); acc[q][6] = fmaf(av, w1.z, acc[q][6]); acc[q][7] = fmaf(av, w1.w, acc[q][7]);
          }
        }
#pragma unroll
        for (int q = 0; q < QR; ++q) {
          const int row = r + RS * q;
          uint16_t hi[8], lo[8];
          const bool live = (row & 15) < kNodes;
#pragma unroll
          for (int j = 0; j < 8; ++j) split16(live ? apply_act(acc[q][j], act) : 0.f, split != 0, hi[j], lo[j]);
          // image offset: [kb = ch/2][part][kc = ch%2][row][8]
          const size_t off = ((((size_t)(ch >> 1) * NP + 0) * 2 + (ch & 1)) * 128 + row) * 16;
          *reinterpret_cast<uint4*>(base + off) = make_uint4((uint32_t)hi[0] | ((uint32_t)hi[1] << 16), (uint32_t)hi[2] | ((uint32_t)hi[3] << 16), (uint32_t)hi[4] | ((uint32_t)hi[5] << 16), (uint32_t)hi[6] | ((uint32_t)hi[7] << 16));
          if (split)
            *reinterpret_cast<uint4*>(base + off + (size_t)2 * 128 * 16) =
                make_uint4((uint32_t)lo[0] | ((uint32_t)lo[1] << 16), (uint32_t)lo[2] | ((uint32_t)lo[3] << 16), (uint32_t)lo[4] | ((uint32_t)lo[5] << 16), (uint32_t)lo[6] | ((uint32_t)lo[7] << 16));
        }
      }
    } else {
      // item = (4 consecutive outputs, frame-row): float4 stores, consecutive threads -> consecutive addresses
      const int quads = (O + 3) / 4;
      const int nrows = 8 * kNodes;
      for (int i = threadIdx.x; i < quads * nrows; i += blockDim.x) {
        const int qd = i % quads, r = i / quads, fr = r / kNodes, m = r % kNodes;
        const int64_t f = grp * 8 + fr;
        if (f >= frames) continue;
        const float* a = xm + (size_t)(fr * 16 + m) * KP;
        float acc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = (qd * 4 + j < O) ? bias[qd * 4 + j] : 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const float av = a[k];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (qd * 4 + j < O) acc[j] = fmaf(av, wt[(size_t)k * O + qd * 4 + j], acc[j]);
        }
        float* yp = y + ((size_t)f * kNodes + m) * O + qd * 4;
        if ((O & 3) == 0) {
          *reinterpret_cast<float4*>(yp) = make_float4(apply_act(acc[0], act), apply_act(acc[1], act), apply_act(acc[2], act), apply_act(acc[3], act));
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (qd * 4 + j < O) yp[j] = apply_act(acc[j], act);
        }
      }
    }
  }
}


// Sum N per-lane values over the 32 lanes of a warp.  Stage MASK: lanes with the bit clear keep the lower half of the values and
// receive the partner's, lanes with the bit set keep the upper half -- (N+1)/2 values remain.  After the last stage value j of
// a lane is the complete sum of the original value butterfly_index(j, lane).
template <int N, int MASK>
__device__ __forceinline__ void xor_butterfly(float* v, int lane) {
  constexpr int H = (N + 1) / 2;
  const bool up = (lane & MASK) != 0;
#pragma unroll
  for (int i = 0; i < H; ++i) {
    const float a = v[i];
    const float b = (i + H < N) ? v[i + H] : 0.f;
    const float send = up ? a : b, keep = up ? b : a;
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, MASK);
  }
  if constexpr (MASK > 1) xor_butterfly<H, MASK / 2>(v, lane);
}
__host__ __device__ constexpr int butterfly_count(int n, int mask) { return mask == 0 ? n : butterfly_count((n + 1) / 2, mask / 2); }
template <int N, int MASK>
__device__ __forceinline__ int butterfly_index(int j, int lane) {
  constexpr int H = (N + 1) / 2;
  int pos = j;
  if constexpr (MASK > 1) {
    pos = butterfly_index<H, MASK / 2>(j, lane);
    if (pos < 0) return -1;
  }
  pos += (lane & MASK) ? H : 0;
  return pos < N ? pos : -1;
}

// ------------------------------------------------------------------------------------------
// gc_out: f_in % 128 == 0, f_out <= 16.  Block = 16 frames (240 rows); each warp takes R rows at a time (8 at f_out <= 4; 4 above:
// half the accumulators, twice the resident warps -- the loads of one warp overlap the reduction of another).
// ------------------------------------------------------------------------------------------
template <int OMAX, int R>
__global__ void __launch_bounds__(kGcThreads, (R == 4 && OMAX <= 9) ? 3 : 2)
gc_out_kernel(a3gc_gc_params p, const float* __restrict__ x, float* __restrict__ y, int64_t frames, int K, int O, int act) {
  extern __shared__ __align__(16) float smem[];
  float* adj = smem;                         // [16][16]
  float* ws = adj + 256;                     // [O][K]
  float* v = ws + (size_t)O * K;             // [16 frames][15][OMAX]
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    const int m = i >> 4, n = i & 15;
    adj[i] = (m < kNodes && n < kNodes) ? p.adj[m * kNodes + n] : 0.f;
  }
  for (int i = threadIdx.x; i < O * K; i += blockDim.x) ws[i] = p.gcn_kernel[i];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t groups = (frames + 15) / 16;
  for (int64_t grp = blockIdx.x; grp < groups; grp += gridDim.x) {
    const int64_t row0 = grp * 16 * kNodes;
    const int64_t nrows = ((frames - grp * 16) < 16 ? (frames - grp * 16) : 16) * kNodes;
    __syncthreads();
    for (int r8 = warp * R; r8 < nrows; r8 += R * (kGcThreads / 32)) {
      float acc[R][OMAX];
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int o = 0; o < OMAX; ++o) acc[r][o] = 0.f;
      for (int k0 = lane * 4; k0 < K; k0 += 128) {
        float4 xv[R];
#pragma unroll
        for (int r = 0; r < R; ++r)
          xv[r] = (r8 + r < nrows) ? __ldg(reinterpret_cast<const float4*>(x + (size_t)(row0 + r8 + r) * K + k0)) : make_float4(0, 0, 0, 0);
#pragma unroll
        for (int o = 0; o < OMAX; ++o) {
          if (o < O) {
            const float4 w = *reinterpret_cast<const float4*>(ws + (size_t)o * K + k0);
#pragma unroll
            for (int r = 0; r < R; ++r)
              acc[r][o] = fmaf(xv[r].x, w.x, fmaf(xv[r].y, w.y, fmaf(xv[r].z, w.z, fmaf(xv[r].w, w.w, acc[r][o]))));
          }
        }
      }
      // transposing butterfly: every stage halves the values a lane carries (71 shuffles for 8 x 9 sums instead of 360); the
      // pairing tree per sum is the xor butterfly 16, 8, 4, 2, 1 as before, so the results are unchanged bit for bit
      float* flat = &acc[0][0];
      xor_butterfly<R * OMAX, 16>(flat, lane);
#pragma unroll
      for (int j = 0; j < butterfly_count(R * OMAX, 16); ++j) {
        const int idx = butterfly_index<R * OMAX, 16>(j, lane);          // = r * OMAX + o of the sum this lane holds
        if (idx >= 0 && r8 + idx / OMAX < nrows) v[(size_t)r8 * OMAX + idx] = flat[j];
      }
    }
    __syncthreads();
    // y[f][m][o] = sum_n adj[m][n] v[f][n][o] + b[o]
    const int outs = (int)(nrows / kNodes) * kNodes * O;
    for (int i = threadIdx.x; i < outs; i += blockDim.x) {
      const int o = i % O, m = (i / O) % kNodes, fr = i / (O * kNodes);
      float s = p.gcn_bias[o];
      const float* vf = v + (size_t)fr * kNodes * OMAX + o;
#pragma unroll
      for (int n = 0; n < kNodes; ++n) s = fmaf(adj[m * 16 + n], vf[n * OMAX], s);
      y[(size_t)(grp * 16) * kNodes * O + i] = apply_act(s, act);
    }
  }
}

template <bool IMG, bool RAW>
int launch_gc_in(int64_t blocks, size_t smem, cudaStream_t stream, const a3gc_gc_params& p, const float* x, const GcRawInput& raw,
                 float* y, uint16_t* img, int64_t frames, int B, int T, int K, int O, int act, int split) {
#define A3GC_GC_IN(KC)                                                                                                             \
  do {                                                                                                                             \
    A3GC_CUDA_TRY(cudaFuncSetAttribute(gc_in_kernel<IMG, RAW, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
    gc_in_kernel<IMG, RAW, KC><<<(unsigned)blocks, kGcThreads, smem, stream>>>(p, x, raw, y, img, frames, B, T, K, O, act, split); \
  } while (0)
  if (K == 12) A3GC_GC_IN(12);
  else if (K == 15) A3GC_GC_IN(15);
  else A3GC_GC_IN(0);
#undef A3GC_GC_IN
  return A3GC_OK;
}

int sm_count() {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

}  // namespace

// 0 = handled here, 1 = shape not covered (caller falls back to the generic kernel)
int gc_forward_fast(const a3gc_gc_params* p, const float* x, float* y, int64_t frames, int f_in, int f_out, int act,
                    cudaStream_t stream, int* handled) {
  *handled = 0;
  if (frames == 0) { *handled = 1; return A3GC_OK; }
  const int sms = sm_count();
  if (f_in <= 32) {
    const size_t smem = (256 + (size_t)f_in * f_out + f_out + 8 * kNodes * f_in + 128 * (f_in | 1) + 144) * sizeof(float);
    if (smem > 160 * 1024) return A3GC_OK;
    int64_t groups = (frames + 7) / 8;
    int64_t blocks = groups < (int64_t)sms * 8 ? groups : (int64_t)sms * 8;
    { const int rc = launch_gc_in<false, false>(blocks, smem, stream, *p, x, GcRawInput{}, y, nullptr, frames, 0, 0, f_in, f_out, act, 0); if (rc != A3GC_OK) return rc; }
    A3GC_LAUNCH_CHECK("gc_in_kernel");
    *handled = 1;
  } else if (f_out <= 16 && f_in % 128 == 0) {
    const size_t smem = (256 + (size_t)f_out * f_in + 16 * kNodes * 16) * sizeof(float);
    if (smem > 200 * 1024) return A3GC_OK;
    int64_t groups = (frames + 15) / 16;
    const int per_sm = (f_out > 4 && f_out <= 9) ? 6 : 4;          // two waves of the 3 (gc_out<9, 4>) or 2 resident CTAs per SM
    int64_t blocks = groups < (int64_t)sms * per_sm ? groups : (int64_t)sms * per_sm;
    if (f_out <= 4) {
      A3GC_CUDA_TRY(cudaFuncSetAttribute(gc_out_kernel<4, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      gc_out_kernel<4, 8><<<(unsigned)blocks, kGcThreads, smem, stream>>>(*p, x, y, frames, f_in, f_out, act);
    } else if (f_out <= 9) {
      A3GC_CUDA_TRY(cudaFuncSetAttribute(gc_out_kernel<9, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      gc_out_kernel<9, 4><<<(unsigned)blocks, kGcThreads, smem, stream>>>(*p, x, y, frames, f_in, f_out, act);
    } else {
      A3GC_CUDA_TRY(cudaFuncSetAttribute(gc_out_kernel<16, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      gc_out_kernel<16, 4><<<(unsigned)blocks, kGcThreads, smem, stream>>>(*p, x, y, frames, f_in, f_out, act);
    }
    A3GC_LAUNCH_CHECK("gc_out_kernel");
    *handled = 1;
  }
  return A3GC_OK;
}

// linear_in fused with the operand-image packing of the first recurrent layer (tensor-core engine); raw != nullptr: also
// fused with prepare_input / the stage concatenation (x is ignored)
int gc_forward_image(const a3gc_gc_params* p, const float* x, const GcRawInput* raw, uint16_t* img, int64_t batch, int64_t steps,
                     int f_in, int f_out, int act, int split, cudaStream_t stream) {
  if (batch == 0 || steps == 0) return A3GC_OK;
  if (f_in > 32 || f_out % 16 != 0) { set_error("gc_forward_image: unsupported shape"); return A3GC_ERR_UNSUPPORTED; }
  const size_t smem = (256 + (size_t)f_in * f_out + f_out + 8 * kNodes * f_in + 128 * (f_in | 1) + 144) * sizeof(float);
  const int64_t groups = ((batch + 7) / 8) * steps;
  const int sms = sm_count();
  int64_t blocks = groups < (int64_t)sms * 8 ? groups : (int64_t)sms * 8;
  if (raw != nullptr) {
    const int rc = launch_gc_in<true, true>(blocks, smem, stream, *p, nullptr, *raw, nullptr, img, batch * steps, (int)batch, (int)steps, f_in, f_out, act, split);
    if (rc != A3GC_OK) return rc;
  } else {
    const int rc = launch_gc_in<true, false>(blocks, smem, stream, *p, x, GcRawInput{}, nullptr, img, batch * steps, (int)batch, (int)steps, f_in, f_out, act, split);
    if (rc != A3GC_OK) return rc;
  }
  A3GC_LAUNCH_CHECK("gc_in_kernel<img>");
  return A3GC_OK;
}

// linear_in from the raw IMU frame to fp32 activations (the engines that do not consume operand images)
int gc_forward_raw(const a3gc_gc_params* p, const GcRawInput* raw, float* y, int64_t frames, int f_in, int f_out, int act,
                   cudaStream_t stream) {
  if (frames == 0) return A3GC_OK;
  const size_t smem = (256 + (size_t)f_in * f_out + f_out + 8 * kNodes * f_in + 128 * (f_in | 1) + 144) * sizeof(float);
  if (f_in > 32 || smem > 160 * 1024) { set_error("gc_forward_raw: unsupported shape"); return A3GC_ERR_UNSUPPORTED; }
  const int sms = sm_count();
  int64_t groups = (frames + 7) / 8;
  int64_t blocks = groups < (int64_t)sms * 8 ? groups : (int64_t)sms * 8;
  { const int rc = launch_gc_in<false, true>(blocks, smem, stream, *p, nullptr, *raw, y, nullptr, frames, 0, 0, f_in, f_out, act, 0); if (rc != A3GC_OK) return rc; }
  A3GC_LAUNCH_CHECK("gc_in_kernel<raw>");
  return A3GC_OK;
}

}  // namespace a3gc

// SIMT (fp32 CUDA-core) engine of the A3GC-IP hot path: works for every variant and every shape,
// is the shape-generic path behind A3GC_ENGINE_SIMT / AUTO, and the on-device cross-check for the
// tcgen05 engine.  Math follows net_aagc.py:61-66 (AAGC), :103-126 / :178-217 / :266-303 (LSTM cells),
// :343-368 (G-GRU) with the graph mix applied on the accumulator side:  P (S W^T) == (P S) W^T.
//
// Data layout in shared memory: every activation operand is kept as [sequence][feature k][16 nodes]
// (node index fastest, padded 15 -> 16 with a zero), so that one (sequence, hidden-unit) task reads
// the 16 node values of feature k as four broadcast 128-bit loads and owns all 15 nodes of its
// column -- which makes the 15x15 adjacency mix a purely in-register operation.
#include "common.cuh"
#include <cuda_bf16.h>

namespace a3gc {
namespace {

constexpr int kThreads = 256;

// ------------------------------------------------------------------------------------------
// weight packing (global -> workspace).  k-major so that consecutive hidden units are contiguous.
// ------------------------------------------------------------------------------------------
struct LstmPacked {
  float4* Wg4;     // [(F+H)][H] float4 = (i, f, c, o) weights of unit j at input feature k
  float* Wa_t;     // [H][H]  Wa_t[k][j] = attention_w [j][k]
  float* Wh_t;     // [H][H]
  float* Wq_t;     // [H][H]
  float* P;        // [4][16][16] zero-padded mixing matrices, P_g[m][n]
  float4* bias4;   // [H]
  float* bs;       // [H]
  float* u;        // [H]
  float* bu;       // [16]
};

struct GruPacked {
  float* Wg_t;     // [H][H]   Wg_t[k][j] = gcn_kernel[j][k]
  float4* Win4;    // [F][H]   (r_in, u_in, c_in, 0)
  float4* Whid4;   // [H][H]   (r_hid, u_hid, 0, c_hid)
  float* P;        // [16][16] P[n][m] = adjacency[m][n]   (used transposed, net_aagc.py:348)
  float4* bias4;   // [H]      (b_r, b_u, b_c, 0)
};

size_t lstm_packed_floats(int F, int H) {
  return (size_t)(F + H) * H * 4 + 3 * (size_t)H * H + 4 * 256 + (size_t)H * 4 + H + H + 16;
}
size_t gru_packed_floats(int F, int H) {
  return (size_t)H * H + (size_t)F * H * 4 + (size_t)H * H * 4 + 256 + (size_t)H * 4;
}

LstmPacked carve_lstm(float* base, int F, int H) {
  LstmPacked p;
  float* q = base;
  p.Wg4 = reinterpret_cast<float4*>(q); q += (size_t)(F + H) * H * 4;
  p.bias4 = reinterpret_cast<float4*>(q); q += (size_t)H * 4;
  p.Wa_t = q; q += (size_t)H * H;
  p.Wh_t = q; q += (size_t)H * H;
  p.Wq_t = q; q += (size_t)H * H;
  p.P = q; q += 4 * 256;
  p.bs = q; q += H;
  p.u = q; q += H;
  p.bu = q; q += 16;
  return p;
}
GruPacked carve_gru(float* base, int F, int H) {
  GruPacked p;
  float* q = base;
  p.Win4 = reinterpret_cast<float4*>(q); q += (size_t)F * H * 4;
  p.Whid4 = reinterpret_cast<float4*>(q); q += (size_t)H * H * 4;
  p.bias4 = reinterpret_cast<float4*>(q); q += (size_t)H * 4;
  p.Wg_t = q; q += (size_t)H * H;
  p.P = q; q += 256;
  return p;
}

__global__ void pack_lstm_kernel(a3gc_cell_params cp, LstmPacked out, int F, int H, int variant) {
  const int K = F + H;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t i = tid; i < (int64_t)K * H; i += stride) {
    int k = (int)(i / H), j = (int)(i % H);
    size_t src = (size_t)j * K + k;
    out.Wg4[i] = make_float4(cp.gcn_kernel[0][src], cp.gcn_kernel[1][src], cp.gcn_kernel[2][src], cp.gcn_kernel[3][src]);
  }
  for (int64_t i = tid; i < H; i += stride)
    out.bias4[i] = make_float4(cp.gcn_bias[0][i], cp.gcn_bias[1][i], cp.gcn_bias[2][i], cp.gcn_bias[3][i]);
  for (int64_t i = tid; i < 4 * 256; i += stride) {
    int g = (int)(i / 256), m = (int)((i % 256) / 16), n = (int)(i % 16);
    float v = 0.f;
    if (m < kNodes && n < kNodes) {
      // A3GC / AAGC: z = adjacency_g @ S (einsum 'bnf,nm->bmf' with adjacency.t(), net_aagc.py:183)
      // AGC:         z = adjacency^T @ S (einsum 'nm,bmf->bnf' with adjacency.t(), net_aagc.py:271)
      v = (variant == A3GC_VARIANT_AGC) ? cp.adjacency[0][n * kNodes + m] : cp.adjacency[g][m * kNodes + n];
    }
    out.P[i] = v;
  }
  if (cp.attention_w != nullptr) {
    for (int64_t i = tid; i < (int64_t)H * H; i += stride) {
      int k = (int)(i / H), j = (int)(i % H);
      size_t src = (size_t)j * H + k;
      out.Wa_t[i] = cp.attention_w[src];
      out.Wh_t[i] = cp.attention_wh[src];
      out.Wq_t[i] = cp.attention_wq[src];
    }
    for (int64_t i = tid; i < H; i += stride) {
      out.bs[i] = cp.attention_bs[i];
      out.u[i] = cp.attention_u[i];
    }
    for (int64_t i = tid; i < 16; i += stride) out.bu[i] = i < kNodes ? cp.attention_bu[i] : 0.f;
  }
}

__global__ void pack_gru_kernel(a3gc_cell_params cp, GruPacked out, int F, int H) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t i = tid; i < (int64_t)F * H; i += stride) {
    int k = (int)(i / H), j = (int)(i % H);
    size_t src = (size_t)j * F + k;
    out.Win4[i] = make_float4(cp.dense_in_w[0][src], cp.dense_in_w[1][src], cp.dense_in_w[2][src], 0.f);
  }
  for (int64_t i = tid; i < (int64_t)H * H; i += stride) {
    int k = (int)(i / H), j = (int)(i % H);
    size_t src = (size_t)j * H + k;
    out.Whid4[i] = make_float4(cp.dense_hid_w[0][src], cp.dense_hid_w[1][src], 0.f, cp.dense_hid_w[2][src]);
    out.Wg_t[i] = cp.g_gcn_kernel[src];
  }
  for (int64_t i = tid; i < H; i += stride)
    out.bias4[i] = make_float4(cp.dense_in_b[0][i], cp.dense_in_b[1][i], cp.dense_in_b[2][i], 0.f);
  for (int64_t i = tid; i < 256; i += stride) {
    int n = (int)(i / 16), m = (int)(i % 16);
    out.P[i] = (m < kNodes && n < kNodes) ? cp.g_adjacency[m * kNodes + n] : 0.f;
  }
}

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
// acc[g][n] += sum_k act[k][n] * w[k].{x,y,z,w}[g]   (act: shared [K][16];  w: global, stride ldw float4)
__device__ __forceinline__ void accum4(float (&acc)[4][16], const float* __restrict__ act,
                                       const float4* __restrict__ w, int K, int ldw) {
#pragma unroll 2
  for (int k = 0; k < K; ++k) {
    const float4 wv = __ldg(w + (size_t)k * ldw);
    const float4* a4 = reinterpret_cast<const float4*>(act + k * kNodesPad);
    float a[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float4 v = a4[q];
      a[4 * q] = v.x; a[4 * q + 1] = v.y; a[4 * q + 2] = v.z; a[4 * q + 3] = v.w;
    }
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      acc[0][n] = fmaf(wv.x, a[n], acc[0][n]);
      acc[1][n] = fmaf(wv.y, a[n], acc[1][n]);
      acc[2][n] = fmaf(wv.z, a[n], acc[2][n]);
      acc[3][n] = fmaf(wv.w, a[n], acc[3][n]);
    }
  }
}

__device__ __forceinline__ void accum1(float (&acc)[16], const float* __restrict__ act,
                                       const float* __restrict__ w, int K, int ldw) {
#pragma unroll 16
  for (int k = 0; k < K; ++k) {
    const float wv = __ldg(w + (size_t)k * ldw);
    const float4* a4 = reinterpret_cast<const float4*>(act + k * kNodesPad);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float4 v = a4[q];
      acc[4 * q] = fmaf(wv, v.x, acc[4 * q]);
      acc[4 * q + 1] = fmaf(wv, v.y, acc[4 * q + 1]);
      acc[4 * q + 2] = fmaf(wv, v.z, acc[4 * q + 2]);
      acc[4 * q + 3] = fmaf(wv, v.w, acc[4 * q + 3]);
    }
  }
}

// out[m] = sum_n P[m][n] * in[n], m, n < 15 (P: shared [16][16], broadcast reads)
__device__ __forceinline__ void mix15(const float (&in)[16], const float* __restrict__ P, float (&out)[16]) {
#pragma unroll
  for (int m = 0; m < kNodes; ++m) {
    const float4* p4 = reinterpret_cast<const float4*>(P + m * 16);
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float4 pv = p4[q];
      s = fmaf(pv.x, in[4 * q], s);
      s = fmaf(pv.y, in[4 * q + 1], s);
      s = fmaf(pv.z, in[4 * q + 2], s);
      s = fmaf(pv.w, in[4 * q + 3], s);   // P[m][15] == 0
    }
    out[m] = s;
  }
  out[15] = 0.f;
}

__device__ __forceinline__ void store16(float* dst, const float (&v)[16]) {
  float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
  for (int q = 0; q < 4; ++q) d4[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}
__device__ __forceinline__ void load16(float (&v)[16], const float* src) {
  const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float4 t = s4[q];
    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
  }
}

// 64-byte rows of the tape in global memory as two 256-bit accesses (sm_100: LDG.256 / STG.256): a warp's 32 rows are 64
// bytes apart, so every access touches 16 lines whatever its width, and the backward chain's pointwise phases are bound by
// exactly those L1 tag lookups -- half as many instructions, half as many lookups.  Rows are 64-byte aligned.
__device__ __forceinline__ void load16g(float (&v)[16], const float* src) {
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(src));
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[8]), "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]) : "l"(src + 8));
}
__device__ __forceinline__ void store16g(float* dst, const float (&v)[16]) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"l"(dst), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"l"(dst + 8), "f"(v[8]), "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]) : "memory");
}

// store16 with the two 8-float halves exchanged when swz: the tensor-core phases of the backward chain read the [k][16]
// arrays as mma B fragments (lane -> k = c, c + 4; node = r), and rows k, k + 2 would otherwise share their banks
__device__ __forceinline__ void store16_swz(float* dst, const float (&v)[16], bool swz) {
  float4* d4 = reinterpret_cast<float4*>(dst);
  const int o = swz ? 2 : 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) d4[q ^ o] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}

struct DirPtrs {
  const float* h0; const float* c0; float* hT; float* cT; int reverse;
};
struct LayerGeom {
  const float* x; int64_t sxb, sxt;
  float* y; int64_t syb, syt, yld;
  int B, T, F, H, out_act, BT;
};

// state [B,15,H] (global) -> [seq][j][16] (shared); zero when src == nullptr or b >= B
__device__ void load_state(float* dst, const float* src, int b0, const LayerGeom& g) {
  for (int i = threadIdx.x; i < g.BT * g.H * kNodesPad; i += blockDim.x) {
    int n = i % kNodesPad, j = (i / kNodesPad) % g.H, s = i / (kNodesPad * g.H);
    int b = b0 + s;
    float v = 0.f;
    if (src != nullptr && n < kNodes && b < g.B) v = src[((size_t)b * kNodes + n) * g.H + j];
    dst[i] = v;
  }
}
__device__ void store_state(float* dst, const float* src, int b0, const LayerGeom& g) {
  if (dst == nullptr) return;
  for (int i = threadIdx.x; i < g.BT * kNodes * g.H; i += blockDim.x) {
    int j = i % g.H, n = (i / g.H) % kNodes, s = i / (g.H * kNodes);
    int b = b0 + s;
    if (b < g.B) dst[((size_t)b * kNodes + n) * g.H + j] = src[((size_t)s * g.H + j) * kNodesPad + n];
  }
}
// x[b, t, n, k] (global) -> xbuf[seq][k][16]; row 15 zero
__device__ void load_x(float* xbuf, int b0, int t, const LayerGeom& g) {
  const int per_seq = kNodesPad * g.F;
  for (int i = threadIdx.x; i < g.BT * per_seq; i += blockDim.x) {
    int k = i % g.F, n = (i / g.F) % kNodesPad, s = i / per_seq;
    int b = b0 + s;
    float v = 0.f;
    if (n < kNodes && b < g.B) v = __ldg(g.x + (size_t)b * g.sxb + (size_t)t * g.sxt + (size_t)n * g.F + k);
    xbuf[((size_t)s * g.F + k) * kNodesPad + n] = v;
  }
}

// ------------------------------------------------------------------------------------------
// LSTM family time loop: one CTA = BT sequences of one direction, all T steps
// ------------------------------------------------------------------------------------------
template <bool ATT>
__global__ void __launch_bounds__(kThreads, 1)
lstm_layer_kernel(LstmPacked w0, LstmPacked w1, DirPtrs d0, DirPtrs d1, LayerGeom g) {
  extern __shared__ __align__(16) float smem[];
  const LstmPacked w = blockIdx.y == 0 ? w0 : w1;
  const DirPtrs d = blockIdx.y == 0 ? d0 : d1;
  const int H = g.H, F = g.F, BT = g.BT;
  const int KX = F > H ? F : H;
  float* hbuf = smem;                                  // [2][BT][H][16]
  float* cbuf = hbuf + (size_t)2 * BT * H * kNodesPad; // [BT][H][16]
  float* xbuf = cbuf + (size_t)BT * H * kNodesPad;     // [BT][KX][16]  (x_t; reused as attention scratch)
  float* qbuf = xbuf + (size_t)BT * KX * kNodesPad;    // [BT][H]
  float* sbuf = qbuf + (size_t)BT * H;                 // [BT][H]
  float* abuf = sbuf + (size_t)BT * H;                 // [BT][16]
  float* Pbuf = abuf + (size_t)BT * kNodesPad;         // [4][16][16]
  const int b0 = blockIdx.x * BT;
  const int y_off = blockIdx.y * H;

  load_state(hbuf, d.h0, b0, g);
  load_state(cbuf, d.c0, b0, g);
  for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x) Pbuf[i] = w.P[i];
  int cur = 0;
  const int ntask = BT * H;

  for (int step = 0; step < g.T; ++step) {
    const int t = d.reverse ? g.T - 1 - step : step;
    float* hcur = hbuf + (size_t)cur * BT * H * kNodesPad;
    float* hnxt = hbuf + (size_t)(cur ^ 1) * BT * H * kNodesPad;
    __syncthreads();                 // previous step done with xbuf / hbuf
    load_x(xbuf, b0, t, g);
    __syncthreads();

    for (int task = threadIdx.x; task < ntask; task += blockDim.x) {
      const int s = task / H, j = task % H;
      float acc[4][16];
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int n = 0; n < 16; ++n) acc[q][n] = 0.f;
      accum4(acc, xbuf + (size_t)s * F * kNodesPad, w.Wg4 + j, F, H);
      accum4(acc, hcur + (size_t)s * H * kNodesPad, w.Wg4 + (size_t)F * H + j, H, H);
      const float4 bias = w.bias4[j];
      float tmp[16], cs[16];
      load16(cs, cbuf + ((size_t)s * H + j) * kNodesPad);
      mix15(acc[0], Pbuf, tmp);                        // i
#pragma unroll
      for (int m = 0; m < kNodes; ++m) acc[0][m] = sigmoidf_(tmp[m] + bias.x);
      mix15(acc[2], Pbuf + 512, tmp);                  // c (candidate)
#pragma unroll
      for (int m = 0; m < kNodes; ++m) acc[0][m] *= tanhf_(tmp[m] + bias.z);
      mix15(acc[1], Pbuf + 256, tmp);                  // f
#pragma unroll
      for (int m = 0; m < kNodes; ++m) cs[m] = fmaf(sigmoidf_(tmp[m] + bias.y), cs[m], acc[0][m]);
      cs[15] = 0.f;
      mix15(acc[3], Pbuf + 768, tmp);                  // o
      float hy[16];
#pragma unroll
      for (int m = 0; m < kNodes; ++m) hy[m] = sigmoidf_(tmp[m] + bias.w) * tanhf_(cs[m]);
      hy[15] = 0.f;
      store16(cbuf + ((size_t)s * H + j) * kNodesPad, cs);
      store16(hnxt + ((size_t)s * H + j) * kNodesPad, hy);
      if (!ATT) {
        const int b = b0 + s;
        if (b < g.B) {
          float* yp = g.y + (size_t)b * g.syb + (size_t)t * g.syt + y_off + j;
#pragma unroll
          for (int m = 0; m < kNodes; ++m) yp[(size_t)m * g.yld] = apply_act(hy[m], g.out_act);
        }
      }
    }

    if (ATT) {
      __syncthreads();
      // s[seq][k] = sum over nodes of hy  (q_t = relu((sum_n hy_n) W_a^T), net_aagc.py:200)
      for (int i = threadIdx.x; i < ntask; i += blockDim.x) {
        float v[16];
        load16(v, hnxt + (size_t)i * kNodesPad);
        float sum = 0.f;
#pragma unroll
        for (int n = 0; n < kNodes; ++n) sum += v[n];
        sbuf[i] = sum;
      }
      __syncthreads();
      for (int task = threadIdx.x; task < ntask; task += blockDim.x) {
        const int s = task / H, j = task % H;
        float q = 0.f;
        const float* sv = sbuf + (size_t)s * H;
        for (int k = 0; k < H; ++k) q = fmaf(sv[k], __ldg(w.Wa_t + (size_t)k * H + j), q);
        qbuf[task] = fmaxf(q, 0.f);
      }
      __syncthreads();
      for (int task = threadIdx.x; task < ntask; task += blockDim.x) {
        const int s = task / H, j = task % H;
        float e[16];
#pragma unroll
        for (int n = 0; n < 16; ++n) e[n] = 0.f;
        accum1(e, hnxt + (size_t)s * H * kNodesPad, w.Wh_t + j, H, H);
        float wq = w.bs[j];
        const float* qv = qbuf + (size_t)s * H;
        for (int k = 0; k < H; ++k) wq = fmaf(qv[k], __ldg(w.Wq_t + (size_t)k * H + j), wq);
        const float uj = w.u[j];
#pragma unroll
        for (int n = 0; n < kNodes; ++n) e[n] = tanhf_(e[n] + wq) * uj;
        e[15] = 0.f;
        store16(xbuf + ((size_t)s * KX + j) * kNodesPad, e);     // scratch [seq][j][16]
      }
      __syncthreads();
      for (int i = threadIdx.x; i < BT * kNodesPad; i += blockDim.x) {
        const int s = i / kNodesPad, n = i % kNodesPad;
        float a = 0.f;
        if (n < kNodes) {
          const float* ev = xbuf + (size_t)s * KX * kNodesPad + n;
          for (int j = 0; j < H; ++j) a += ev[(size_t)j * kNodesPad];
          a = 1.0f + sigmoidf_(a + w.bu[n]);                     // hy + hy * a_t  (net_aagc.py:212-213)
        }
        abuf[i] = a;
      }
      __syncthreads();
      for (int task = threadIdx.x; task < ntask; task += blockDim.x) {
        const int s = task / H, j = task % H;
        float hy[16], av[16];
        load16(hy, hnxt + (size_t)task * kNodesPad);
        load16(av, abuf + (size_t)s * kNodesPad);
#pragma unroll
        for (int m = 0; m < 16; ++m) hy[m] *= av[m];
        store16(hnxt + (size_t)task * kNodesPad, hy);
        const int b = b0 + s;
        if (b < g.B) {
          float* yp = g.y + (size_t)b * g.syb + (size_t)t * g.syt + y_off + j;
#pragma unroll
          for (int m = 0; m < kNodes; ++m) yp[(size_t)m * g.yld] = apply_act(hy[m], g.out_act);
        }
      }
    }
    cur ^= 1;
  }
  __syncthreads();
  store_state(d.hT, hbuf + (size_t)cur * BT * H * kNodesPad, b0, g);
  store_state(d.cT, cbuf, b0, g);
}

// ------------------------------------------------------------------------------------------
// Training path of the LSTM family (BPTT): forward with a tape, and the reverse-time backward chain.
//
// Tape layout (all fp32, allocated by the caller): one record per (direction d, time t, sequence b),
// idx = (d*T + t)*B + b, unit-major with the node index fastest and padded to 16 -- the shared-memory
// layout of the kernels, so every record is written / read as contiguous 64-byte rows; node slot 15 is
// always 0 so that reductions over the padded axis are exact.
//   gates [idx][4][H][16]  activated i, f, c~, o          (overwritten in place with dz by the backward)
//   u     [idx][4][H][16]  pre-mix accumulators S W_g^T   (for the adjacency gradients)
//   c, hh, e [idx][H][16]   c'_t, hy_t (before attention), e_t = tanh(..)
//   hp    [D][B][T][15][H]  h'_t, node-major like the layer input x, so that [x | h_prev] and the mixed gate
//                           gradients dzm [D][B][T][15][4H] are the row-major operands of the hoisted GEMMs
//   a [idx][16] (sigmoid output), q [idx][H], s [idx][H] (node sums)
// ------------------------------------------------------------------------------------------
struct TrainGeom {
  LayerGeom g;
  a3gc_tape tape;
  const float* hmask;   // [D][B][T][15][H] multiplicative recurrent-dropout mask (0 or 1/(1-p)); nullptr = none
};

template <bool ATT>
__global__ void __launch_bounds__(kThreads, 1)
lstm_train_fwd_kernel(LstmPacked w0, LstmPacked w1, DirPtrs d0, DirPtrs d1, TrainGeom tg) {
  extern __shared__ __align__(16) float smem[];
  const LayerGeom& g = tg.g;
  const a3gc_tape& tp = tg.tape;
  const LstmPacked w = blockIdx.y == 0 ? w0 : w1;
  const DirPtrs d = blockIdx.y == 0 ? d0 : d1;
  const int H = g.H, F = g.F, BT = g.BT;
  const int KX = F > H ? F : H;
  float* hbuf = smem;                                  // [2][BT][H][16]
  float* cbuf = hbuf + (size_t)2 * BT * H * kNodesPad; // [BT][H][16]
  float* xbuf = cbuf + (size_t)BT * H * kNodesPad;     // [BT][KX][16]
  float* qbuf = xbuf + (size_t)BT * KX * kNodesPad;    // [BT][H]
  float* sbuf = qbuf + (size_t)BT * H;                 // [BT][H]
  float* abuf = sbuf + (size_t)BT * H;                 // [BT][16]
  float* Pbuf = abuf + (size_t)BT * kNodesPad;         // [4][16][16]
  const int b0 = blockIdx.x * BT;
  const int y_off = blockIdx.y * H;
  const size_t HN = (size_t)H * kNodesPad;

  load_state(hbuf, d.h0, b0, g);
  load_state(cbuf, d.c0, b0, g);
  for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x) Pbuf[i] = w.P[i];
  int cur = 0;
  const int ntask = BT * H;

  for (int step = 0; step < g.T; ++step) {
    const int t = d.reverse ? g.T - 1 - step : step;
    const size_t rec0 = ((size_t)blockIdx.y * g.T + t) * g.B;      // record index of sequence 0 at (d, t)
    // node-major arrays [D][B][T][15][H]: offset of (d, b = 0, t, n = 0, j = 0); add ((b*T*15) + n) * H + j
    const size_t nm0 = ((size_t)blockIdx.y * g.B * g.T + t) * kNodes * H;
    float* hcur = hbuf + (size_t)cur * BT * HN;
    float* hnxt = hbuf + (size_t)(cur ^ 1) * BT * HN;
    __syncthreads();
    load_x(xbuf, b0, t, g);
    if (tg.hmask != nullptr) {   // recurrent dropout acts on the h that enters the gates only (net_aagc.py:181)
      for (int i = threadIdx.x; i < BT * kNodes * H; i += blockDim.x) {
        const int j = i % H, n = (i / H) % kNodes, s = i / (H * kNodes), b = b0 + s;
        if (b < g.B) hcur[((size_t)s * H + j) * kNodesPad + n] *= tg.hmask[nm0 + ((size_t)b * g.T * kNodes + n) * H + j];
      }
    }
    __syncthreads();

    for (int task = threadIdx.x; task < ntask; task += blockDim.x) {
      const int s = task / H, j = task % H;
      const int b = b0 + s;
      const bool live = b < g.B;
      const size_t rec = rec0 + b;
      float acc[4][16];
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int n = 0; n < 16; ++n) acc[q][n] = 0.f;
      accum4(acc, xbuf + (size_t)s * F * kNodesPad, w.Wg4 + j, F, H);
      accum4(acc, hcur + (size_t)s * HN, w.Wg4 + (size_t)F * H + j, H, H);
      const float4 bias = w.bias4[j];
      const float bb[4] = {bias.x, bias.y, bias.z, bias.w};
      float gate[4][16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        acc[q][15] = 0.f;
        if (live && tp.u != nullptr) store16(tp.u + ((rec * 4 + q) * H + j) * kNodesPad, acc[q]);
        float tmp[16];
        mix15(acc[q], Pbuf + q * 256, tmp);
#pragma unroll
        for (int m = 0; m < kNodes; ++m) gate[q][m] = q == 2 ? tanhf_(tmp[m] + bb[q]) : sigmoidf_(tmp[m] + bb[q]);
        gate[q][15] = 0.f;
        if (live) store16(tp.gates + ((rec * 4 + q) * H + j) * kNodesPad, gate[q]);
      }
      float cs[16], hy[16];
      load16(cs, cbuf + ((size_t)s * H + j) * kNodesPad);
#pragma unroll
      for (int m = 0; m < kNodes; ++m) {
        cs[m] = fmaf(gate[1][m], cs[m], gate[0][m] * gate[2][m]);
        hy[m] = gate[3][m] * tanhf_(cs[m]);
      }
      cs[15] = 0.f; hy[15] = 0.f;
      store16(cbuf + ((size_t)s * H + j) * kNodesPad, cs);
      store16(hnxt + ((size_t)s * H + j) * kNodesPad, hy);
      if (live) {
        store16(tp.c + (rec * H + j) * kNodesPad, cs);
        store16(tp.hh + (rec * H + j) * kNodesPad, hy);
        if (!ATT) {
          float* hpp = tp.hp + nm0 + (size_t)b * g.T * kNodes * H + j;
          float* yp = g.y + (size_t)b * g.syb + (size_t)t * g.syt + y_off + j;
#pragma unroll
          for (int m = 0; m < kNodes; ++m) { hpp[(size_t)m * H] = hy[m]; yp[(size_t)m * g.yld] = apply_act(hy[m], g.out_act); }
        }
      }
    }

    if (ATT) {
      __syncthreads();
      for (int i = threadIdx.x; i < ntask; i += blockDim.x) {
        float v[16];
        load16(v, hnxt + (size_t)i * kNodesPad);
        float sum = 0.f;
#pragma unroll
        for (int n = 0; n < kNodes; ++n) sum += v[n];
        sbuf[i] = sum;
        const int b = b0 + i / H;
        if (b < g.B) tp.s[(rec0 + b) * H + i % H] = sum;
      }
      __syncthreads();
      for (int task = threadIdx.x; task < ntask; task += blockDim.x) {
        const int s = task / H, j = task % H;
        float q = 0.f;
        const float* sv = sbuf + (size_t)s * H;
        for (int k = 0; k < H; ++k) q = fmaf(sv[k], __ldg(w.Wa_t + (size_t)k * H + j), q);
        q = fmaxf(q, 0.f);
        qbuf[task] = q;
        if (b0 + s < g.B) tp.q[(rec0 + b0 + s) * H + j] = q;
      }
      __syncthreads();
      for (int task = threadIdx.x; task < ntask; task += blockDim.x) {
        const int s = task / H, j = task % H;
        float e[16];
#pragma unroll
        for (int n = 0; n < 16; ++n) e[n] = 0.f;
        accum1(e, hnxt + (size_t)s * HN, w.Wh_t + j, H, H);
        float wq = w.bs[j];
        const float* qv = qbuf + (size_t)s * H;
        for (int k = 0; k < H; ++k) wq = fmaf(qv[k], __ldg(w.Wq_t + (size_t)k * H + j), wq);
        const float uj = w.u[j];
#pragma unroll
        for (int n = 0; n < kNodes; ++n) e[n] = tanhf_(e[n] + wq);
        e[15] = 0.f;
        if (b0 + s < g.B) store16(tp.e + ((rec0 + b0 + s) * H + j) * kNodesPad, e);
#pragma unroll
        for (int n = 0; n < kNodes; ++n) e[n] *= uj;
        store16(xbuf + ((size_t)s * KX + j) * kNodesPad, e);     // scratch [seq][j][16]
      }
      __syncthreads();
      for (int i = threadIdx.x; i < BT * kNodesPad; i += blockDim.x) {
        const int s = i / kNodesPad, n = i % kNodesPad;
        float a = 0.f, sg = 0.f;
        if (n < kNodes) {
          const float* ev = xbuf + (size_t)s * KX * kNodesPad + n;
          for (int j = 0; j < H; ++j) a += ev[(size_t)j * kNodesPad];
          sg = sigmoidf_(a + w.bu[n]);
          a = 1.0f + sg;                                           // hy + hy * a_t  (net_aagc.py:212-213)
        }
        abuf[i] = a;
        if (b0 + s < g.B) tp.a[(rec0 + b0 + s) * kNodesPad + n] = sg;
      }
      __syncthreads();
      for (int task = threadIdx.x; task < ntask; task += blockDim.x) {
        const int s = task / H, j = task % H;
        float hy[16], av[16];
        load16(hy, hnxt + (size_t)task * kNodesPad);
        load16(av, abuf + (size_t)s * kNodesPad);
#pragma unroll
        for (int m = 0; m < 16; ++m) hy[m] *= av[m];
        store16(hnxt + (size_t)task * kNodesPad, hy);
        const int b = b0 + s;
        if (b < g.B) {
          float* hpp = tp.hp + nm0 + (size_t)b * g.T * kNodes * H + j;
          float* yp = g.y + (size_t)b * g.syb + (size_t)t * g.syt + y_off + j;
#pragma unroll
          for (int m = 0; m < kNodes; ++m) { hpp[(size_t)m * H] = hy[m]; yp[(size_t)m * g.yld] = apply_act(hy[m], g.out_act); }
        }
      }
    }
    cur ^= 1;
  }
  __syncthreads();
  store_state(d.hT, hbuf + (size_t)cur * BT * HN, b0, g);
  store_state(d.cT, cbuf, b0, g);
}

// Reverse-time chain of one direction.  Consumes dY (gradient of the layer output act(h'_t)), the incoming
// gradients of the final state, and the tape; produces per step the gate-pre-activation gradients dz (in place
// of tape.gates), their mixed form dzm = P_g^T dz_g (the operand of the hoisted weight / input gradient GEMMs),
// the attention gradients dep / dqs / dqp / dap, and finally the gradients of the initial state.
struct BwdDir {
  const float* Wg[4];      // gcn_kernel_g [H][F+H]  (reference layout)
  const float* Wh;         // attention_wh [H][H]
  const float* Wq;         // attention_wq [H][H]
  const float* Wa;         // attention_w  [H][H]
  const float* u;          // attention_u  [H]
  const float* PT;         // [4][16][16]  PT_g[n][m] = P_g[m][n]  (zero padded)
  const uint4* wfrag;      // blocked kernel, modes 3 / 4: the seven H x H weight blocks in mma.sync A-fragment order (pack_bwd_frag_kernel)
  const float* c0;         // [B][15][H] initial cell state (nullptr = zeros)
  const float* dhT; const float* dcT;   // [B][15][H] gradients of the final state (nullptr = zeros)
  float* dh0; float* dc0;               // [B][15][H] gradients of the initial state (nullptr = skip)
  int reverse;
};
// One element of grads.dzm: plain fp32, or -- when the caller passed the bf16 arrays (the "mixed" form of the hoisted GEMMs,
// include/a3gc_b200.h) -- the TF32-exact head in dzm plus bf16(head) and bf16(remainder), which saves the separate split
// pass over the largest operand of the training step.  (Packing the bf16 halves of a unit pair into one 32-bit store through a
// lane shuffle was measured: phase F 38 k -> 56 k cycles per step, not kept.)
__device__ __forceinline__ void emit_dzm(const a3gc_tape_grads& gr, size_t idx, float v) {
  if (gr.dzm_hi16 == nullptr) { gr.dzm[idx] = v; return; }
  const float h = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u);
  gr.dzm[idx] = h;
  gr.dzm_hi16[idx] = __bfloat16_as_ushort(__float2bfloat16_rn(h));
  gr.dzm_lo16[idx] = __bfloat16_as_ushort(__float2bfloat16_rn(v - h));
}

struct BwdGeom {
  const float* dy; int64_t syb, syt, yld;   // dY element (b, t, n, d*H + j)
  a3gc_tape tape;
  a3gc_tape_grads gr;
  const float* hmask;
  int B, T, F, H, out_act, BT;
  long long* trace;                  // diagnostics: per-phase cycle sums of CTA (0, 0), or nullptr
  int tcore;                         // blocked kernel: how the weight products run (A3GC_BWD_MMA: 0 FFMA, 1..4 mma.sync forms; default 4)
};

template <bool ATT>
__global__ void __launch_bounds__(kThreads, 1)
lstm_train_bwd_kernel(BwdDir d0, BwdDir d1, BwdGeom g) {
  extern __shared__ __align__(16) float smem[];
  const BwdDir d = blockIdx.y == 0 ? d0 : d1;
  const a3gc_tape& tp = g.tape;
  const a3gc_tape_grads& gr = g.gr;
  const int H = g.H, F = g.F, BT = g.BT, K = F + H;
  const size_t HN = (size_t)H * kNodesPad;
  float* dh = smem;                       // [BT][H][16]  gradient wrt h'_t carried from the later step
  float* dc = dh + BT * HN;               // [BT][H][16]
  float* dhh = dc + BT * HN;              // [BT][H][16]  gradient wrt hy_t
  float* dep = dhh + BT * HN;             // [BT][H][16]  (ATT) gradient wrt the tanh pre-activation; also scratch
  float* dzm = dep + BT * HN;             // [4][BT][H][16]
  float* v1 = dzm + 4 * BT * HN;          // [BT][H]   dqs, then ds
  float* v2 = v1 + BT * H;                // [BT][H]   dqp
  float* abuf = v2 + BT * H;              // [BT][16]  dap
  float* al = abuf + BT * kNodesPad;      // [BT][16]  alpha
  float* PT = al + BT * kNodesPad;        // [4][16][16]
  const int b0 = blockIdx.x * BT;
  const int y_off = blockIdx.y * H;
  const int ntask = BT * H;
  LayerGeom lg; lg.B = g.B; lg.H = H; lg.BT = BT;
  load_state(dh, d.dhT, b0, lg);
  load_state(dc, d.dcT, b0, lg);
  for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x) PT[i] = d.PT[i];

  for (int step = g.T - 1; step >= 0; --step) {
    const int t = d.reverse ? g.T - 1 - step : step;                 // time of the forward step being undone
    const int tprev = d.reverse ? t + 1 : t - 1;                       // time whose c' was this step's c input
    const size_t rec0 = ((size_t)blockIdx.y * g.T + t) * g.B;
    const size_t recp0 = ((size_t)blockIdx.y * g.T + tprev) * g.B;
    const size_t nm0 = ((size_t)blockIdx.y * g.B * g.T + t) * kNodes;   // node-major row of (d, b = 0, t, n = 0); add b*T*15 + n
    __syncthreads();
    // ---- A: dh' += dY (1 - y^2);  dhy = dh' (1 + alpha);  partial dalpha[n] = sum_j dh' hy
    if (ATT) {
      for (int i = threadIdx.x; i < BT * kNodesPad; i += blockDim.x) {
        const int b = b0 + i / kNodesPad;
        al[i] = b < g.B ? tp.a[(rec0 + b) * kNodesPad + i % kNodesPad] : 0.f;
      }
      __syncthreads();
    }
    for (int task = threadIdx.x; task < ntask; task += blockDim.x) {
      const int s = task / H, j = task % H, b = b0 + s;
      float dv[16];
      load16(dv, dh + (size_t)task * kNodesPad);
      if (b < g.B) {
        const float* hpp = tp.hp + (nm0 + (size_t)b * g.T * kNodes) * H + j;
        const float* yp = g.dy + (size_t)b * g.syb + (size_t)t * g.syt + y_off + j;
#pragma unroll
        for (int n = 0; n < kNodes; ++n) {
          const float gy = __ldg(yp + (size_t)n * g.yld);
          if (g.out_act == A3GC_ACT_TANH) { const float y = tanhf(hpp[(size_t)n * H]); dv[n] = fmaf(gy, 1.0f - y * y, dv[n]); }
          else dv[n] += gy;
        }
        dv[15] = 0.f;
        if (ATT) {
          float hh[16], pa[16], av[16];
          load16(hh, tp.hh + ((rec0 + b) * H + j) * kNodesPad);
          load16(av, al + (size_t)s * kNodesPad);
#pragma unroll
          for (int n = 0; n < 16; ++n) { pa[n] = dv[n] * hh[n]; dv[n] *= 1.0f + av[n]; }
          store16(dep + (size_t)task * kNodesPad, pa);          // scratch: partial dalpha
        }
      } else {
#pragma unroll
        for (int n = 0; n < 16; ++n) dv[n] = 0.f;
        if (ATT) store16(dep + (size_t)task * kNodesPad, dv);
      }
      store16(dhh + (size_t)task * kNodesPad, dv);
    }
    if (ATT) {
      __syncthreads();
      // ---- B: dalpha[n] = sum_j partial;  dap = dalpha * a (1 - a)
      for (int i = threadIdx.x; i < BT * kNodesPad; i += blockDim.x) {
        const int s = i / kNodesPad, n = i % kNodesPad;
        float a = 0.f;
        const float* pv = dep + (size_t)s * HN + n;
        for (int j = 0; j < H; ++j) a += pv[(size_t)j * kNodesPad];
        const float sg = al[i];
        a *= sg * (1.0f - sg);
        abuf[i] = a;
        if (b0 + s < g.B) gr.dap[(rec0 + b0 + s) * kNodesPad + n] = a;
      }
      __syncthreads();
      // ---- C: dep[n][j] = dap[n] u_j (1 - e^2);  dqs_j = sum_n dep
      for (int task = threadIdx.x; task < ntask; task += blockDim.x) {
        const int s = task / H, j = task % H, b = b0 + s;
        float e[16], av[16], o[16];
        float sum = 0.f;
        if (b < g.B) {
          load16(e, tp.e + ((rec0 + b) * H + j) * kNodesPad);
          load16(av, abuf + (size_t)s * kNodesPad);
          const float uj = d.u[j];
#pragma unroll
          for (int n = 0; n < 16; ++n) { o[n] = av[n] * uj * (1.0f - e[n] * e[n]); sum += o[n]; }
          store16(gr.dep + ((rec0 + b) * H + j) * kNodesPad, o);
          gr.dqs[(rec0 + b) * H + j] = sum;
        } else {
#pragma unroll
          for (int n = 0; n < 16; ++n) o[n] = 0.f;
        }
        store16(dep + (size_t)task * kNodesPad, o);
        v1[task] = sum;
      }
      __syncthreads();
      // ---- D: dq_j = sum_k dqs_k Wq[k][j];  dqp = dq [q > 0]
      for (int task = threadIdx.x; task < ntask; task += blockDim.x) {
        const int s = task / H, j = task % H, b = b0 + s;
        float dq = 0.f;
        const float* sv = v1 + (size_t)s * H;
        for (int k = 0; k < H; ++k) dq = fmaf(sv[k], __ldg(d.Wq + (size_t)k * H + j), dq);
        float r = 0.f;
        if (b < g.B) {
          r = tp.q[(rec0 + b) * H + j] > 0.f ? dq : 0.f;
          gr.dqp[(rec0 + b) * H + j] = r;
        }
        v2[task] = r;
      }
      __syncthreads();
      // ---- E: ds_j = sum_k dqp_k Wa[k][j];  dhy[n][j] += ds_j + sum_k dep[n][k] Wh[k][j]
      for (int task = threadIdx.x; task < ntask; task += blockDim.x) {
        const int s = task / H, j = task % H;
        float ds = 0.f;
        const float* sv = v2 + (size_t)s * H;
        for (int k = 0; k < H; ++k) ds = fmaf(sv[k], __ldg(d.Wa + (size_t)k * H + j), ds);
        float acc[16];
        load16(acc, dhh + (size_t)task * kNodesPad);
        accum1(acc, dep + (size_t)s * HN, d.Wh + j, H, H);
#pragma unroll
        for (int n = 0; n < kNodes; ++n) acc[n] += ds;
        acc[15] = 0.f;
        store16(dhh + (size_t)task * kNodesPad, acc);
      }
    }
    __syncthreads();
    // ---- F: LSTM pointwise backward, dz (global, in place of the gates) and dzm = P_g^T dz_g
    for (int task = threadIdx.x; task < ntask; task += blockDim.x) {
      const int s = task / H, j = task % H, b = b0 + s;
      float dz[4][16];
      float dcv[16];
      if (b < g.B) {
        float gi[16], gf[16], gg[16], go[16], cc[16], cp[16], dv[16];
        float* gp = tp.gates + ((rec0 + b) * 4 * H + j) * kNodesPad;
        load16(gi, gp); load16(gf, gp + HN); load16(gg, gp + 2 * HN); load16(go, gp + 3 * HN);
        load16(cc, tp.c + ((rec0 + b) * H + j) * kNodesPad);
        if (step > 0) load16(cp, tp.c + ((recp0 + b) * H + j) * kNodesPad);
        else {
#pragma unroll
          for (int n = 0; n < 16; ++n) cp[n] = (d.c0 != nullptr && n < kNodes) ? d.c0[((size_t)b * kNodes + n) * H + j] : 0.f;
        }
        load16(dv, dhh + (size_t)task * kNodesPad);
        load16(dcv, dc + (size_t)task * kNodesPad);
#pragma unroll
        for (int n = 0; n < 16; ++n) {
          const float tc = tanhf(cc[n]);
          const float dcn = fmaf(dv[n] * go[n], 1.0f - tc * tc, dcv[n]);
          dz[3][n] = dv[n] * tc * go[n] * (1.0f - go[n]);
          dz[0][n] = dcn * gg[n] * gi[n] * (1.0f - gi[n]);
          dz[1][n] = dcn * cp[n] * gf[n] * (1.0f - gf[n]);
          dz[2][n] = dcn * gi[n] * (1.0f - gg[n] * gg[n]);
          dcv[n] = dcn * gf[n];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) { dz[q][15] = 0.f; store16(gp + q * HN, dz[q]); }
        dcv[15] = 0.f;
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int n = 0; n < 16; ++n) dz[q][n] = 0.f;
#pragma unroll
        for (int n = 0; n < 16; ++n) dcv[n] = 0.f;
      }
      store16(dc + (size_t)task * kNodesPad, dcv);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float m[16];
        mix15(dz[q], PT + q * 256, m);
        store16(dzm + ((size_t)q * BT * H + task) * kNodesPad, m);
        if (b < g.B) {   // [D][B][T][15][4H], column q*H + j: rows match x [B,T,15,F], so dW and dX are plain GEMMs
          const size_t zi = (nm0 + (size_t)b * g.T * kNodes) * 4 * H + (size_t)q * H + j;
#pragma unroll
          for (int n = 0; n < kNodes; ++n) emit_dzm(gr, zi + (size_t)n * 4 * H, m[n]);
        }
      }
    }
    __syncthreads();
    // ---- G: dh'_{prev}[n][k] = sum_g sum_j dzm_g[n][j] W_g[j][F + k]   (then the recurrent-dropout mask of this step)
    for (int task = threadIdx.x; task < ntask; task += blockDim.x) {
      const int s = task / H, k = task % H, b = b0 + s;
      float acc[16];
#pragma unroll
      for (int n = 0; n < 16; ++n) acc[n] = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) accum1(acc, dzm + ((size_t)q * BT + s) * HN, d.Wg[q] + F + k, H, K);
      acc[15] = 0.f;
      if (g.hmask != nullptr && b < g.B) {
        const float* mp = g.hmask + (nm0 + (size_t)b * g.T * kNodes) * H + k;
#pragma unroll
        for (int n = 0; n < kNodes; ++n) acc[n] *= mp[(size_t)n * H];
      }
      store16(dh + (size_t)task * kNodesPad, acc);
    }
  }
  __syncthreads();
  store_state(d.dh0, dh, b0, lg);
  store_state(d.dc0, dc, b0, lg);
}

// acc[si][ui][n] += sum_{k0 <= k < k1} act[si][k][n] * w[k * ldw + ui * ustride]   (act: shared [seq][k][16])
// Register blocking over NS sequences x NU units: one weight load feeds 16 * NS FMAs, one activation float4 4 * NU.
template <int NS, int NU>
__device__ __forceinline__ void accum_blk(float (&acc)[NS][NU][16], const float* __restrict__ act, size_t seq_stride,
                                          const float* __restrict__ w, int ldw, int ustride, int k0, int k1) {
  // weights of the next group of kG contraction steps are fetched (L2 latency) while the current group is consumed;
  // k1 - k0 is a multiple of kG for every shape the blocked kernel accepts
  constexpr int kG = 8;
  float wn[kG][NU];
#pragma unroll
  for (int i = 0; i < kG; ++i)
#pragma unroll
    for (int ui = 0; ui < NU; ++ui) wn[i][ui] = __ldg(w + (size_t)(k0 + i) * ldw + ui * ustride);
#pragma unroll 1
  for (int k = k0; k < k1; k += kG) {
    float wc[kG][NU];
#pragma unroll
    for (int i = 0; i < kG; ++i)
#pragma unroll
      for (int ui = 0; ui < NU; ++ui) wc[i][ui] = wn[i][ui];
    if (k + kG < k1) {
#pragma unroll
      for (int i = 0; i < kG; ++i)
#pragma unroll
        for (int ui = 0; ui < NU; ++ui) wn[i][ui] = __ldg(w + (size_t)(k + kG + i) * ldw + ui * ustride);
    }
#pragma unroll
    for (int i = 0; i < kG; ++i) {
#pragma unroll
      for (int si = 0; si < NS; ++si) {
        const float4* a4 = reinterpret_cast<const float4*>(act + si * seq_stride + (size_t)(k + i) * kNodesPad);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 v = a4[q];
#pragma unroll
          for (int ui = 0; ui < NU; ++ui) {
            acc[si][ui][4 * q] = fmaf(wc[i][ui], v.x, acc[si][ui][4 * q]);
            acc[si][ui][4 * q + 1] = fmaf(wc[i][ui], v.y, acc[si][ui][4 * q + 1]);
            acc[si][ui][4 * q + 2] = fmaf(wc[i][ui], v.z, acc[si][ui][4 * q + 2]);
            acc[si][ui][4 * q + 3] = fmaf(wc[i][ui], v.w, acc[si][ui][4 * q + 3]);
          }
        }
      }
    }
  }
}

// Partial products of a batched GEMV out[s][j] = sum_k vec[s][k] W[k][j] for all BT sequences of the CTA: this thread
// covers units j4..j4+3 (one 16-byte load per k) and contraction steps [kd*kq, (kd+1)*kq); G loads are in flight
// per thread (G of 16 bytes) because the loop is bound by L2 latency, not bandwidth.  part -> scr[(kd*BT + s)*H + j]
template <int G>
__device__ __forceinline__ void gemv_part(const float* __restrict__ vec, const float* __restrict__ W, float* __restrict__ scr,
                                          int H, int BT, int j4, int kd, int kq) {
  float part[8][4];
#pragma unroll
  for (int s = 0; s < 8; ++s)
#pragma unroll
    for (int i = 0; i < 4; ++i) part[s][i] = 0.f;
  const float4* wp = reinterpret_cast<const float4*>(W + (size_t)kd * kq * H + j4);
  const float* vp = vec + kd * kq;
  const int ld4 = H / 4;
#pragma unroll 1
  for (int k = 0; k < kq; k += G) {
    float4 wv[G];
#pragma unroll
    for (int i = 0; i < G; ++i) wv[i] = __ldg(wp + (size_t)(k + i) * ld4);
#pragma unroll
    for (int i = 0; i < G; ++i)
#pragma unroll
      for (int s = 0; s < 8; ++s)
        if (s < BT) {
          const float v = vp[s * H + k + i];
          part[s][0] = fmaf(v, wv[i].x, part[s][0]);
          part[s][1] = fmaf(v, wv[i].y, part[s][1]);
          part[s][2] = fmaf(v, wv[i].z, part[s][2]);
          part[s][3] = fmaf(v, wv[i].w, part[s][3]);
        }
  }
#pragma unroll
  for (int s = 0; s < 8; ++s)
    if (s < BT) *reinterpret_cast<float4*>(scr + ((size_t)kd * BT + s) * H + j4) = make_float4(part[s][0], part[s][1], part[s][2], part[s][3]);
}

// ---- 3xTF32 tensor-core contraction for the two weight products of the backward chain ---------------------------------
// D[m][n] += sum_k W[k][m] act[k][n] with m = output unit, n = (sequence, node) row and k the contraction index: the
// units are the M side of mma.sync.m16n8k8 (row-major A = W^T, read straight from the reference layout in L2: a lane's four
// A registers are 32-byte segments of four weight rows), the BT * 16 node rows the N side (B fragments from the fp32
// shared-memory arrays [seq][k][16]).  Both operands are split into tf32 hi + tf32 lo in registers and three products are
// accumulated (lo x hi, hi x lo, hi x hi; the dropped lo x lo term is 2^-22 relative), the precision the hoisted gradient
// GEMMs of the training step use as well.  One warp owns MT x NT tiles for the whole contraction: no partial sums meet.
// hi = x rounded to TF32 (nearest, ties away: add half an ulp of the 10-bit mantissa to the magnitude and clear the 13 low
// bits -- two integer instructions; cvt.rna.tf32.f32 expands to a ten-instruction sequence on sm_100 and made the contraction
// loops ALU-bound), lo = x - hi exactly; the tensor core ignores the 13 low mantissa bits of lo (2^-11 of lo = 2^-22 of x).
__device__ __forceinline__ void tf32_split(float x, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(uint32_t lo_bits, uint32_t hi_bits) {      // {bf16(lo) in bits 0..15, bf16(hi) in 16..31}
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(__uint_as_float(hi_bits)), "f"(__uint_as_float(lo_bits)));
  return r;
}

// act: shared [seq][k][16] (seq_stride floats per sequence); n tiles nt0 .. nt0 + NT - 1 (tile = 8 rows: sequence tile / 2,
// nodes 8 (tile % 2) ..);  w: element (k, m) at w[k * ldw + m], m tiles m0 .. m0 + MT - 1;  kcount % 32 == 0
// MIXED: the two correction products (lo x hi, hi x lo) of a PAIR of k tiles run as one bf16 m16n8k16 product each (their
// operands rounded to bf16: 2^-9 on a term that is 2^-11 of the product), the hi x hi product stays TF32 -- 4 tensor
// instructions per 16 contraction steps instead of 6.  The k16 slots {2c, 2c+1, 2c+8, 2c+9} of lane c carry the contraction
// indices {c, c+4, c+8, c+12} it already holds for the two TF32 tiles, on the A and the B side alike.
// accc: accumulators of the correction products (MIXED); the callers pass acc itself (a separate set was measured: no faster).
template <int MT, int NT, bool MIXED>
__device__ __forceinline__ void contract_tf32(float (&acc)[MT][NT][4], float (&accc)[MT][NT][4], const float* __restrict__ act,
                                              size_t seq_stride, int nt0, const float* __restrict__ w, int ldw, int m0, int kcount) {
  constexpr int kD = 4;                                   // k tiles of weights in flight per warp (L2 latency)
  const int lane = threadIdx.x & 31, r = lane >> 2, c = lane & 3;
  const float* wp = w + (size_t)c * ldw + m0 * 16 + r;    // a0: (k = c, m = r)  a1: m + 8  a2: k + 4  a3: both
  const size_t k4 = (size_t)4 * ldw, k8 = (size_t)8 * ldw;
  float wn[kD][MT][4];
#pragma unroll
  for (int i = 0; i < kD; ++i)
#pragma unroll
    for (int mi = 0; mi < MT; ++mi) {
      const float* q = wp + i * k8 + mi * 16;
      wn[i][mi][0] = __ldg(q); wn[i][mi][1] = __ldg(q + 8); wn[i][mi][2] = __ldg(q + k4); wn[i][mi][3] = __ldg(q + k4 + 8);
    }
  const float* bp[NT];
#pragma unroll
  for (int ni = 0; ni < NT; ++ni) {
    const int nt = nt0 + ni;
    bp[ni] = act + (size_t)(nt >> 1) * seq_stride + (((nt & 1) ^ ((c >> 1) & 1)) * 8) + r + c * kNodesPad;   // writer: store16_swz
  }
  const int nkt = kcount / 8;
  auto refill = [&](int i, int kt) {
    if (kt < nkt) {
#pragma unroll
      for (int mi = 0; mi < MT; ++mi) {
        const float* q = wp + (size_t)kt * k8 + mi * 16;
        wn[i][mi][0] = __ldg(q); wn[i][mi][1] = __ldg(q + 8); wn[i][mi][2] = __ldg(q + k4); wn[i][mi][3] = __ldg(q + k4 + 8);
      }
    }
  };
#pragma unroll 1
  for (int kt0 = 0; kt0 < nkt; kt0 += kD) {
    if constexpr (!MIXED) {
#pragma unroll
      for (int i = 0; i < kD; ++i) {
        const int kt = kt0 + i;
        uint32_t ah[MT][4], al[MT][4];
#pragma unroll
        for (int mi = 0; mi < MT; ++mi)
#pragma unroll
          for (int e = 0; e < 4; ++e) tf32_split(wn[i][mi][e], ah[mi][e], al[mi][e]);
        refill(i, kt + kD);
        uint32_t bh[NT][2], bl[NT][2];
#pragma unroll
        for (int ni = 0; ni < NT; ++ni) {
          const float* q = bp[ni] + (size_t)kt * 8 * kNodesPad;
          tf32_split(q[0], bh[ni][0], bl[ni][0]);
          tf32_split(q[4 * kNodesPad], bh[ni][1], bl[ni][1]);
        }
#pragma unroll
        for (int mi = 0; mi < MT; ++mi)
#pragma unroll
          for (int ni = 0; ni < NT; ++ni) mma_tf32(acc[mi][ni], al[mi], bh[ni][0], bh[ni][1]);
#pragma unroll
        for (int mi = 0; mi < MT; ++mi)
#pragma unroll
          for (int ni = 0; ni < NT; ++ni) mma_tf32(acc[mi][ni], ah[mi], bl[ni][0], bl[ni][1]);
#pragma unroll
        for (int mi = 0; mi < MT; ++mi)
#pragma unroll
          for (int ni = 0; ni < NT; ++ni) mma_tf32(acc[mi][ni], ah[mi], bh[ni][0], bh[ni][1]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < kD; i += 2) {
        const int kt = kt0 + i;
        uint32_t ah[2][MT][4], ahb[MT][4], alb[MT][4];       // TF32 heads of the two tiles; bf16 pairs of heads / tails
#pragma unroll
        for (int mi = 0; mi < MT; ++mi) {
          uint32_t al[2][4];
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) tf32_split(wn[i + j][mi][e], ah[j][mi][e], al[j][e]);
          // k16 A registers: (row r, slots 2c, 2c+1) (row r+8, same) (row r, slots 2c+8, 2c+9) (row r+8, same)
          ahb[mi][0] = pack_bf16(ah[0][mi][0], ah[0][mi][2]); ahb[mi][1] = pack_bf16(ah[0][mi][1], ah[0][mi][3]);
          ahb[mi][2] = pack_bf16(ah[1][mi][0], ah[1][mi][2]); ahb[mi][3] = pack_bf16(ah[1][mi][1], ah[1][mi][3]);
          alb[mi][0] = pack_bf16(al[0][0], al[0][2]); alb[mi][1] = pack_bf16(al[0][1], al[0][3]);
          alb[mi][2] = pack_bf16(al[1][0], al[1][2]); alb[mi][3] = pack_bf16(al[1][1], al[1][3]);
        }
        refill(i, kt + kD);
        refill(i + 1, kt + 1 + kD);
        uint32_t bh[2][NT][2], bhb[NT][2], blb[NT][2];
#pragma unroll
        for (int ni = 0; ni < NT; ++ni) {
          uint32_t bl[2][2];
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const float* q = bp[ni] + (size_t)(kt + j) * 8 * kNodesPad;
            tf32_split(q[0], bh[j][ni][0], bl[j][0]);
            tf32_split(q[4 * kNodesPad], bh[j][ni][1], bl[j][1]);
          }
          bhb[ni][0] = pack_bf16(bh[0][ni][0], bh[0][ni][1]); bhb[ni][1] = pack_bf16(bh[1][ni][0], bh[1][ni][1]);
          blb[ni][0] = pack_bf16(bl[0][0], bl[0][1]); blb[ni][1] = pack_bf16(bl[1][0], bl[1][1]);
        }
#pragma unroll
        for (int mi = 0; mi < MT; ++mi)
#pragma unroll
          for (int ni = 0; ni < NT; ++ni) mma_bf16(accc[mi][ni], alb[mi], bhb[ni][0], bhb[ni][1]);
#pragma unroll
        for (int mi = 0; mi < MT; ++mi)
#pragma unroll
          for (int ni = 0; ni < NT; ++ni) mma_bf16(accc[mi][ni], ahb[mi], blb[ni][0], blb[ni][1]);
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int mi = 0; mi < MT; ++mi)
#pragma unroll
            for (int ni = 0; ni < NT; ++ni) mma_tf32(acc[mi][ni], ah[j][mi], bh[j][ni][0], bh[j][ni][1]);
      }
    }
  }
}

// The same contraction as contract_tf32<.., MIXED = true> with the weight side prepared once per launch: per (unit tile, 16
// contraction steps) four 16-byte words per lane -- the TF32 heads of the two k8 tiles, the bf16 pairs of the heads, the bf16
// pairs of the tails (pack_bwd_frag_kernel) -- so the loop neither splits nor packs weights and fetches them with four
// coalesced 128-bit loads instead of sixteen 32-bit ones.  Bit-identical to the in-register form (same split, same products).
// FULL = false: the fragments hold the fp32 weights themselves in fragment order (two words per lane, the bytes of the
// reference layout: at H = 256 the full form doubles the L2 traffic of every step and is slower); heads, tails and bf16 pairs
// are then formed in registers as in contract_tf32, but the fetch is two coalesced 128-bit loads instead of eight 32-bit
// loads that touch four lines each (the L1 tag stage was the busiest unit of the kernel).
template <int MT, int NT, bool FULL>
__device__ __forceinline__ void contract_packed(float (&acc)[MT][NT][4], float (&accc)[MT][NT][4], const float* __restrict__ act,
                                                size_t seq_stride, int nt0, const uint4* __restrict__ wpk, int nk16, int m0) {
  constexpr int kD = 2;                                   // 16-step groups of weights in flight per warp
  constexpr int kW = FULL ? 4 : 2;                        // 16-byte words per lane and (unit tile, 16 steps)
  const int lane = threadIdx.x & 31, r = lane >> 2, c = lane & 3;
  const uint4* wp = wpk + (size_t)m0 * nk16 * (kW * 32) + lane;   // tile (mi, k16), word w: wp[((mi * nk16 + k16) * kW + w) * 32]
  uint4 wn[kD][MT][kW];
#pragma unroll
  for (int i = 0; i < kD; ++i)
#pragma unroll
    for (int mi = 0; mi < MT; ++mi)
#pragma unroll
      for (int w = 0; w < kW; ++w) wn[i][mi][w] = __ldg(wp + ((size_t)(mi * nk16 + i) * kW + w) * 32);
  const float* bp[NT];
#pragma unroll
  for (int ni = 0; ni < NT; ++ni) {
    const int nt = nt0 + ni;
    bp[ni] = act + (size_t)(nt >> 1) * seq_stride + (((nt & 1) ^ ((c >> 1) & 1)) * 8) + r + c * kNodesPad;   // writer: store16_swz
  }
#pragma unroll 1
  for (int k0 = 0; k0 < nk16; k0 += kD) {
#pragma unroll
    for (int i = 0; i < kD; ++i) {
      const int k16 = k0 + i;
      uint32_t ah[2][MT][4], ahb[MT][4], alb[MT][4];
#pragma unroll
      for (int mi = 0; mi < MT; ++mi) {
        if constexpr (FULL) {
          ah[0][mi][0] = wn[i][mi][0].x; ah[0][mi][1] = wn[i][mi][0].y; ah[0][mi][2] = wn[i][mi][0].z; ah[0][mi][3] = wn[i][mi][0].w;
          ah[1][mi][0] = wn[i][mi][1].x; ah[1][mi][1] = wn[i][mi][1].y; ah[1][mi][2] = wn[i][mi][1].z; ah[1][mi][3] = wn[i][mi][1].w;
          ahb[mi][0] = wn[i][mi][2].x; ahb[mi][1] = wn[i][mi][2].y; ahb[mi][2] = wn[i][mi][2].z; ahb[mi][3] = wn[i][mi][2].w;
          alb[mi][0] = wn[i][mi][3].x; alb[mi][1] = wn[i][mi][3].y; alb[mi][2] = wn[i][mi][3].z; alb[mi][3] = wn[i][mi][3].w;
        } else {
          uint32_t al[2][4];
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            tf32_split(__uint_as_float(wn[i][mi][j].x), ah[j][mi][0], al[j][0]);
            tf32_split(__uint_as_float(wn[i][mi][j].y), ah[j][mi][1], al[j][1]);
            tf32_split(__uint_as_float(wn[i][mi][j].z), ah[j][mi][2], al[j][2]);
            tf32_split(__uint_as_float(wn[i][mi][j].w), ah[j][mi][3], al[j][3]);
          }
          ahb[mi][0] = pack_bf16(ah[0][mi][0], ah[0][mi][2]); ahb[mi][1] = pack_bf16(ah[0][mi][1], ah[0][mi][3]);
          ahb[mi][2] = pack_bf16(ah[1][mi][0], ah[1][mi][2]); ahb[mi][3] = pack_bf16(ah[1][mi][1], ah[1][mi][3]);
          alb[mi][0] = pack_bf16(al[0][0], al[0][2]); alb[mi][1] = pack_bf16(al[0][1], al[0][3]);
          alb[mi][2] = pack_bf16(al[1][0], al[1][2]); alb[mi][3] = pack_bf16(al[1][1], al[1][3]);
        }
      }
      if (k16 + kD < nk16) {
#pragma unroll
        for (int mi = 0; mi < MT; ++mi)
#pragma unroll
          for (int w = 0; w < kW; ++w) wn[i][mi][w] = __ldg(wp + ((size_t)(mi * nk16 + k16 + kD) * kW + w) * 32);
      }
      uint32_t bh[2][NT][2], bhb[NT][2], blb[NT][2];
#pragma unroll
      for (int ni = 0; ni < NT; ++ni) {
        uint32_t bl[2][2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const float* q = bp[ni] + (size_t)(2 * k16 + j) * 8 * kNodesPad;
          tf32_split(q[0], bh[j][ni][0], bl[j][0]);
          tf32_split(q[4 * kNodesPad], bh[j][ni][1], bl[j][1]);
        }
        bhb[ni][0] = pack_bf16(bh[0][ni][0], bh[0][ni][1]); bhb[ni][1] = pack_bf16(bh[1][ni][0], bh[1][ni][1]);
        blb[ni][0] = pack_bf16(bl[0][0], bl[0][1]); blb[ni][1] = pack_bf16(bl[1][0], bl[1][1]);
      }
#pragma unroll
      for (int mi = 0; mi < MT; ++mi)
#pragma unroll
        for (int ni = 0; ni < NT; ++ni) mma_bf16(accc[mi][ni], alb[mi], bhb[ni][0], bhb[ni][1]);
#pragma unroll
      for (int mi = 0; mi < MT; ++mi)
#pragma unroll
        for (int ni = 0; ni < NT; ++ni) mma_bf16(accc[mi][ni], ahb[mi], blb[ni][0], blb[ni][1]);
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int mi = 0; mi < MT; ++mi)
#pragma unroll
          for (int ni = 0; ni < NT; ++ni) mma_tf32(acc[mi][ni], ah[j][mi], bh[j][ni][0], bh[j][ni][1]);
    }
  }
}

// Weight blocks of the backward chain as mma.sync A fragments: block 0..3 = W_g[:, F:] (contraction index j = row of
// gcn_kernel_g, unit = column F + m), blocks 4, 5, 6 = attention_wh, attention_wq, attention_w (ATT).  One warp per (block,
// unit tile, 16 contraction steps).
__global__ void pack_bwd_frag_kernel(BwdDir d, uint4* __restrict__ out, int F, int H, int nblocks, int full) {
  const int nt = H / 16, lane = threadIdx.x & 31, r = lane >> 2, c = lane & 3;
  const int tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (tile >= nblocks * nt * nt) return;
  const int blk = tile / (nt * nt), mt = (tile / nt) % nt, k16 = tile % nt;
  const float* w = blk < 4 ? d.Wg[blk] + F : blk == 4 ? d.Wh : blk == 5 ? d.Wq : d.Wa;
  const int ld = blk < 4 ? F + H : H;
  uint32_t hi[2][4], lo[2][4];
  float raw[2][4];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const float* q = w + (size_t)(k16 * 16 + 8 * j + c) * ld + mt * 16 + r;
    raw[j][0] = q[0]; raw[j][1] = q[8]; raw[j][2] = q[(size_t)4 * ld]; raw[j][3] = q[(size_t)4 * ld + 8];
#pragma unroll
    for (int e = 0; e < 4; ++e) tf32_split(raw[j][e], hi[j][e], lo[j][e]);
  }
  if (!full) {                       // the fp32 weights in fragment order
    float4* o4 = reinterpret_cast<float4*>(out) + (size_t)tile * 64 + lane;
    o4[0] = make_float4(raw[0][0], raw[0][1], raw[0][2], raw[0][3]);
    o4[32] = make_float4(raw[1][0], raw[1][1], raw[1][2], raw[1][3]);
    return;
  }
  uint4* o = out + (size_t)tile * 128 + lane;
  o[0] = make_uint4(hi[0][0], hi[0][1], hi[0][2], hi[0][3]);
  o[32] = make_uint4(hi[1][0], hi[1][1], hi[1][2], hi[1][3]);
  o[64] = make_uint4(pack_bf16(hi[0][0], hi[0][2]), pack_bf16(hi[0][1], hi[0][3]), pack_bf16(hi[1][0], hi[1][2]), pack_bf16(hi[1][1], hi[1][3]));
  o[96] = make_uint4(pack_bf16(lo[0][0], lo[0][2]), pack_bf16(lo[0][1], lo[0][3]), pack_bf16(lo[1][0], lo[1][2]), pack_bf16(lo[1][1], lo[1][3]));
}

// The two H x H vector products of the q chain (phases D, D2) as tensor-core products: out[m][n] = sum_k W[k][m] vec[n][k] with
// the sequences on the N side (n < BT <= 8, one n tile) -- a handful of mma instructions; the point is the weight fetch: each
// warp owns MT unit tiles for the whole contraction and keeps KD k8 tiles of weights in flight (the FFMA form reached
// 20 B/clk/SM on a 256 KB matrix).  Same TF32 head + bf16 correction scheme as contract_tf32<.., true>.
// acc[mi]: {(unit r, seq 2c), (unit r, seq 2c+1), (unit r+8, seq 2c), (unit r+8, seq 2c+1)}
template <int MT, int KD, bool PACKED>
__device__ __forceinline__ void gemv_tc(float (&acc)[MT][4], const float* __restrict__ vec, int BT, int H,
                                        const float* __restrict__ w, const uint4* __restrict__ wpk, int m0) {
  const int lane = threadIdx.x & 31, r = lane >> 2, c = lane & 3;
  const int nkt = H / 8, nk16 = H / 16;
  // PACKED: the fp32 weights in fragment order (pack_bwd_frag_kernel, full = 0): word j of (unit tile, 16 steps) = k8 tile j
  const float* wp = w + (size_t)c * H + m0 * 16 + r;
  const uint4* wq = wpk + (size_t)m0 * nk16 * 64 + lane;
  const size_t k4 = (size_t)4 * H, k8 = (size_t)8 * H;
  float wn[KD][MT][4];
  auto fetch = [&](int i, int kt) {
#pragma unroll
    for (int mi = 0; mi < MT; ++mi) {
      if constexpr (PACKED) {
        const uint4 v = __ldg(wq + ((size_t)(mi * nk16 + (kt >> 1)) * 2 + (kt & 1)) * 32);
        wn[i][mi][0] = __uint_as_float(v.x); wn[i][mi][1] = __uint_as_float(v.y);
        wn[i][mi][2] = __uint_as_float(v.z); wn[i][mi][3] = __uint_as_float(v.w);
      } else {
        const float* q = wp + (size_t)kt * k8 + mi * 16;
        wn[i][mi][0] = __ldg(q); wn[i][mi][1] = __ldg(q + 8); wn[i][mi][2] = __ldg(q + k4); wn[i][mi][3] = __ldg(q + k4 + 8);
      }
    }
  };
#pragma unroll
  for (int i = 0; i < KD; ++i) fetch(i, i);
  float accc[MT][4];
#pragma unroll
  for (int mi = 0; mi < MT; ++mi)
#pragma unroll
    for (int e = 0; e < 4; ++e) { acc[mi][e] = 0.f; accc[mi][e] = 0.f; }
  const float* bp = vec + (size_t)(r < BT ? r : 0) * H + c;
  const bool bv = r < BT;
#pragma unroll 1
  for (int kt0 = 0; kt0 < nkt; kt0 += KD) {
#pragma unroll
    for (int i = 0; i < KD; i += 2) {
      const int kt = kt0 + i;
      uint32_t ah[2][MT][4], ahb[MT][4], alb[MT][4];
#pragma unroll
      for (int mi = 0; mi < MT; ++mi) {
        uint32_t al[2][4];
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) tf32_split(wn[i + j][mi][e], ah[j][mi][e], al[j][e]);
        ahb[mi][0] = pack_bf16(ah[0][mi][0], ah[0][mi][2]); ahb[mi][1] = pack_bf16(ah[0][mi][1], ah[0][mi][3]);
        ahb[mi][2] = pack_bf16(ah[1][mi][0], ah[1][mi][2]); ahb[mi][3] = pack_bf16(ah[1][mi][1], ah[1][mi][3]);
        alb[mi][0] = pack_bf16(al[0][0], al[0][2]); alb[mi][1] = pack_bf16(al[0][1], al[0][3]);
        alb[mi][2] = pack_bf16(al[1][0], al[1][2]); alb[mi][3] = pack_bf16(al[1][1], al[1][3]);
      }
#pragma unroll
      for (int j = 0; j < 2; ++j)
        if (kt + j + KD < nkt) fetch(i + j, kt + j + KD);
      uint32_t bh[2][2], bl[2][2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float x0 = bv ? bp[(kt + j) * 8] : 0.f, x1 = bv ? bp[(kt + j) * 8 + 4] : 0.f;
        tf32_split(x0, bh[j][0], bl[j][0]);
        tf32_split(x1, bh[j][1], bl[j][1]);
      }
      const uint32_t bhb0 = pack_bf16(bh[0][0], bh[0][1]), bhb1 = pack_bf16(bh[1][0], bh[1][1]);
      const uint32_t blb0 = pack_bf16(bl[0][0], bl[0][1]), blb1 = pack_bf16(bl[1][0], bl[1][1]);
#pragma unroll
      for (int mi = 0; mi < MT; ++mi) mma_bf16(accc[mi], alb[mi], bhb0, bhb1);
#pragma unroll
      for (int mi = 0; mi < MT; ++mi) mma_bf16(accc[mi], ahb[mi], blb0, blb1);
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int mi = 0; mi < MT; ++mi) mma_tf32(acc[mi], ah[j][mi], bh[j][0], bh[j][1]);
    }
  }
#pragma unroll
  for (int mi = 0; mi < MT; ++mi)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[mi][e] += accc[mi][e];
}

// phases D and D2 on the tensor cores: dq = dqs Wq, dqp = dq [q > 0] (-> v2, gr.dqp);  ds = dqp Wa (-> v1)
template <int MT, bool PACKED>
__device__ __forceinline__ void bwd_phase_d_tc(float* v1, float* v2, int BT, int H, const BwdDir& d, int m0, bool active,
                                               const float* __restrict__ tq, float* __restrict__ dqp, int nvalid) {
  const int lane = threadIdx.x & 31, r = lane >> 2, c = lane & 3;
  float acc[MT][4];
  if (active) {
    // the relu mask's operand is fetched before the contraction
    float qv[MT][4];
#pragma unroll
    for (int mi = 0; mi < MT; ++mi)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int unit = (m0 + mi) * 16 + r + 8 * (e >> 1), sq = 2 * c + (e & 1);
        qv[mi][e] = sq < nvalid ? __ldg(tq + (size_t)sq * H + unit) : 0.f;
      }
    gemv_tc<MT, 8, PACKED>(acc, v1, BT, H, d.Wq, d.wfrag + (size_t)5 * (H / 16) * (H / 16) * 64, m0);
#pragma unroll
    for (int mi = 0; mi < MT; ++mi)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int unit = (m0 + mi) * 16 + r + 8 * (e >> 1), sq = 2 * c + (e & 1);
        if (sq < BT) {
          const float rv = (sq < nvalid && qv[mi][e] > 0.f) ? acc[mi][e] : 0.f;
          if (sq < nvalid) dqp[(size_t)sq * H + unit] = rv;
          v2[sq * H + unit] = rv;
        }
      }
  }
  __syncthreads();
  if (active) {
    gemv_tc<MT, 8, PACKED>(acc, v2, BT, H, d.Wa, d.wfrag + (size_t)6 * (H / 16) * (H / 16) * 64, m0);
#pragma unroll
    for (int mi = 0; mi < MT; ++mi)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int unit = (m0 + mi) * 16 + r + 8 * (e >> 1), sq = 2 * c + (e & 1);
        if (sq < BT) v1[sq * H + unit] = acc[mi][e];
      }
  }
}

// phase E on the tensor cores: dhy[n][j] += ds_j + sum_k dep[n][k] Wh[k][j]   (the barrier: v1 = ds is complete)
template <int MT, int NT, int MODE>
__device__ __forceinline__ void bwd_phase_e_tc(const float* dep, size_t HN, const BwdDir& d, int H, int m0, int nt0,
                                               float* dh, const float* v1) {
  float acc[MT][NT][4];
#pragma unroll
  for (int mi = 0; mi < MT; ++mi)
#pragma unroll
    for (int ni = 0; ni < NT; ++ni)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[mi][ni][e] = 0.f;
  if constexpr (MODE >= 3) contract_packed<MT, NT, MODE == 3>(acc, acc, dep, HN, nt0, d.wfrag + (size_t)4 * (H / 16) * (H / 16) * (MODE == 3 ? 128 : 64), H / 16, m0);
  else contract_tf32<MT, NT, MODE == 2>(acc, acc, dep, HN, nt0, d.Wh, H, m0, H);
  __syncthreads();
  const int lane = threadIdx.x & 31, r = lane >> 2, c = lane & 3;
#pragma unroll
  for (int mi = 0; mi < MT; ++mi)
#pragma unroll
    for (int ni = 0; ni < NT; ++ni)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int nt = nt0 + ni, unit = (m0 + mi) * 16 + r + 8 * hf, node = (nt & 1) * 8 + 2 * c;
        const int task = (nt >> 1) * H + unit;
        float2* p = reinterpret_cast<float2*>(dh + (size_t)task * kNodesPad + node);
        float2 cur = *p;
        const float ds = v1[task];
        cur.x += acc[mi][ni][2 * hf] + ds;
        cur.y = node + 1 < kNodes ? cur.y + acc[mi][ni][2 * hf + 1] + ds : 0.f;
        *p = cur;
      }
}

// phase G on the tensor cores: dh'_{prev}[n][k] = sum_g sum_j dzm_g[n][j] W_g[j][F + k], then the recurrent-dropout mask
template <int MT, int NT, int MODE>
__device__ __forceinline__ void bwd_phase_g_tc(const float* dzm, int BT, size_t HN, const BwdDir& d, int F, int H, int m0, int nt0,
                                               float* dh, const float* hmask_t, int b0, int B, int T) {
  float acc[MT][NT][4];
#pragma unroll
  for (int mi = 0; mi < MT; ++mi)
#pragma unroll
    for (int ni = 0; ni < NT; ++ni)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[mi][ni][e] = 0.f;
  const int lane = threadIdx.x & 31, r = lane >> 2, c = lane & 3;
  // the recurrent-dropout mask of this step for the thread's outputs is fetched before the contraction (hmask_t: mask
  // element (b = 0, this step, node 0, unit 0)), not after it
  float mk[MT][NT][4];
#pragma unroll
  for (int mi = 0; mi < MT; ++mi)
#pragma unroll
    for (int ni = 0; ni < NT; ++ni)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int nt = nt0 + ni, unit = (m0 + mi) * 16 + r + 8 * hf, node = (nt & 1) * 8 + 2 * c;
        const int b = b0 + (nt >> 1);
        mk[mi][ni][2 * hf] = 1.0f; mk[mi][ni][2 * hf + 1] = 1.0f;
        if (hmask_t != nullptr && b < B) {
          const float* mp = hmask_t + ((size_t)b * T * kNodes + node) * H + unit;
          mk[mi][ni][2 * hf] = __ldg(mp);
          if (node + 1 < kNodes) mk[mi][ni][2 * hf + 1] = __ldg(mp + H);
        }
      }
#pragma unroll 1
  for (int q = 0; q < 4; ++q) {
    if constexpr (MODE >= 3) contract_packed<MT, NT, MODE == 3>(acc, acc, dzm + (size_t)q * BT * HN, HN, nt0, d.wfrag + (size_t)q * (H / 16) * (H / 16) * (MODE == 3 ? 128 : 64), H / 16, m0);
    else contract_tf32<MT, NT, MODE == 2>(acc, acc, dzm + (size_t)q * BT * HN, HN, nt0, d.Wg[q] + F, F + H, m0, H);
  }
#pragma unroll
  for (int mi = 0; mi < MT; ++mi)
#pragma unroll
    for (int ni = 0; ni < NT; ++ni)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int nt = nt0 + ni, unit = (m0 + mi) * 16 + r + 8 * hf, node = (nt & 1) * 8 + 2 * c;
        const int s = nt >> 1;
        float2 v = make_float2(acc[mi][ni][2 * hf] * mk[mi][ni][2 * hf], node + 1 < kNodes ? acc[mi][ni][2 * hf + 1] * mk[mi][ni][2 * hf + 1] : 0.f);
        *reinterpret_cast<float2*>(dh + ((size_t)s * H + unit) * kNodesPad + node) = v;
      }
}

// Same chain as lstm_train_bwd_kernel for H in {64, 128, 256}, arranged around the two weight contractions that
// dominate it (phase E: dep x Wh, phase G: dzm x W[:, F:]).  TCORE != 0 (the default, 4): those two and the q chain's
// vector products run as mma.sync tensor-core products (contract_tf32 / contract_packed / gemv_tc above); the rest of
// this comment describes the FFMA form (TCORE = 0), which stays as the A/B reference.  BT * H = 512 (or 256): a thread owns 2 sequences x 2
// units (u, u + H/2) and 1/KS of the contraction range, so the weights cross L2 once per sequence PAIR and every
// activation float4 read from shared memory feeds 8 FMAs; the KS partial sums meet in shared memory.  dhy overwrites
// dh' in place and dep lives in dzm[0], which leaves room for two H=256 sequences per CTA.
template <bool ATT, int TCORE>      // TCORE: 0 = FFMA contractions, 1 = 3xTF32 mma.sync, 2 = TF32 head + bf16 corrections, 3 = 2 with pre-built weight fragments, 4 = 2 with the fp32 weights in fragment order
__global__ void __launch_bounds__(kThreads, 1)
lstm_train_bwd_blk_kernel(BwdDir d0, BwdDir d1, BwdGeom g) {
  extern __shared__ __align__(16) float smem[];
  const BwdDir d = blockIdx.y == 0 ? d0 : d1;
  const a3gc_tape& tp = g.tape;
  const a3gc_tape_grads& gr = g.gr;
  const int H = g.H, F = g.F, BT = g.BT, K = F + H;
  const size_t HN = (size_t)H * kNodesPad;
  float* dh = smem;                       // [BT][H][16]  gradient wrt h'_t carried from the later step; then dhy_t in place
  float* dc = dh + BT * HN;               // [BT][H][16]
  float* dzm = dc + BT * HN;              // [4][BT][H][16]
  float* dep = dzm;                       // (ATT) [BT][H][16] until phase F; dzm[1] is scratch for the GEMV partial sums
  float* scr = dzm + BT * HN;
  float* v1 = dzm + 4 * BT * HN;          // [BT][H]   dqs, then ds
  float* v2 = v1 + BT * H;                // [BT][H]   dqp
  float* abuf = v2 + BT * H;              // [BT][16]  dap
  float* al = abuf + BT * kNodesPad;      // [BT][16]  alpha
  float* red = al + BT * kNodesPad;       // [256]
  float* PT = red + 256;                  // [4][16][16]
  const int b0 = blockIdx.x * BT;
  const int y_off = blockIdx.y * H;
  const int ntask = BT * H;
  // blocked phases: thread -> (kh, sequence pair sp, unit pair u / u + UH)
  const int UH = H / 2, ntile = UH * (BT / 2), KS = kThreads / ntile;
  const int u = threadIdx.x % UH, s0 = 2 * ((threadIdx.x / UH) % (BT / 2)), kh = threadIdx.x / ntile;
  const int kq = H / KS, k0 = kh * kq, k1 = k0 + kq;
  // tensor-core phases: warp -> tc_mt unit tiles (16 units) from tc_m0 x tc_nt row tiles (8 rows) from tc_nt0:
  // H = 256: 2 x 4 (all rows of the CTA's two sequences), H = 128: 1 x 2 BT, H = 64: 1 x BT (warp pairs share a unit tile)
  const int tc_warp = threadIdx.x >> 5;
  const int tc_mt = H / 16 >= 8 ? H / 128 : 1, tc_m0 = H / 16 >= 8 ? tc_warp * tc_mt : tc_warp % 4;
  const int tc_nt = H / 16 >= 8 ? 2 * BT : BT, tc_nt0 = H / 16 >= 8 ? 0 : (tc_warp / 4) * BT;
  // GEMV phases: thread -> (kd, 4 consecutive units jd..jd+3), all BT sequences, 1/KD of the contraction range
  const int KD = 4 * kThreads / H, jd = 4 * (threadIdx.x % (H / 4)), kd = threadIdx.x / (H / 4);
  const int kdq = H / KD;
  LayerGeom lg; lg.B = g.B; lg.H = H; lg.BT = BT;
  load_state(dh, d.dhT, b0, lg);
  load_state(dc, d.dcT, b0, lg);
  for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x) PT[i] = d.PT[i];
  // diagnostics (A3GC_BWD_TRACE=1): cycles of CTA (0, 0) between the phase barriers, summed over the steps
  const bool tr = g.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0;
  long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tlast = tr ? clock64() : 0;
#define BWD_MARK(i) do { if (tr) { const long long now_ = clock64(); tacc[i] += now_ - tlast; tlast = now_; } } while (0)

  for (int step = g.T - 1; step >= 0; --step) {
    const int t = d.reverse ? g.T - 1 - step : step;
    const int tprev = d.reverse ? t + 1 : t - 1;
    const size_t rec0 = ((size_t)blockIdx.y * g.T + t) * g.B;
    const size_t recp0 = ((size_t)blockIdx.y * g.T + tprev) * g.B;
    const size_t nm0 = ((size_t)blockIdx.y * g.B * g.T + t) * kNodes;
    __syncthreads();
    // ---- A: dh' += dY (1 - y^2);  dhy = dh' (1 + alpha) (in place);  partial dalpha[n] = sum_j dh' hy
    if (ATT) {
      for (int i = threadIdx.x; i < BT * kNodesPad; i += blockDim.x) {
        const int b = b0 + i / kNodesPad;
        al[i] = b < g.B ? tp.a[(rec0 + b) * kNodesPad + i % kNodesPad] : 0.f;
      }
      __syncthreads();
    }
    for (int task = threadIdx.x; task < ntask; task += blockDim.x) {
      const int s = task / H, j = task % H, b = b0 + s;
      float dv[16];
      load16(dv, dh + (size_t)task * kNodesPad);
      if (b < g.B) {
        const float* hpp = tp.hp + (nm0 + (size_t)b * g.T * kNodes) * H + j;
        const float* yp = g.dy + (size_t)b * g.syb + (size_t)t * g.syt + y_off + j;
        // all 30 loads are issued before the first use (one DRAM latency per task instead of one per node)
        float gy[kNodes], hv[kNodes];
#pragma unroll
        for (int n = 0; n < kNodes; ++n) gy[n] = __ldg(yp + (size_t)n * g.yld);
        if (g.out_act == A3GC_ACT_TANH) {
#pragma unroll
          for (int n = 0; n < kNodes; ++n) hv[n] = __ldg(hpp + (size_t)n * H);
#pragma unroll
          for (int n = 0; n < kNodes; ++n) { const float y = fast_tanh(hv[n]); dv[n] = fmaf(gy[n], 1.0f - y * y, dv[n]); }
        } else {
#pragma unroll
          for (int n = 0; n < kNodes; ++n) dv[n] += gy[n];
        }
        dv[15] = 0.f;
        if (ATT) {
          float hh[16], pa[16], av[16];
          load16g(hh, tp.hh + ((rec0 + b) * H + j) * kNodesPad);
          load16(av, al + (size_t)s * kNodesPad);
#pragma unroll
          for (int n = 0; n < 16; ++n) { pa[n] = dv[n] * hh[n]; dv[n] *= 1.0f + av[n]; }
          store16(dep + (size_t)task * kNodesPad, pa);
        }
      } else {
#pragma unroll
        for (int n = 0; n < 16; ++n) dv[n] = 0.f;
        if (ATT) store16(dep + (size_t)task * kNodesPad, dv);
      }
      store16(dh + (size_t)task * kNodesPad, dv);
    }
    if (ATT) {
      __syncthreads();
      BWD_MARK(0);
      // ---- B: dalpha[n] = sum_j partial (all threads, then BT*16 of them);  dap = dalpha * a (1 - a)
      {
        const int nI = BT * kNodesPad, parts = kThreads / nI;
        const int i = threadIdx.x % nI, part = threadIdx.x / nI;
        const int s = i / kNodesPad, n = i % kNodesPad;
        const int jq = H / parts;
        const float* pv = dep + (size_t)s * HN + (size_t)part * jq * kNodesPad + n;
        float a = 0.f;
#pragma unroll 8
        for (int j = 0; j < jq; ++j) a += pv[(size_t)j * kNodesPad];
        red[threadIdx.x] = a;
        __syncthreads();
        if (threadIdx.x < nI) {
          float tot = 0.f;
          for (int q = 0; q < parts; ++q) tot += red[q * nI + i];
          const float sg = al[i];
          tot *= sg * (1.0f - sg);
          abuf[i] = tot;
          if (b0 + s < g.B) gr.dap[(rec0 + b0 + s) * kNodesPad + n] = tot;
        }
      }
      __syncthreads();
      BWD_MARK(1);
      // ---- C: dep[n][j] = dap[n] u_j (1 - e^2);  dqs_j = sum_n dep
      for (int task = threadIdx.x; task < ntask; task += blockDim.x) {
        const int s = task / H, j = task % H, b = b0 + s;
        float e[16], av[16], o[16];
        float sum = 0.f;
        if (b < g.B) {
          load16g(e, tp.e + ((rec0 + b) * H + j) * kNodesPad);
          load16(av, abuf + (size_t)s * kNodesPad);
          const float uj = d.u[j];
#pragma unroll
          for (int n = 0; n < 16; ++n) { o[n] = av[n] * uj * (1.0f - e[n] * e[n]); sum += o[n]; }
          store16g(gr.dep + ((rec0 + b) * H + j) * kNodesPad, o);
          gr.dqs[(rec0 + b) * H + j] = sum;
        } else {
#pragma unroll
          for (int n = 0; n < 16; ++n) o[n] = 0.f;
        }
        store16_swz(dep + (size_t)task * kNodesPad, o, TCORE != 0 && ((j >> 1) & 1));
        v1[task] = sum;
      }
      __syncthreads();
      BWD_MARK(2);
      if constexpr (TCORE >= 2) {
        // ---- D, D2 on the tensor cores (v1 = dqs -> v2 = dqp -> v1 = ds); the reads of v1 end before the barrier inside
        const int nval = g.B - b0 < BT ? g.B - b0 : BT;
        const float* tq = tp.q + (rec0 + b0) * H;
        float* dqp = gr.dqp + (rec0 + b0) * H;
        if (H >= 256) bwd_phase_d_tc<2, TCORE == 4>(v1, v2, BT, H, d, 2 * tc_warp, true, tq, dqp, nval);
        else bwd_phase_d_tc<1, TCORE == 4>(v1, v2, BT, H, d, tc_warp, tc_warp < H / 16, tq, dqp, nval);
        BWD_MARK(3);
      } else {
        // ---- D: dq_j = sum_k dqs_k Wq[k][j] for all BT sequences at once;  dqp = dq [q > 0]
        if (kdq % 8 == 0) gemv_part<8>(v1, d.Wq, scr, H, BT, jd, kd, kdq); else gemv_part<4>(v1, d.Wq, scr, H, BT, jd, kd, kdq);
        __syncthreads();
        for (int task = threadIdx.x; task < ntask; task += blockDim.x) {
          const int s = task / H, j = task % H, b = b0 + s;
          float dq = 0.f;
          for (int q = 0; q < KD; ++q) dq += scr[(q * BT + s) * H + j];
          float r = 0.f;
          if (b < g.B) {
            r = tp.q[(rec0 + b) * H + j] > 0.f ? dq : 0.f;
            gr.dqp[(rec0 + b) * H + j] = r;
          }
          v2[task] = r;
        }
        __syncthreads();
        BWD_MARK(3);
        // ---- D2: ds_j = sum_k dqp_k Wa[k][j]  (-> v1)
        if (kdq % 8 == 0) gemv_part<8>(v2, d.Wa, scr, H, BT, jd, kd, kdq); else gemv_part<4>(v2, d.Wa, scr, H, BT, jd, kd, kdq);
        __syncthreads();
        for (int task = threadIdx.x; task < ntask; task += blockDim.x) {
          float ds = 0.f;
          for (int q = 0; q < KD; ++q) ds += scr[(q * BT + task / H) * H + task % H];
          v1[task] = ds;
        }
      }
      BWD_MARK(4);
      // ---- E: dhy[n][j] += ds_j + sum_k dep[n][k] Wh[k][j]
      if constexpr (TCORE != 0) {
        if (tc_mt == 2) bwd_phase_e_tc<2, 4, TCORE>(dep, HN, d, H, tc_m0, tc_nt0, dh, v1);
        else if (tc_nt == 8) bwd_phase_e_tc<1, 8, TCORE>(dep, HN, d, H, tc_m0, tc_nt0, dh, v1);
        else bwd_phase_e_tc<1, 4, TCORE>(dep, HN, d, H, tc_m0, tc_nt0, dh, v1);
      } else {
        float acc[2][2][16];
#pragma unroll
        for (int si = 0; si < 2; ++si)
#pragma unroll
          for (int ui = 0; ui < 2; ++ui)
#pragma unroll
            for (int n = 0; n < 16; ++n) acc[si][ui][n] = 0.f;
        accum_blk<2, 2>(acc, dep + (size_t)s0 * HN, HN, d.Wh + u, H, UH, k0, k1);
        __syncthreads();                                           // v1 = ds complete
        for (int r = 0; r < KS; ++r) {
          if (kh == r) {
#pragma unroll
            for (int si = 0; si < 2; ++si)
#pragma unroll
              for (int ui = 0; ui < 2; ++ui) {
                const int task = (s0 + si) * H + u + ui * UH;
                float cur[16];
                load16(cur, dh + (size_t)task * kNodesPad);
                const float ds = r == 0 ? v1[task] : 0.f;
#pragma unroll
                for (int n = 0; n < kNodes; ++n) cur[n] += acc[si][ui][n] + ds;
                cur[15] = 0.f;
                store16(dh + (size_t)task * kNodesPad, cur);
              }
          }
          if (r + 1 < KS) __syncthreads();
        }
      }
    }
    __syncthreads();
    BWD_MARK(5);
    // ---- F: LSTM pointwise backward, dz (global, in place of the gates) and dzm = P_g^T dz_g
    for (int task = threadIdx.x; task < ntask; task += blockDim.x) {
      const int s = task / H, j = task % H, b = b0 + s;
      float dz[4][16];
      float dcv[16];
      if (b < g.B) {
        float gi[16], gf[16], gg[16], go[16], cc[16], cp[16], dv[16];
        float* gp = tp.gates + ((rec0 + b) * 4 * H + j) * kNodesPad;
        load16g(gi, gp); load16g(gf, gp + HN); load16g(gg, gp + 2 * HN); load16g(go, gp + 3 * HN);
        load16g(cc, tp.c + ((rec0 + b) * H + j) * kNodesPad);
        if (step > 0) load16g(cp, tp.c + ((recp0 + b) * H + j) * kNodesPad);
        else {
#pragma unroll
          for (int n = 0; n < 16; ++n) cp[n] = (d.c0 != nullptr && n < kNodes) ? d.c0[((size_t)b * kNodes + n) * H + j] : 0.f;
        }
        load16(dv, dh + (size_t)task * kNodesPad);
        load16(dcv, dc + (size_t)task * kNodesPad);
#pragma unroll
        for (int n = 0; n < 16; ++n) {
          const float tc = fast_tanh(cc[n]);
          const float dcn = fmaf(dv[n] * go[n], 1.0f - tc * tc, dcv[n]);
          dz[3][n] = dv[n] * tc * go[n] * (1.0f - go[n]);
          dz[0][n] = dcn * gg[n] * gi[n] * (1.0f - gi[n]);
          dz[1][n] = dcn * cp[n] * gf[n] * (1.0f - gf[n]);
          dz[2][n] = dcn * gi[n] * (1.0f - gg[n] * gg[n]);
          dcv[n] = dcn * gf[n];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) { dz[q][15] = 0.f; store16g(gp + q * HN, dz[q]); }
        dcv[15] = 0.f;
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int n = 0; n < 16; ++n) dz[q][n] = 0.f;
#pragma unroll
        for (int n = 0; n < 16; ++n) dcv[n] = 0.f;
      }
      store16(dc + (size_t)task * kNodesPad, dcv);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float m[16];
        mix15(dz[q], PT + q * 256, m);
        store16_swz(dzm + ((size_t)q * BT * H + task) * kNodesPad, m, TCORE != 0 && ((j >> 1) & 1));
        if (b < g.B) {
          const size_t zi = (nm0 + (size_t)b * g.T * kNodes) * 4 * H + (size_t)q * H + j;
#pragma unroll
          for (int n = 0; n < kNodes; ++n) emit_dzm(gr, zi + (size_t)n * 4 * H, m[n]);
        }
      }
    }
    __syncthreads();
    BWD_MARK(6);
    // ---- G: dh'_{prev}[n][k] = sum_g sum_j dzm_g[n][j] W_g[j][F + k]   (then the recurrent-dropout mask of this step)
    if constexpr (TCORE != 0) {
      const float* hm = g.hmask != nullptr ? g.hmask + nm0 * H : nullptr;
      if (tc_mt == 2) bwd_phase_g_tc<2, 4, TCORE>(dzm, BT, HN, d, F, H, tc_m0, tc_nt0, dh, hm, b0, g.B, g.T);
      else if (tc_nt == 8) bwd_phase_g_tc<1, 8, TCORE>(dzm, BT, HN, d, F, H, tc_m0, tc_nt0, dh, hm, b0, g.B, g.T);
      else bwd_phase_g_tc<1, 4, TCORE>(dzm, BT, HN, d, F, H, tc_m0, tc_nt0, dh, hm, b0, g.B, g.T);
    } else {
      float acc[2][2][16];
#pragma unroll
      for (int si = 0; si < 2; ++si)
#pragma unroll
        for (int ui = 0; ui < 2; ++ui)
#pragma unroll
          for (int n = 0; n < 16; ++n) acc[si][ui][n] = 0.f;
#pragma unroll 1
      for (int q = 0; q < 4; ++q) accum_blk<2, 2>(acc, dzm + ((size_t)q * BT + s0) * HN, HN, d.Wg[q] + F + u, K, UH, k0, k1);
      for (int r = 0; r < KS; ++r) {
        if (kh == r) {
#pragma unroll
          for (int si = 0; si < 2; ++si)
#pragma unroll
            for (int ui = 0; ui < 2; ++ui) {
              const int k = u + ui * UH, b = b0 + s0 + si;
              const int task = (s0 + si) * H + k;
              float cur[16];
              if (r > 0) {
                load16(cur, dh + (size_t)task * kNodesPad);
#pragma unroll
                for (int n = 0; n < 16; ++n) cur[n] += acc[si][ui][n];
              } else {
#pragma unroll
                for (int n = 0; n < 16; ++n) cur[n] = acc[si][ui][n];
              }
              if (r == KS - 1) {
                cur[15] = 0.f;
                if (g.hmask != nullptr && b < g.B) {
                  const float* mp = g.hmask + (nm0 + (size_t)b * g.T * kNodes) * H + k;
#pragma unroll
                  for (int n = 0; n < kNodes; ++n) cur[n] *= mp[(size_t)n * H];
                }
              }
              store16(dh + (size_t)task * kNodesPad, cur);
            }
        }
        if (r + 1 < KS) __syncthreads();
      }
    }
    BWD_MARK(7);
  }
  __syncthreads();
  if (tr) for (int i = 0; i < 8; ++i) g.trace[i] = tacc[i];
#undef BWD_MARK
  store_state(d.dh0, dh, b0, lg);
  store_state(d.dc0, dc, b0, lg);
}

// Reverse-time chain of the graph-GRU (net_aagc.py:343-368).  Per step, with dh' = carried gradient + dY_t:
//   du = dh' (h_prev - c), dc = dh' (1 - u), dzc = dc (1 - c^2), dzr = dzc zch r (1 - r), dzu = du u (1 - u), dzch = dzc r
//   dmsg = dzr Wrh + dzu Wuh + dzch Wch,  dM = P^T dmsg,  dh_prev = dh' u + dM Wg
// Stored for the hoisted GEMMs: (dzr, dzu, dzcx = dzc, dzch) node-major [D][B][T][15][4H] (grads.dzm), dmsg unit-major
// (grads.dep), dM node-major [D][B][T][15][H] (grads.dqs).
struct GruBwdDir {
  const float* Whid[3];    // dense_{r,u,c}_hid.weight [H][H]
  const float* Wg;         // gcn_kernel [H][H]
  const float* PT;         // [16][16]  PT[m][n] = P[n][m]
  const float* h0; const float* dhT; float* dh0;
  int reverse;
};

__global__ void __launch_bounds__(kThreads, 1)
gru_train_bwd_kernel(GruBwdDir d0, GruBwdDir d1, BwdGeom g) {
  extern __shared__ __align__(16) float smem[];
  const GruBwdDir d = blockIdx.y == 0 ? d0 : d1;
  const a3gc_tape& tp = g.tape;
  const a3gc_tape_grads& gr = g.gr;
  const int H = g.H, BT = g.BT;
  const size_t HN = (size_t)H * kNodesPad;
  float* dh = smem;                        // [BT][H][16]
  float* dzs = dh + BT * HN;               // [3][BT][H][16]  dzr, dzu, dzch
  float* dM = dzs + 3 * BT * HN;           // [BT][H][16]
  float* PT = dM + BT * HN;                // [16][16]
  const int b0 = blockIdx.x * BT;
  const int y_off = blockIdx.y * H;
  const int ntask = BT * H;
  LayerGeom lg; lg.B = g.B; lg.H = H; lg.BT = BT;
  load_state(dh, d.dhT, b0, lg);
  for (int i = threadIdx.x; i < 256; i += blockDim.x) PT[i] = d.PT[i];

  for (int step = g.T - 1; step >= 0; --step) {
    const int t = d.reverse ? g.T - 1 - step : step;
    const int tprev = d.reverse ? t + 1 : t - 1;
    const size_t rec0 = ((size_t)blockIdx.y * g.T + t) * g.B;
    const size_t nm0 = ((size_t)blockIdx.y * g.B * g.T + t) * kNodes;
    __syncthreads();
    // ---- A: gate gradients; dh <- dh' u (the direct path to h_prev)
    for (int task = threadIdx.x; task < ntask; task += blockDim.x) {
      const int s = task / H, j = task % H, b = b0 + s;
      float dv[16], zr[16], zu[16], zx[16], zh[16];
      load16(dv, dh + (size_t)task * kNodesPad);
      if (b < g.B) {
        float r[16], u[16], c[16], zch[16];
        const float* gp = tp.gates + (((rec0 + b) * 4) * H + j) * kNodesPad;
        load16(r, gp); load16(u, gp + HN); load16(c, gp + 2 * HN); load16(zch, gp + 3 * HN);
        const float* yp = g.dy + (size_t)b * g.syb + (size_t)t * g.syt + y_off + j;
        const float* hprev = step > 0 ? tp.hp + ((((size_t)blockIdx.y * g.B + b) * g.T + tprev) * kNodes) * H + j : nullptr;
#pragma unroll
        for (int n = 0; n < 16; ++n) {
          float hp = 0.f, gy = 0.f;
          if (n < kNodes) {
            gy = __ldg(yp + (size_t)n * g.yld);
            hp = step > 0 ? hprev[(size_t)n * H] : (d.h0 != nullptr ? d.h0[((size_t)b * kNodes + n) * H + j] : 0.f);
          }
          const float dhp = dv[n] + gy;
          const float dzc = dhp * (1.0f - u[n]) * (1.0f - c[n] * c[n]);
          zr[n] = dzc * zch[n] * r[n] * (1.0f - r[n]);
          zu[n] = dhp * (hp - c[n]) * u[n] * (1.0f - u[n]);
          zx[n] = dzc;
          zh[n] = dzc * r[n];
          dv[n] = dhp * u[n];
        }
        zr[15] = zu[15] = zx[15] = zh[15] = dv[15] = 0.f;
        float* zp = gr.dzm + (nm0 + (size_t)b * g.T * kNodes) * 4 * H + j;
#pragma unroll
        for (int n = 0; n < kNodes; ++n) {
          zp[(size_t)n * 4 * H] = zr[n]; zp[(size_t)n * 4 * H + H] = zu[n]; zp[(size_t)n * 4 * H + 2 * H] = zx[n]; zp[(size_t)n * 4 * H + 3 * H] = zh[n];
        }
      } else {
#pragma unroll
        for (int n = 0; n < 16; ++n) { zr[n] = zu[n] = zh[n] = dv[n] = 0.f; }
      }
      store16(dh + (size_t)task * kNodesPad, dv);
      store16(dzs + (size_t)task * kNodesPad, zr);
      store16(dzs + ((size_t)BT * H + task) * kNodesPad, zu);
      store16(dzs + ((size_t)2 * BT * H + task) * kNodesPad, zh);
    }
    __syncthreads();
    // ---- B: dmsg[n][k] = sum_g sum_j dz_g[n][j] Whid_g[j][k];  dM = P^T dmsg
    for (int task = threadIdx.x; task < ntask; task += blockDim.x) {
      const int s = task / H, k = task % H, b = b0 + s;
      float acc[16], m[16];
#pragma unroll
      for (int n = 0; n < 16; ++n) acc[n] = 0.f;
#pragma unroll
      for (int q = 0; q < 3; ++q) accum1(acc, dzs + ((size_t)q * BT + s) * HN, d.Whid[q] + k, H, H);
      acc[15] = 0.f;
      mix15(acc, PT, m);
      store16(dM + (size_t)task * kNodesPad, m);
      if (b < g.B) {
        store16(gr.dep + ((rec0 + b) * H + k) * kNodesPad, acc);
        float* mp = gr.dqs + (nm0 + (size_t)b * g.T * kNodes) * H + k;
#pragma unroll
        for (int n = 0; n < kNodes; ++n) mp[(size_t)n * H] = m[n];
      }
    }
    __syncthreads();
    // ---- C: dh_prev[n][k'] = dh' u + sum_j dM[n][j] Wg[j][k']
    for (int task = threadIdx.x; task < ntask; task += blockDim.x) {
      const int s = task / H, k = task % H;
      float acc[16];
      load16(acc, dh + (size_t)task * kNodesPad);
      accum1(acc, dM + (size_t)s * HN, d.Wg + k, H, H);
      acc[15] = 0.f;
      store16(dh + (size_t)task * kNodesPad, acc);
    }
  }
  __syncthreads();
  store_state(d.dh0, dh, b0, lg);
}

__global__ void pack_gru_pt_kernel(a3gc_cell_params cp, float* out) {
  // P[n][m] = g_adjacency[m][n] (pack_gru_kernel)  ->  PT[m][n] = P[n][m] = g_adjacency[m][n]
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    const int m = i / 16, n = i % 16;
    out[i] = (m < kNodes && n < kNodes) ? cp.g_adjacency[m * kNodes + n] : 0.f;
  }
}

__global__ void pack_pt_kernel(a3gc_cell_params cp, float* out, int variant) {
  // PT_g[n][m] = P_g[m][n] with P as in pack_lstm_kernel
  for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x) {
    const int g = i / 256, n = (i % 256) / 16, m = i % 16;
    float v = 0.f;
    if (m < kNodes && n < kNodes)
      v = (variant == A3GC_VARIANT_AGC) ? cp.adjacency[0][n * kNodes + m] : cp.adjacency[g][m * kNodes + n];
    out[i] = v;
  }
}

// ------------------------------------------------------------------------------------------
// G-GRU time loop
// ------------------------------------------------------------------------------------------
// TRAIN: also keeps the tape of the graph-GRU backward (a3gc_tape fields, G-GRU meaning): gates = (r, u, c, zch = Wch msg),
// c = M = h Wg^T (before the node mix), hh = msg, hp = h' (node-major)
template <bool TRAIN>
__global__ void __launch_bounds__(kThreads, 1)
gru_layer_kernel(GruPacked w0, GruPacked w1, DirPtrs d0, DirPtrs d1, LayerGeom g, a3gc_tape tp) {
  extern __shared__ __align__(16) float smem[];
  const GruPacked w = blockIdx.y == 0 ? w0 : w1;
  const DirPtrs d = blockIdx.y == 0 ? d0 : d1;
  const int H = g.H, F = g.F, BT = g.BT;
  float* hbuf = smem;                                    // [2][BT][H][16]
  float* mbuf = hbuf + (size_t)2 * BT * H * kNodesPad;   // [BT][H][16]  message
  float* xbuf = mbuf + (size_t)BT * H * kNodesPad;       // [BT][F][16]
  float* Pbuf = xbuf + (size_t)BT * F * kNodesPad;       // [16][16]
  const int b0 = blockIdx.x * BT;
  const int y_off = blockIdx.y * H;
  load_state(hbuf, d.h0, b0, g);
  for (int i = threadIdx.x; i < 256; i += blockDim.x) Pbuf[i] = w.P[i];
  int cur = 0;
  const int ntask = BT * H;
  for (int step = 0; step < g.T; ++step) {
    const int t = d.reverse ? g.T - 1 - step : step;
    float* hcur = hbuf + (size_t)cur * BT * H * kNodesPad;
    float* hnxt = hbuf + (size_t)(cur ^ 1) * BT * H * kNodesPad;
    __syncthreads();
    load_x(xbuf, b0, t, g);
    // msg = adjacency^T (h W_g^T)   (net_aagc.py:347-348)
    for (int task = threadIdx.x; task < ntask; task += blockDim.x) {
      const int s = task / H, j = task % H;
      float m[16], msg[16];
#pragma unroll
      for (int n = 0; n < 16; ++n) m[n] = 0.f;
      accum1(m, hcur + (size_t)s * H * kNodesPad, w.Wg_t + j, H, H);
      mix15(m, Pbuf, msg);
      store16(mbuf + (size_t)task * kNodesPad, msg);
      if (TRAIN && b0 + s < g.B) {
        const size_t rec = ((size_t)blockIdx.y * g.T + t) * g.B + b0 + s;
        m[15] = 0.f;
        store16(tp.c + (rec * H + j) * kNodesPad, m);
        store16(tp.hh + (rec * H + j) * kNodesPad, msg);
      }
    }
    __syncthreads();
    for (int task = threadIdx.x; task < ntask; task += blockDim.x) {
      const int s = task / H, j = task % H;
      float acc[4][16];
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int n = 0; n < 16; ++n) acc[q][n] = 0.f;
      accum4(acc, xbuf + (size_t)s * F * kNodesPad, w.Win4 + j, F, H);       // r, u, c_in
      accum4(acc, mbuf + (size_t)s * H * kNodesPad, w.Whid4 + j, H, H);      // r, u, -, c_hid
      const float4 bias = w.bias4[j];
      float hold[16], hy[16];
      float gr_[16], gu_[16], gc_[16];
      load16(hold, hcur + (size_t)task * kNodesPad);
#pragma unroll
      for (int n = 0; n < kNodes; ++n) {
        const float r = sigmoidf_(acc[0][n] + bias.x);
        const float u = sigmoidf_(acc[1][n] + bias.y);
        const float c = tanhf_(acc[2][n] + bias.z + r * acc[3][n]);
        hy[n] = u * hold[n] + (1.0f - u) * c;                                  // net_aagc.py:364
        if (TRAIN) { gr_[n] = r; gu_[n] = u; gc_[n] = c; }
      }
      hy[15] = 0.f;
      store16(hnxt + (size_t)task * kNodesPad, hy);
      const int b = b0 + s;
      if (b < g.B) {
        float* yp = g.y + (size_t)b * g.syb + (size_t)t * g.syt + y_off + j;
#pragma unroll
        for (int n = 0; n < kNodes; ++n) yp[(size_t)n * g.yld] = hy[n];        // returns (h, h): no activation
        if (TRAIN) {
          const size_t rec = ((size_t)blockIdx.y * g.T + t) * g.B + b;
          float* gp = tp.gates + ((rec * 4) * H + j) * kNodesPad;
          gr_[15] = 0.f; gu_[15] = 0.f; gc_[15] = 0.f; acc[3][15] = 0.f;
          store16(gp, gr_); store16(gp + (size_t)H * kNodesPad, gu_); store16(gp + (size_t)2 * H * kNodesPad, gc_);
          store16(gp + (size_t)3 * H * kNodesPad, acc[3]);
          float* hpp = tp.hp + ((((size_t)blockIdx.y * g.B + b) * g.T + t) * kNodes) * H + j;
#pragma unroll
          for (int n = 0; n < kNodes; ++n) hpp[(size_t)n * H] = hy[n];
        }
      }
    }
    cur ^= 1;
  }
  __syncthreads();
  store_state(d.hT, hbuf + (size_t)cur * BT * H * kNodesPad, b0, g);
}

// ------------------------------------------------------------------------------------------
// AAGC graph convolution: y = act((adj @ x) @ W^T + b)    (net_aagc.py:61-66)
// one CTA processes FR frames at a time: stage x, mix over nodes, then 15*f_out dot products
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
gc_kernel(a3gc_gc_params p, const float* __restrict__ x, float* __restrict__ y, int64_t frames,
          int f_in, int f_out, int act, int FR) {
  extern __shared__ __align__(16) float smem[];
  float* adj = smem;                       // [15][16]
  float* xs = adj + 256;                   // [FR][15][f_in]
  float* xm = xs + (size_t)FR * kNodes * f_in;
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    int m = i / 16, n = i % 16;
    adj[i] = (m < kNodes && n < kNodes) ? p.adj[m * kNodes + n] : 0.f;
  }
  const int per_frame = kNodes * f_in;
  for (int64_t f0 = (int64_t)blockIdx.x * FR; f0 < frames; f0 += (int64_t)gridDim.x * FR) {
    const int nf = (int)((frames - f0) < FR ? (frames - f0) : FR);
    __syncthreads();
    for (int i = threadIdx.x; i < nf * per_frame; i += blockDim.x) xs[i] = __ldg(x + (size_t)f0 * per_frame + i);
    __syncthreads();
    for (int i = threadIdx.x; i < nf * per_frame; i += blockDim.x) {
      const int k = i % f_in, m = (i / f_in) % kNodes, fr = i / per_frame;
      const float* col = xs + (size_t)fr * per_frame + k;
      float s = 0.f;
#pragma unroll
      for (int n = 0; n < kNodes; ++n) s = fmaf(adj[m * 16 + n], col[n * f_in], s);
      xm[i] = s;
    }
    __syncthreads();
    const int outs = nf * kNodes * f_out;
    for (int i = threadIdx.x; i < outs; i += blockDim.x) {
      const int o = i % f_out, row = i / f_out;      // row = fr*15 + m
      const float* a = xm + (size_t)row * f_in;
      const float* wr = p.gcn_kernel + (size_t)o * f_in;
      float s = p.gcn_bias[o];
      for (int k = 0; k < f_in; ++k) s = fmaf(a[k], __ldg(wr + k), s);
      y[((size_t)f0 * kNodes + row) * f_out + o] = apply_act(s, act);
    }
  }
}

// prepare_input (evaluate_a3gc_tp.py:64-94): one thread per (frame, node, channel<12)
__global__ void prepare_input_kernel(const float* __restrict__ acc, const float* __restrict__ ori,
                                     const float* acc_mean, const float* acc_std, const float* ori_mean,
                                     const float* ori_std, float* __restrict__ x, int64_t frames, int ld_x) {
  // node -> IMU index; input_joints = [3, 4, 13, 14, 10]  (evaluate_a3gc_tp.py:65)
  const int64_t total = frames * kNodes * 12;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % 12), n = (int)((i / 12) % kNodes);
    const int64_t f = i / (12 * kNodes);
    int imu = -1;
    switch (n) { case 3: imu = 0; break; case 4: imu = 1; break; case 13: imu = 2; break; case 14: imu = 3; break; case 10: imu = 4; break; default: break; }
    float v = 0.f;
    if (imu >= 0) {
      if (c < 3) {
        const int ch = imu * 3 + c;
        v = acc[f * 18 + ch];
        if (acc_mean != nullptr) v = (v - acc_mean[ch]) / acc_std[ch];
      } else {
        const int ch = imu * 9 + (c - 3);
        v = ori[f * 54 + ch];
        if (ori_mean != nullptr) v = (v - ori_mean[ch]) / ori_std[ch];
      }
    }
    x[(f * kNodes + n) * ld_x + c] = v;
  }
}

__global__ void concat_stage_input_kernel(const float* __restrict__ x, const float* __restrict__ pos,
                                          float* __restrict__ dst, int64_t rows) {
  const int64_t total = rows * 15;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % 15);
    const int64_t r = i / 15;
    dst[i] = c < 12 ? x[r * 12 + c] : pos[r * 3 + (c - 12)];
  }
}

// reduced-global -> full-local pose (PoseNet3._reduced_glb_to_full_local_mat, net_aagc.py:788-800): scatter the 15
// predicted global rotations into the 24 SMPL joints (identity elsewhere), R_local[i] = R_global[parent[i]]^T R_global[i]
// along the SMPL tree (articulate/math/spatial.py:115-123), identity on the ignored joints.  One thread per (frame, joint).
__constant__ int c_smpl_parent[24] = {-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21};   // SMPL kintree_table[0]
__constant__ int c_full_to_reduced[24] = {-1, 0, 1, 2, 3, 4, 5, -1, -1, 6, -1, -1, 7, 8, 9, 10, 11, 12, 13, 14, -1, -1, -1, -1};   // inverse of joint_set.reduced (config.py:29)

__device__ __forceinline__ void load_global_rot(const float* __restrict__ pose, int64_t f, int joint, int rotsize, float (&R)[9]) {
  const int r = joint < 0 ? -1 : c_full_to_reduced[joint];
  if (r < 0) {
    R[0] = 1.f; R[1] = 0.f; R[2] = 0.f; R[3] = 0.f; R[4] = 1.f; R[5] = 0.f; R[6] = 0.f; R[7] = 0.f; R[8] = 1.f;
    return;
  }
  const float* p = pose + ((size_t)f * kNodes + r) * rotsize;
  if (rotsize == 9) {
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = p[i];
    return;
  }
  // 6D -> rotation matrix (articulate/math/angular.py:167-182): columns c0, c1, c0 x c1; NaN -> 0
  float a[3] = {p[0], p[1], p[2]}, b[3] = {p[3], p[4], p[5]};
  const float na = sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
  float c0[3] = {a[0] / na, a[1] / na, a[2] / na};
  const float dt = c0[0] * b[0] + c0[1] * b[1] + c0[2] * b[2];
  float v[3] = {b[0] - dt * c0[0], b[1] - dt * c0[1], b[2] - dt * c0[2]};
  const float nv = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  float c1[3] = {v[0] / nv, v[1] / nv, v[2] / nv};
  float c2[3] = {c0[1] * c1[2] - c0[2] * c1[1], c0[2] * c1[0] - c0[0] * c1[2], c0[0] * c1[1] - c0[1] * c1[0]};
#pragma unroll
  for (int i = 0; i < 3; ++i) { R[3 * i] = c0[i]; R[3 * i + 1] = c1[i]; R[3 * i + 2] = c2[i]; }
#pragma unroll
  for (int i = 0; i < 9; ++i) if (R[i] != R[i]) R[i] = 0.f;
}

__global__ void reduced_to_full_local_kernel(const float* __restrict__ pose, float* __restrict__ out, int64_t frames, int rotsize) {
  const int64_t total = frames * 24;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t f = i / 24;
    const int j = (int)(i % 24);
    float L[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
    if (c_full_to_reduced[j] >= 0) {                      // joints outside joint_set.reduced are exactly joint_set.ignored
      float G[9], P[9];
      load_global_rot(pose, f, j, rotsize, G);
      load_global_rot(pose, f, c_smpl_parent[j], rotsize, P);
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) L[3 * r + c] = P[r] * G[c] + P[3 + r] * G[3 + c] + P[6 + r] * G[6 + c];   // (P^T G)[r][c]
    }
    float* o = out + (size_t)i * 9;
#pragma unroll
    for (int k = 0; k < 9; ++k) o[k] = L[k];
  }
}

int max_optin_smem() {
  int dev = 0, v = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return 0;
  return v;
}

}  // namespace

size_t simt_layer_workspace_bytes(int variant, int f_in, int hidden, int num_dirs) {
  size_t per = variant == A3GC_VARIANT_GGRU ? gru_packed_floats(f_in, hidden) : lstm_packed_floats(f_in, hidden);
  return (size_t)num_dirs * align_up(per * sizeof(float), 256);
}

int simt_layer_forward(const LayerArgs& a, void* ws, size_t ws_bytes, cudaStream_t stream) {
  const int F = a.f_in, H = a.hidden;
  const size_t need = simt_layer_workspace_bytes(a.variant, F, H, a.num_dirs);
  if (ws_bytes < need || ws == nullptr) {
    set_error("a3gc_layer_forward: workspace too small (%zu < %zu bytes)", ws_bytes, need);
    return A3GC_ERR_WORKSPACE;
  }
  const int smem_max = max_optin_smem();
  if (smem_max <= 0) { set_error("no CUDA device"); return A3GC_ERR_NO_DEVICE; }
  const bool gru = a.variant == A3GC_VARIANT_GGRU;
  const bool att = a.variant == A3GC_VARIANT_A3GC || a.variant == A3GC_VARIANT_AGC;
  const int KX = F > H ? F : H;
  // shared-memory bytes as a function of the batch tile
  auto smem_bytes = [&](int bt) -> size_t {
    size_t fl = gru ? ((size_t)3 * bt * H * 16 + (size_t)bt * F * 16 + 256)
                    : ((size_t)3 * bt * H * 16 + (size_t)bt * KX * 16 + (size_t)2 * bt * H + (size_t)bt * 16 + 1024);
    return fl * sizeof(float);
  };
  int BT = 8;
  while (BT > 1 && smem_bytes(BT) > (size_t)smem_max) --BT;
  // do not leave SMs idle when the batch is small: shrink the tile until the grid covers the chip
  int sms = 148;
  { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  while (BT > 1 && ((a.batch + BT - 1) / BT) * a.num_dirs < sms) --BT;
  if (smem_bytes(BT) > (size_t)smem_max) {
    set_error("SIMT engine: hidden=%d f_in=%d needs %zu bytes of shared memory (> %d)", H, F, smem_bytes(1), smem_max);
    return A3GC_ERR_UNSUPPORTED;
  }
  char* base = static_cast<char*>(ws);
  const size_t per = need / a.num_dirs;
  DirPtrs dp[2] = {};
  for (int d = 0; d < a.num_dirs; ++d) {
    dp[d].h0 = a.h0[d]; dp[d].c0 = a.c0[d]; dp[d].hT = a.hT[d]; dp[d].cT = a.cT[d]; dp[d].reverse = a.reverse[d];
  }
  LayerGeom g;
  g.x = a.x; g.sxb = a.x_stride_b; g.sxt = a.x_stride_t;
  g.y = a.y; g.syb = a.y_stride_b; g.syt = a.y_stride_t; g.yld = a.y_ld;
  g.B = (int)a.batch; g.T = (int)a.steps; g.F = F; g.H = H; g.out_act = a.out_act; g.BT = BT;
  dim3 grid((unsigned)((a.batch + BT - 1) / BT), (unsigned)a.num_dirs);
  const size_t smem = smem_bytes(BT);
  if (gru) {
    GruPacked pk[2];
    for (int d = 0; d < a.num_dirs; ++d) {
      pk[d] = carve_gru(reinterpret_cast<float*>(base + d * per), F, H);
      pack_gru_kernel<<<64, 256, 0, stream>>>(a.cells[d], pk[d], F, H);
      A3GC_LAUNCH_CHECK("pack_gru_kernel");
    }
    if (a.num_dirs == 1) pk[1] = pk[0];
    A3GC_CUDA_TRY(cudaFuncSetAttribute(gru_layer_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gru_layer_kernel<false><<<grid, kThreads, smem, stream>>>(pk[0], pk[1], dp[0], dp[1], g, a3gc_tape{});
    A3GC_LAUNCH_CHECK("gru_layer_kernel");
  } else {
    LstmPacked pk[2];
    for (int d = 0; d < a.num_dirs; ++d) {
      pk[d] = carve_lstm(reinterpret_cast<float*>(base + d * per), F, H);
      pack_lstm_kernel<<<64, 256, 0, stream>>>(a.cells[d], pk[d], F, H, a.variant);
      A3GC_LAUNCH_CHECK("pack_lstm_kernel");
    }
    if (a.num_dirs == 1) pk[1] = pk[0];
    if (att) {
      A3GC_CUDA_TRY(cudaFuncSetAttribute(lstm_layer_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      lstm_layer_kernel<true><<<grid, kThreads, smem, stream>>>(pk[0], pk[1], dp[0], dp[1], g);
    } else {
      A3GC_CUDA_TRY(cudaFuncSetAttribute(lstm_layer_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      lstm_layer_kernel<false><<<grid, kThreads, smem, stream>>>(pk[0], pk[1], dp[0], dp[1], g);
    }
    A3GC_LAUNCH_CHECK("lstm_layer_kernel");
  }
  return A3GC_OK;
}


// ------------------------------------------------------------------------------------------
// training path: host side
// ------------------------------------------------------------------------------------------
size_t simt_train_workspace_bytes(int variant, int f_in, int hidden, int num_dirs) {
  const size_t packed = variant == A3GC_VARIANT_GGRU ? gru_packed_floats(f_in, hidden) : lstm_packed_floats(f_in, hidden);
  // + (LSTM family) the seven H x H weight blocks of the backward chain as mma.sync fragments: up to 8 bytes per weight
  const size_t frag = variant == A3GC_VARIANT_GGRU ? 0 : align_up((size_t)7 * hidden * hidden * 8, 256);
  const size_t per = align_up(packed * sizeof(float), 256) + 4 * 256 * sizeof(float) + frag;
  return (size_t)num_dirs * per;
}

// graph-GRU training forward: the SIMT layer kernel with the tape
static int simt_gru_train_forward(const LayerArgs& a, const a3gc_tape& tape, void* ws, size_t ws_bytes, cudaStream_t stream) {
  const int F = a.f_in, H = a.hidden;
  const size_t need = simt_train_workspace_bytes(a.variant, F, H, a.num_dirs);
  if (ws_bytes < need || ws == nullptr) { set_error("a3gc_layer_train_forward: workspace too small (%zu < %zu bytes)", ws_bytes, need); return A3GC_ERR_WORKSPACE; }
  const int smem_max = max_optin_smem();
  if (smem_max <= 0) { set_error("no CUDA device"); return A3GC_ERR_NO_DEVICE; }
  auto smem_bytes = [&](int bt) -> size_t { return ((size_t)3 * bt * H * 16 + (size_t)bt * F * 16 + 256) * sizeof(float); };
  int BT = 8;
  while (BT > 1 && smem_bytes(BT) > (size_t)smem_max) --BT;
  int sms = 148;
  { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  while (BT > 1 && ((a.batch + BT - 1) / BT) * a.num_dirs < sms) --BT;
  if (smem_bytes(BT) > (size_t)smem_max) { set_error("G-GRU training forward: hidden=%d f_in=%d does not fit in shared memory", H, F); return A3GC_ERR_UNSUPPORTED; }
  char* base = static_cast<char*>(ws);
  const size_t per = need / a.num_dirs;
  DirPtrs dp[2] = {};
  GruPacked pk[2];
  for (int d = 0; d < a.num_dirs; ++d) {
    dp[d].h0 = a.h0[d]; dp[d].hT = a.hT[d]; dp[d].reverse = a.reverse[d];
    pk[d] = carve_gru(reinterpret_cast<float*>(base + d * per), F, H);
    pack_gru_kernel<<<64, 256, 0, stream>>>(a.cells[d], pk[d], F, H);
    A3GC_LAUNCH_CHECK("pack_gru_kernel");
  }
  if (a.num_dirs == 1) pk[1] = pk[0];
  LayerGeom g;
  g.x = a.x; g.sxb = a.x_stride_b; g.sxt = a.x_stride_t;
  g.y = a.y; g.syb = a.y_stride_b; g.syt = a.y_stride_t; g.yld = a.y_ld;
  g.B = (int)a.batch; g.T = (int)a.steps; g.F = F; g.H = H; g.out_act = a.out_act; g.BT = BT;
  dim3 grid((unsigned)((a.batch + BT - 1) / BT), (unsigned)a.num_dirs);
  const size_t smem = smem_bytes(BT);
  A3GC_CUDA_TRY(cudaFuncSetAttribute(gru_layer_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gru_layer_kernel<true><<<grid, kThreads, smem, stream>>>(pk[0], pk[1], dp[0], dp[1], g, tape);
  A3GC_LAUNCH_CHECK("gru_layer_kernel");
  return A3GC_OK;
}

static int simt_gru_train_backward(const TrainBwdArgs& a, void* ws, size_t ws_bytes, cudaStream_t stream) {
  const int F = a.f_in, H = a.hidden;
  const size_t need = simt_train_workspace_bytes(a.variant, F, H, a.num_dirs);
  if (ws_bytes < need || ws == nullptr) { set_error("a3gc_layer_backward: workspace too small (%zu < %zu bytes)", ws_bytes, need); return A3GC_ERR_WORKSPACE; }
  const int smem_max = max_optin_smem();
  if (smem_max <= 0) { set_error("no CUDA device"); return A3GC_ERR_NO_DEVICE; }
  auto smem_bytes = [&](int bt) -> size_t { return ((size_t)5 * bt * H * 16 + 256) * sizeof(float); };
  int BT = 8;
  while (BT > 1 && smem_bytes(BT) > (size_t)smem_max) --BT;
  int sms = 148;
  { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  while (BT > 1 && ((a.batch + BT - 1) / BT) * a.num_dirs < sms) --BT;
  if (smem_bytes(BT) > (size_t)smem_max) { set_error("G-GRU training backward: hidden=%d does not fit in shared memory", H); return A3GC_ERR_UNSUPPORTED; }
  char* base = static_cast<char*>(ws);
  const size_t per = need / a.num_dirs;
  GruBwdDir bd[2] = {};
  for (int d = 0; d < a.num_dirs; ++d) {
    float* pt = reinterpret_cast<float*>(base + d * per + align_up(gru_packed_floats(F, H) * sizeof(float), 256));
    pack_gru_pt_kernel<<<1, 256, 0, stream>>>(a.cells[d], pt);
    A3GC_LAUNCH_CHECK("pack_gru_pt_kernel");
    for (int q = 0; q < 3; ++q) bd[d].Whid[q] = a.cells[d].dense_hid_w[q];
    bd[d].Wg = a.cells[d].g_gcn_kernel; bd[d].PT = pt;
    bd[d].h0 = a.c0[d];                               // for G-GRU the caller passes the forward's initial h here
    bd[d].dhT = a.dhT[d]; bd[d].dh0 = a.dh0[d]; bd[d].reverse = a.reverse[d];
  }
  if (a.num_dirs == 1) bd[1] = bd[0];
  BwdGeom g;
  g.dy = a.dy; g.syb = a.dy_stride_b; g.syt = a.dy_stride_t; g.yld = a.dy_ld;
  g.tape = a.tape; g.gr = a.grads; g.hmask = nullptr;
  g.B = (int)a.batch; g.T = (int)a.steps; g.F = F; g.H = H; g.out_act = a.out_act; g.BT = BT;
  dim3 grid((unsigned)((a.batch + BT - 1) / BT), (unsigned)a.num_dirs);
  const size_t smem = smem_bytes(BT);
  A3GC_CUDA_TRY(cudaFuncSetAttribute(gru_train_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gru_train_bwd_kernel<<<grid, kThreads, smem, stream>>>(bd[0], bd[1], g);
  A3GC_LAUNCH_CHECK("gru_train_bwd_kernel");
  return A3GC_OK;
}

int simt_train_forward(const LayerArgs& a, const a3gc_tape& tape, const float* hmask, void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (a.variant == A3GC_VARIANT_GGRU) return simt_gru_train_forward(a, tape, ws, ws_bytes, stream);
  const int F = a.f_in, H = a.hidden;
  const size_t need = simt_train_workspace_bytes(a.variant, F, H, a.num_dirs);
  if (ws_bytes < need || ws == nullptr) {
    set_error("a3gc_layer_train_forward: workspace too small (%zu < %zu bytes)", ws_bytes, need);
    return A3GC_ERR_WORKSPACE;
  }
  const int smem_max = max_optin_smem();
  if (smem_max <= 0) { set_error("no CUDA device"); return A3GC_ERR_NO_DEVICE; }
  const bool att = a.variant == A3GC_VARIANT_A3GC || a.variant == A3GC_VARIANT_AGC;
  const int KX = F > H ? F : H;
  auto smem_bytes = [&](int bt) -> size_t {
    return ((size_t)3 * bt * H * 16 + (size_t)bt * KX * 16 + (size_t)2 * bt * H + (size_t)bt * 16 + 1024) * sizeof(float);
  };
  int BT = 8;
  while (BT > 1 && smem_bytes(BT) > (size_t)smem_max) --BT;
  int sms = 148;
  { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  while (BT > 1 && ((a.batch + BT - 1) / BT) * a.num_dirs < sms) --BT;
  if (smem_bytes(BT) > (size_t)smem_max) {
    set_error("training forward: hidden=%d f_in=%d needs %zu bytes of shared memory (> %d)", H, F, smem_bytes(1), smem_max);
    return A3GC_ERR_UNSUPPORTED;
  }
  char* base = static_cast<char*>(ws);
  const size_t per = need / a.num_dirs;
  DirPtrs dp[2] = {};
  LstmPacked pk[2];
  for (int d = 0; d < a.num_dirs; ++d) {
    dp[d].h0 = a.h0[d]; dp[d].c0 = a.c0[d]; dp[d].hT = a.hT[d]; dp[d].cT = a.cT[d]; dp[d].reverse = a.reverse[d];
    pk[d] = carve_lstm(reinterpret_cast<float*>(base + d * per), F, H);
    pack_lstm_kernel<<<64, 256, 0, stream>>>(a.cells[d], pk[d], F, H, a.variant);
    A3GC_LAUNCH_CHECK("pack_lstm_kernel");
  }
  if (a.num_dirs == 1) pk[1] = pk[0];
  TrainGeom tg;
  LayerGeom& g = tg.g;
  g.x = a.x; g.sxb = a.x_stride_b; g.sxt = a.x_stride_t;
  g.y = a.y; g.syb = a.y_stride_b; g.syt = a.y_stride_t; g.yld = a.y_ld;
  g.B = (int)a.batch; g.T = (int)a.steps; g.F = F; g.H = H; g.out_act = a.out_act; g.BT = BT;
  tg.tape = tape; tg.hmask = hmask;
  dim3 grid((unsigned)((a.batch + BT - 1) / BT), (unsigned)a.num_dirs);
  const size_t smem = smem_bytes(BT);
  if (att) {
    A3GC_CUDA_TRY(cudaFuncSetAttribute(lstm_train_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lstm_train_fwd_kernel<true><<<grid, kThreads, smem, stream>>>(pk[0], pk[1], dp[0], dp[1], tg);
  } else {
    A3GC_CUDA_TRY(cudaFuncSetAttribute(lstm_train_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lstm_train_fwd_kernel<false><<<grid, kThreads, smem, stream>>>(pk[0], pk[1], dp[0], dp[1], tg);
  }
  A3GC_LAUNCH_CHECK("lstm_train_fwd_kernel");
  return A3GC_OK;
}

int simt_train_backward(const TrainBwdArgs& a, void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (a.variant == A3GC_VARIANT_GGRU) return simt_gru_train_backward(a, ws, ws_bytes, stream);
  const int F = a.f_in, H = a.hidden;
  const size_t need = simt_train_workspace_bytes(a.variant, F, H, a.num_dirs);
  if (ws_bytes < need || ws == nullptr) {
    set_error("a3gc_layer_backward: workspace too small (%zu < %zu bytes)", ws_bytes, need);
    return A3GC_ERR_WORKSPACE;
  }
  const int smem_max = max_optin_smem();
  if (smem_max <= 0) { set_error("no CUDA device"); return A3GC_ERR_NO_DEVICE; }
  const bool att = a.variant == A3GC_VARIANT_A3GC || a.variant == A3GC_VARIANT_AGC;
  auto smem_bytes = [&](int bt) -> size_t {
    return ((size_t)8 * bt * H * 16 + (size_t)2 * bt * H + (size_t)2 * bt * 16 + 1024) * sizeof(float);
  };
  int sms = 148;
  { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  // blocked kernel (H in {64, 128, 256}): BT * H = 512, halved once when that would leave more than half of the SMs idle
  const char* blk_env = getenv("A3GC_BWD_BLK");
  const bool blk = (H == 64 || H == 128 || H == 256) && !(blk_env != nullptr && blk_env[0] == '0');
  auto blk_smem_bytes = [&](int bt) -> size_t {
    return ((size_t)6 * bt * H * 16 + (size_t)2 * bt * H + (size_t)2 * bt * 16 + 256 + 1024) * sizeof(float);
  };
  int BT = 8;
  if (blk) {
    BT = 512 / H;
    if (((a.batch + BT - 1) / BT) * a.num_dirs * 2 <= sms && BT >= 4) BT /= 2;
    if (const char* e = getenv("A3GC_BWD_BT")) { const int v = atoi(e); if (v == 512 / H || (v == 256 / H && v >= 2)) BT = v; }
    if (blk_smem_bytes(BT) > (size_t)smem_max) {
      set_error("training backward: hidden=%d needs %zu bytes of shared memory (> %d)", H, blk_smem_bytes(BT), smem_max);
      return A3GC_ERR_UNSUPPORTED;
    }
  } else {
    while (BT > 1 && smem_bytes(BT) > (size_t)smem_max) --BT;
    while (BT > 1 && ((a.batch + BT - 1) / BT) * a.num_dirs < sms) --BT;
    if (smem_bytes(BT) > (size_t)smem_max) {
      set_error("training backward: hidden=%d needs %zu bytes of shared memory (> %d)", H, smem_bytes(1), smem_max);
      return A3GC_ERR_UNSUPPORTED;
    }
  }
  char* base = static_cast<char*>(ws);
  const size_t per = need / a.num_dirs;
  BwdDir bd[2] = {};
  for (int d = 0; d < a.num_dirs; ++d) {
    float* pt = reinterpret_cast<float*>(base + d * per + align_up(lstm_packed_floats(F, H) * sizeof(float), 256));
    pack_pt_kernel<<<1, 256, 0, stream>>>(a.cells[d], pt, a.variant);
    A3GC_LAUNCH_CHECK("pack_pt_kernel");
    for (int q = 0; q < 4; ++q) bd[d].Wg[q] = a.cells[d].gcn_kernel[q];
    bd[d].Wh = a.cells[d].attention_wh; bd[d].Wq = a.cells[d].attention_wq; bd[d].Wa = a.cells[d].attention_w;
    bd[d].u = a.cells[d].attention_u; bd[d].PT = pt;
    bd[d].wfrag = reinterpret_cast<const uint4*>(pt + 4 * 256);
    bd[d].c0 = a.c0[d]; bd[d].dhT = a.dhT[d]; bd[d].dcT = a.dcT[d]; bd[d].dh0 = a.dh0[d]; bd[d].dc0 = a.dc0[d];
    bd[d].reverse = a.reverse[d];
  }
  if (a.num_dirs == 1) bd[1] = bd[0];
  BwdGeom g;
  g.dy = a.dy; g.syb = a.dy_stride_b; g.syt = a.dy_stride_t; g.yld = a.dy_ld;
  g.tape = a.tape; g.gr = a.grads; g.hmask = a.hmask;
  g.B = (int)a.batch; g.T = (int)a.steps; g.F = F; g.H = H; g.out_act = a.out_act; g.BT = BT;
  {
    // tile shapes the tensor-core phases are instantiated for: (H / 128, 2 BT) in {(2, 4), (1, 8), (1, 4)} or H = 64 with BT in {4, 8}
    const int nt = H >= 128 ? 2 * BT : BT;
    const char* e = getenv("A3GC_BWD_MMA");
    const int mode = e != nullptr ? atoi(e) : 4;
    g.tcore = (blk && ((H == 256 && nt == 4) || (H <= 128 && (nt == 4 || nt == 8)))) ? (mode < 0 || mode > 4 ? 4 : mode) : 0;
  }
  dim3 grid((unsigned)((a.batch + BT - 1) / BT), (unsigned)a.num_dirs);
  if (blk) {
    const size_t bsmem = blk_smem_bytes(BT);
    auto launch = [&](auto kern) -> int {
      A3GC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsmem));
      kern<<<grid, kThreads, bsmem, stream>>>(bd[0], bd[1], g);
      return A3GC_OK;
    };
    g.trace = nullptr;
    const char* tre = getenv("A3GC_BWD_TRACE");
    if (tre != nullptr && tre[0] == '1') A3GC_CUDA_TRY(cudaMalloc(&g.trace, 8 * sizeof(long long)));
    int rc;
    if (g.tcore >= 3) {
      const int nblk = att ? 7 : 4, ntile = nblk * (H / 16) * (H / 16);
      for (int d = 0; d < a.num_dirs; ++d) {
        pack_bwd_frag_kernel<<<(ntile + 7) / 8, 256, 0, stream>>>(bd[d], const_cast<uint4*>(bd[d].wfrag), F, H, nblk, g.tcore == 3);
        A3GC_LAUNCH_CHECK("pack_bwd_frag_kernel");
      }
    }
    if (att) rc = g.tcore == 4 ? launch(lstm_train_bwd_blk_kernel<true, 4>) : g.tcore == 3 ? launch(lstm_train_bwd_blk_kernel<true, 3>)
                : g.tcore == 2 ? launch(lstm_train_bwd_blk_kernel<true, 2>) : g.tcore == 1 ? launch(lstm_train_bwd_blk_kernel<true, 1>)
                                                                              : launch(lstm_train_bwd_blk_kernel<true, 0>);
    else rc = g.tcore == 4 ? launch(lstm_train_bwd_blk_kernel<false, 4>) : g.tcore == 3 ? launch(lstm_train_bwd_blk_kernel<false, 3>)
              : g.tcore == 2 ? launch(lstm_train_bwd_blk_kernel<false, 2>) : g.tcore == 1 ? launch(lstm_train_bwd_blk_kernel<false, 1>)
                                                                            : launch(lstm_train_bwd_blk_kernel<false, 0>);
    if (rc != A3GC_OK) return rc;
    if (g.trace != nullptr) {
      long long h[8];
      A3GC_CUDA_TRY(cudaStreamSynchronize(stream));
      A3GC_CUDA_TRY(cudaMemcpy(h, g.trace, sizeof(h), cudaMemcpyDeviceToHost));
      cudaFree(g.trace);
      fprintf(stderr, "[a3gc bwd trace] H=%d BT=%d T=%d tcore=%d cycles/step: A %lld  B %lld  C %lld  D %lld  D2in %lld  E+D2red %lld  F %lld  G %lld\n",
              H, BT, g.T, g.tcore, h[0] / g.T, h[1] / g.T, h[2] / g.T, h[3] / g.T, h[4] / g.T, h[5] / g.T, h[6] / g.T, h[7] / g.T);
    }
    A3GC_LAUNCH_CHECK("lstm_train_bwd_blk_kernel");
    return A3GC_OK;
  }
  const size_t smem = smem_bytes(BT);
  if (att) {
    A3GC_CUDA_TRY(cudaFuncSetAttribute(lstm_train_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lstm_train_bwd_kernel<true><<<grid, kThreads, smem, stream>>>(bd[0], bd[1], g);
  } else {
    A3GC_CUDA_TRY(cudaFuncSetAttribute(lstm_train_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lstm_train_bwd_kernel<false><<<grid, kThreads, smem, stream>>>(bd[0], bd[1], g);
  }
  A3GC_LAUNCH_CHECK("lstm_train_bwd_kernel");
  return A3GC_OK;
}

int simt_gc_forward(const a3gc_gc_params* p, const float* x, float* y, int64_t frames, int f_in,
                    int f_out, int act, cudaStream_t stream) {
  if (frames == 0) return A3GC_OK;
  {
    int handled = 0;
    int rc = gc_forward_fast(p, x, y, frames, f_in, f_out, act, stream, &handled);
    if (rc != A3GC_OK || handled) return rc;
  }
  const int smem_max = max_optin_smem();
  if (smem_max <= 0) { set_error("no CUDA device"); return A3GC_ERR_NO_DEVICE; }
  int FR = 8;
  auto bytes = [&](int fr) { return (256 + (size_t)2 * fr * kNodes * f_in) * sizeof(float); };
  while (FR > 1 && bytes(FR) > (size_t)96 * 1024) --FR;
  if (bytes(FR) > (size_t)smem_max) { set_error("gc_forward: f_in=%d too large", f_in); return A3GC_ERR_UNSUPPORTED; }
  int sms = 148;
  { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  int64_t blocks = (frames + FR - 1) / FR;
  if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
  A3GC_CUDA_TRY(cudaFuncSetAttribute(gc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes(FR)));
  gc_kernel<<<(unsigned)blocks, kThreads, bytes(FR), stream>>>(*p, x, y, frames, f_in, f_out, act, FR);
  A3GC_LAUNCH_CHECK("gc_kernel");
  return A3GC_OK;
}

int simt_prepare_input(const float* acc, const float* ori, const float* acc_mean, const float* acc_std,
                       const float* ori_mean, const float* ori_std, float* x, int64_t frames, int ld_x,
                       cudaStream_t stream) {
  if (frames == 0) return A3GC_OK;
  int64_t total = frames * kNodes * 12;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  prepare_input_kernel<<<(unsigned)blocks, 256, 0, stream>>>(acc, ori, acc_mean, acc_std, ori_mean, ori_std, x, frames, ld_x);
  A3GC_LAUNCH_CHECK("prepare_input_kernel");
  return A3GC_OK;
}

int simt_concat_stage_input(const float* x, const float* pos, float* dst, int64_t frames, cudaStream_t stream) {
  if (frames == 0) return A3GC_OK;
  int64_t rows = frames * kNodes;
  int64_t blocks = (rows * 15 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  concat_stage_input_kernel<<<(unsigned)blocks, 256, 0, stream>>>(x, pos, dst, rows);
  A3GC_LAUNCH_CHECK("concat_stage_input_kernel");
  return A3GC_OK;
}

int simt_reduced_to_full_local(const float* pose, float* out, int64_t frames, int rotsize, cudaStream_t stream) {
  if (frames == 0) return A3GC_OK;
  int64_t blocks = (frames * 24 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  reduced_to_full_local_kernel<<<(unsigned)blocks, 256, 0, stream>>>(pose, out, frames, rotsize);
  A3GC_LAUNCH_CHECK("reduced_to_full_local_kernel");
  return A3GC_OK;
}

}  // namespace a3gc

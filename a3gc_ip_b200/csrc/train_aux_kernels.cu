// Helpers of the training step that are not part of the recurrent chain: operand preparation for the hoisted
// weight / input gradient GEMMs (train_a3gc_tp.py:74-84 runs them as autograd mm/bmm nodes in fp32).
//
// The hoisted GEMMs run on the tensor cores as three TF32 passes (hi*hi + lo*hi + hi*lo): every fp32 operand is cut
// into a head that is exactly representable in TF32 (cvt.rna, 11 significand bits) and the fp32 remainder, so the
// library's TF32 conversion of the head is exact and the dropped lo*lo term is 2^-22 relative.
//
// "Mixed" form (round 2): only the hi*hi product needs TF32; the two correction products carry 2^-11 of the result, so
// their operands may be rounded to bf16 (2^-9 of a 2^-11 term) and run at the bf16 tensor rate -- measured 4-5x the
// TF32 rate of the library on these shapes (tests/diag_dw_gemm.py).  The builders below therefore emit hi (fp32,
// TF32-exact), bf16(hi) and bf16(lo) in one pass, into row-strided buffers so that x and h_prev land side by side
// in one S = [x | h_prev] operand (one GEMM instead of two per weight-gradient block).
#include "common.cuh"

namespace a3gc {
namespace {

__device__ __forceinline__ float tf32_head(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

__global__ void __launch_bounds__(256) split_tf32_kernel(const float4* __restrict__ x, float4* __restrict__ hi,
                                                         float4* __restrict__ lo, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(x + i);
    float4 h, l;
    h.x = tf32_head(v.x); h.y = tf32_head(v.y); h.z = tf32_head(v.z); h.w = tf32_head(v.w);
    l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
    hi[i] = h; lo[i] = l;
  }
}

__global__ void split_tf32_tail_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo,
                                       int64_t first, int64_t n) {
  const int64_t i = first + threadIdx.x;
  if (i < n) { const float h = tf32_head(x[i]); hi[i] = h; lo[i] = x[i] - h; }
}

// S_h operand of dW_h = dzm^T S_h: the state that entered step t of one direction, i.e. h' of the previous step of
// that direction (h0 at its first step) times the recurrent-dropout mask of step t (net_aagc.py:181-182), split.
//   hp [B][T][15][H] (tape), h0 [B][15][H] or nullptr (zeros), mask [B][T][15][H] or nullptr
__global__ void __launch_bounds__(256) hprev_split_kernel(const float4* __restrict__ hp, const float4* __restrict__ h0,
                                                          const float4* __restrict__ mask, float4* __restrict__ hi,
                                                          float4* __restrict__ lo, int B, int T, int row4, int reverse) {
  const int64_t n4 = (int64_t)B * T * row4;                     // row4 = 15*H/4 float4 per (b, t)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t bt = i / row4;
    const int r = (int)(i - bt * row4);
    const int b = (int)(bt / T), t = (int)(bt - (int64_t)b * T);
    const int tp = reverse ? t + 1 : t - 1;
    float4 v;
    if (tp < 0 || tp >= T) v = h0 != nullptr ? __ldg(h0 + (int64_t)b * row4 + r) : make_float4(0.f, 0.f, 0.f, 0.f);
    else v = __ldg(hp + ((int64_t)b * T + tp) * row4 + r);
    if (mask != nullptr) { const float4 m = __ldg(mask + i); v.x *= m.x; v.y *= m.y; v.z *= m.z; v.w *= m.w; }
    float4 h, l;
    h.x = tf32_head(v.x); h.y = tf32_head(v.y); h.z = tf32_head(v.z); h.w = tf32_head(v.w);
    l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
    hi[i] = h; lo[i] = l;
  }
}

__device__ __forceinline__ uint2 bf16x4(float a, float b, float c, float d) {
  uint2 r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r.x) : "f"(b), "f"(a));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r.y) : "f"(d), "f"(c));
  return r;
}

__device__ __forceinline__ void emit_mixed(const float4& v, float4* hi, uint2* hi16, uint2* lo16, int64_t o4) {
  float4 h;
  h.x = tf32_head(v.x); h.y = tf32_head(v.y); h.z = tf32_head(v.z); h.w = tf32_head(v.w);
  hi[o4] = h;
  hi16[o4] = bf16x4(h.x, h.y, h.z, h.w);
  lo16[o4] = bf16x4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
}

// x [rows][cols] -> hi / hi16 / lo16 [rows][ld] at column col0 (all in units of 4 elements)
__global__ void __launch_bounds__(256) split_mixed_kernel(const float4* __restrict__ x, float4* __restrict__ hi, uint2* __restrict__ hi16,
                                                          uint2* __restrict__ lo16, int64_t rows, int cols4, int64_t ld4, int64_t col04) {
  const int64_t n4 = rows * cols4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / cols4;
    emit_mixed(__ldg(x + i), hi, hi16, lo16, row * ld4 + col04 + (i - row * cols4));
  }
}

// hprev_split_kernel in mixed form; output rows are (b, t, node), H/4 units each, at column col0 of a [rows][ld] buffer
__global__ void __launch_bounds__(256) hprev_split_mixed_kernel(const float4* __restrict__ hp, const float4* __restrict__ h0,
                                                                const float4* __restrict__ mask, float4* __restrict__ hi,
                                                                uint2* __restrict__ hi16, uint2* __restrict__ lo16, int B, int T,
                                                                int h4, int64_t ld4, int64_t col04, int reverse) {
  const int row4 = kNodes * h4;
  const int64_t n4 = (int64_t)B * T * row4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t bt = i / row4;
    const int r = (int)(i - bt * row4);
    const int b = (int)(bt / T), t = (int)(bt - (int64_t)b * T);
    const int tp = reverse ? t + 1 : t - 1;
    float4 v;
    if (tp < 0 || tp >= T) v = h0 != nullptr ? __ldg(h0 + (int64_t)b * row4 + r) : make_float4(0.f, 0.f, 0.f, 0.f);
    else v = __ldg(hp + ((int64_t)b * T + tp) * row4 + r);
    if (mask != nullptr) { const float4 m = __ldg(mask + i); v.x *= m.x; v.y *= m.y; v.z *= m.z; v.w *= m.w; }
    const int64_t row = i / h4;                                  // (b, t, node)
    emit_mixed(v, hi, hi16, lo16, row * ld4 + col04 + (i - row * h4));
  }
}

// Adjacency gradients of one direction: dP_g[m][n] = sum over records r and units j of dz[r][g][j][m] * u[r][g][j][n]
// (the autograd bmm + sum of training.py; dz is what the backward chain left in place of the gates, u the pre-mix
// accumulators of the tape, both [records][4][H][16]).  One pass over the two arrays at HBM speed instead of four batched
// 16 x H x 16 GEMMs: block (x, g) walks records x, x + gridDim.x, ... of gate g; the two H x 16 tiles of a record (16 KB each
// at H = 256, contiguous) are staged in shared memory by cp.async, double buffered -- a first version that read them
// straight into registers had 16 KB per SM in flight and reached a third of the HBM rate; thread (js, mb, nb) owns the
// 4 x 4 block (m = 4 mb.., n = 4 nb..) for the units j = js (mod 16); the 16 unit slices meet in shared memory and the
// blocks' partial results are summed in a fixed order by adjacency_grad_reduce_kernel (deterministic).
__global__ void __launch_bounds__(256) adjacency_grad_kernel(const float4* __restrict__ dz, const float4* __restrict__ u,
                                                             float* __restrict__ partial, int64_t records, int H) {
  extern __shared__ __align__(16) float4 stage[];         // [2 buffers][dz tile | u tile], H * 4 float4 per tile
  const int g = blockIdx.y, tile4 = H * 4;
  const int js = threadIdx.x >> 4, mb = (threadIdx.x >> 2) & 3, nb = threadIdx.x & 3;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  auto fetch = [&](int64_t r, int buf) {
    const float4* dr = dz + (r * 4 + g) * tile4;
    const float4* ur = u + (r * 4 + g) * tile4;
    float4* sd = stage + (size_t)buf * 2 * tile4;
    for (int i = threadIdx.x; i < tile4; i += blockDim.x) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(sd + i)), "l"(dr + i));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(sd + tile4 + i)), "l"(ur + i));
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int buf = 0;
  if ((int64_t)blockIdx.x < records) fetch(blockIdx.x, 0);
  for (int64_t r = blockIdx.x; r < records; r += gridDim.x) {
    const int64_t rn = r + gridDim.x;
    if (rn < records) {
      fetch(rn, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const float4* sd = stage + (size_t)buf * 2 * tile4;
    const float4* su = sd + tile4;
#pragma unroll 4
    for (int j = js; j < H; j += 16) {
      const float4 a = sd[j * 4 + mb], b = su[j * 4 + nb];
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] = fmaf(av[x], bv[y], acc[x][y]);
    }
    __syncthreads();                                       // the buffer is refilled two iterations later
    buf ^= 1;
  }
  float* red = reinterpret_cast<float*>(stage);            // [16 slices][256]
#pragma unroll
  for (int x = 0; x < 4; ++x)
#pragma unroll
    for (int y = 0; y < 4; ++y) red[js * 256 + (4 * mb + x) * 16 + 4 * nb + y] = acc[x][y];
  __syncthreads();
  float sum = 0.f;
  for (int sl = 0; sl < 16; ++sl) sum += red[sl * 256 + threadIdx.x];
  partial[((size_t)blockIdx.x * 4 + g) * 256 + threadIdx.x] = sum;
}

__global__ void adjacency_grad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dP, int nblocks) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= 1024) return;
  float sum = 0.f;
  for (int b = 0; b < nblocks; ++b) sum += partial[(size_t)b * 1024 + o];
  dP[o] = sum;
}

}  // namespace

int train_split_tf32(const float* x, float* hi, float* lo, int64_t n, cudaStream_t stream) {
  if (n == 0) return A3GC_OK;
  const int64_t n4 = n / 4;
  if (n4 > 0) {
    const int64_t blocks = (n4 + 255) / 256;
    split_tf32_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, stream>>>(
        reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(hi), reinterpret_cast<float4*>(lo), n4);
    A3GC_LAUNCH_CHECK("split_tf32_kernel");
  }
  if (n4 * 4 < n) {
    split_tf32_tail_kernel<<<1, 32, 0, stream>>>(x, hi, lo, n4 * 4, n);
    A3GC_LAUNCH_CHECK("split_tf32_tail_kernel");
  }
  return A3GC_OK;
}

int train_hprev_split(const float* hp, const float* h0, const float* mask, float* hi, float* lo, int64_t batch,
                      int64_t steps, int hidden, int reverse, cudaStream_t stream) {
  if (batch == 0 || steps == 0) return A3GC_OK;
  const int row4 = kNodes * hidden / 4;
  const int64_t n4 = batch * steps * row4;
  const int64_t blocks = (n4 + 255) / 256;
  hprev_split_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, stream>>>(
      reinterpret_cast<const float4*>(hp), reinterpret_cast<const float4*>(h0), reinterpret_cast<const float4*>(mask),
      reinterpret_cast<float4*>(hi), reinterpret_cast<float4*>(lo), (int)batch, (int)steps, row4, reverse);
  A3GC_LAUNCH_CHECK("hprev_split_kernel");
  return A3GC_OK;
}

int train_split_mixed(const float* x, int64_t rows, int cols, float* hi, uint16_t* hi16, uint16_t* lo16, int64_t ld, int64_t col0,
                      cudaStream_t stream) {
  if (rows == 0 || cols == 0) return A3GC_OK;
  const int64_t n4 = rows * (cols / 4);
  const int64_t blocks = (n4 + 255) / 256;
  split_mixed_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, stream>>>(
      reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(hi), reinterpret_cast<uint2*>(hi16), reinterpret_cast<uint2*>(lo16),
      rows, cols / 4, ld / 4, col0 / 4);
  A3GC_LAUNCH_CHECK("split_mixed_kernel");
  return A3GC_OK;
}

int train_hprev_split_mixed(const float* hp, const float* h0, const float* mask, float* hi, uint16_t* hi16, uint16_t* lo16,
                            int64_t batch, int64_t steps, int hidden, int64_t ld, int64_t col0, int reverse, cudaStream_t stream) {
  if (batch == 0 || steps == 0) return A3GC_OK;
  const int64_t n4 = batch * steps * kNodes * (hidden / 4);
  const int64_t blocks = (n4 + 255) / 256;
  hprev_split_mixed_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, stream>>>(
      reinterpret_cast<const float4*>(hp), reinterpret_cast<const float4*>(h0), reinterpret_cast<const float4*>(mask),
      reinterpret_cast<float4*>(hi), reinterpret_cast<uint2*>(hi16), reinterpret_cast<uint2*>(lo16), (int)batch, (int)steps, hidden / 4,
      ld / 4, col0 / 4, reverse);
  A3GC_LAUNCH_CHECK("hprev_split_mixed_kernel");
  return A3GC_OK;
}

int train_adjacency_grad(const float* dz, const float* u, int64_t records, int hidden, float* partial, int nblocks, float* dP,
                         cudaStream_t stream) {
  size_t smem = (size_t)2 * 2 * hidden * 16 * sizeof(float);           // two buffers of (dz tile | u tile)
  if (smem < (size_t)16 * 256 * sizeof(float)) smem = (size_t)16 * 256 * sizeof(float);
  A3GC_CUDA_TRY(cudaFuncSetAttribute(adjacency_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  adjacency_grad_kernel<<<dim3((unsigned)nblocks, 4), 256, smem, stream>>>(reinterpret_cast<const float4*>(dz), reinterpret_cast<const float4*>(u), partial,
                                                        records, hidden);
  A3GC_LAUNCH_CHECK("adjacency_grad_kernel");
  adjacency_grad_reduce_kernel<<<4, 256, 0, stream>>>(partial, dP, nblocks);
  A3GC_LAUNCH_CHECK("adjacency_grad_reduce_kernel");
  return A3GC_OK;
}

}  // namespace a3gc

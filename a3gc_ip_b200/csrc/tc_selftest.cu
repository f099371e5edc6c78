// Self-test of the tcgen05 building blocks the tensor-core engine relies on: 1-D bulk copies onto an
// mbarrier, UMMA shared-memory / instruction descriptors for the no-swizzle K-major layout, the
// TMEM accumulator layout read back with tcgen05.ld.  D[128, N] = A[128, K] * B[N, K]^T in fp16/bf16
// with fp32 accumulation, one CTA.  Exposed as a3gc_tc_selftest (include/a3gc_b200.h) and checked
// against a CPU product in tests/test_gpu_tc.py.
#include "common.cuh"
#include "tc_ptx.cuh"
#include <cstring>

namespace a3gc {
namespace {

// operand image: [K/8 chunks][rows][8 elements] 16-bit
__global__ void __launch_bounds__(128, 1)
tc_selftest_kernel(const uint16_t* __restrict__ a_img, const uint16_t* __restrict__ b_img, float* __restrict__ d,
                   int K, int N, int bf16, int swap_lbo_sbo, int vec_a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_full, bar_mma;
  __shared__ uint32_t tmem_base_slot;
  const int warp = threadIdx.x / 32;
  uint8_t* sa = smem;                                  // [K/8][128][16 B]
  uint8_t* sb = smem + (size_t)K * 128 * 2;            // [K/8][N][16 B]
  // vec_a: the A image holds only 8 rows per K chunk ([K/8][8][16 B]) and the descriptor's 8-row-group stride (SBO) is 0, so
  // every one of the 16 row groups of the M = 128 tile aliases the same 8 rows: D row r = A row (r & 7).  The layer kernel
  // feeds its per-sequence vectors (node sums / q of the attention block) to the tensor core this way.
  const uint32_t a_rows = vec_a ? 8u : 128u;
  const uint32_t bytes_a = (uint32_t)K * a_rows * 2, bytes_b = (uint32_t)K * N * 2;

  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar_full, 1);
    ptx::mbar_init(&bar_mma, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) ptx::tmem_alloc(&tmem_base_slot, 256);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_base_slot;

  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(&bar_full, bytes_a + bytes_b);
    ptx::bulk_g2s(sa, a_img, bytes_a, &bar_full);
    ptx::bulk_g2s(sb, b_img, bytes_b, &bar_full);
    ptx::mbar_wait(&bar_full, 0);
    ptx::tc_fence_after();
    const uint32_t idesc = ptx::make_idesc_f16(128, N, bf16 != 0);
    for (int kk = 0; kk < K / 16; ++kk) {
      // K-chunk stride (LBO) = rows * 16 B; 8-row group stride (SBO) = 128 B
      uint32_t a_lbo = a_rows * 16, a_sbo = vec_a ? 0u : 128u, b_lbo = (uint32_t)N * 16, b_sbo = 128;
      if (swap_lbo_sbo) { uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t; t = b_lbo; b_lbo = b_sbo; b_sbo = t; }
      const uint64_t ad = ptx::make_smem_desc(ptx::smem_u32(sa) + (uint32_t)kk * 2 * a_rows * 16, a_lbo, a_sbo);
      const uint64_t bd = ptx::make_smem_desc(ptx::smem_u32(sb) + (uint32_t)kk * 2 * N * 16, b_lbo, b_sbo);
      ptx::umma_f16(tmem, ad, bd, idesc, kk > 0 ? 1u : 0u);
    }
    ptx::umma_commit(&bar_mma);
  }
  __syncwarp();
  ptx::mbar_wait(&bar_mma, 0);
  ptx::tc_fence_after();
  // each warp reads its 32 TMEM lanes (= rows 32*warp .. +31)
  const int row = warp * 32 + (threadIdx.x & 31);
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    ptx::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
    for (int i = 0; i < 32; ++i) if (c0 + i < N) d[(size_t)row * N + c0 + i] = v[i];
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 256);
}


// Micro-benchmark of the UMMA issue rate for a given shared-memory operand layout: `iters` back-to-back
// tcgen05.mma (M=128, N=n, K=16, fp16) cycling over `nk` K-steps of one resident tile; reports clock64 cycles per
// MMA of CTA 0.  Operand contents are zeros -- only the addressing pattern matters.
__global__ void __launch_bounds__(128, 1)
tc_mma_bench_kernel(int n, int layout_type, int a_lbo, int a_sbo, int b_lbo, int b_sbo, int a_kstep, int b_kstep,
                    int nk, int iters, float* __restrict__ cycles_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_mma;
  __shared__ uint32_t tmem_base_slot;
  const int warp = threadIdx.x / 32;
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { ptx::mbar_init(&bar_mma, 1); ptx::fence_mbar_init(); }
  if (warp == 0) ptx::tmem_alloc(&tmem_base_slot, 256);
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_base_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::make_idesc_f16(128, n, false);
    const uint32_t sa = ptx::smem_u32(smem), sb = sa + 64 * 1024;
    const uint64_t lt = (uint64_t)layout_type << 61;
    // descriptors of the (up to) 4 K-steps are built once: the issue loop is the MMA instructions alone
    uint64_t ad[4], bd[4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      ad[kk] = ptx::make_smem_desc(sa + (uint32_t)(kk % nk) * a_kstep, a_lbo, a_sbo) | lt;
      bd[kk] = ptx::make_smem_desc(sb + (uint32_t)(kk % nk) * b_kstep, b_lbo, b_sbo) | lt;
    }
    ptx::umma_f16(tmem, ad[0], bd[0], idesc, 0u);
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += 4) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) ptx::umma_f16(tmem, ad[kk], bd[kk], idesc, 1u);
    }
    ptx::umma_commit(&bar_mma);
    ptx::mbar_wait(&bar_mma, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) cycles_out[0] = (float)(t1 - t0) / (float)iters;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 256);
}


// Micro-benchmark of the L2 -> shared-memory bulk-copy stream: every CTA pulls `iters` chunks of `chunk` bytes from a
// `span`-byte (L2-resident) buffer through a `depth`-slot ring, no compute.  `nprod` producer threads (one per warp) each
// own every nprod-th slot, so the per-copy issue / wait overhead of a single thread can be told apart from the link rate.
// mcast = 1 (nprod must be 1): the CTAs of a 2-CTA cluster each fetch HALF of every chunk and multicast it to both.
// Reports bytes per cycle per CTA (CTA 0's clock, slowest producer).
__global__ void __launch_bounds__(256, 1)
tc_stream_bench_kernel(const uint8_t* __restrict__ src, size_t span, int chunk, int depth, int iters, int mcast, int nprod, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full[8], empty[8];
  __shared__ long long tend[8];
  const uint32_t rank = mcast ? ptx::cluster_ctarank() : 0u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], mcast ? 2 : 1); }
    ptx::fence_mbar_init();
  }
  __syncthreads();
  if (mcast) ptx::cluster_sync_all();
  const int w = threadIdx.x >> 5;
  const long long t0 = clock64();
  if ((threadIdx.x & 31) == 0 && w < nprod) {
    const size_t base = ((size_t)(blockIdx.x >> (mcast ? 1 : 0)) * 1315423911ull) % (span - (size_t)chunk * 64);
    const int my_depth = depth / nprod;                     // this producer's private slots: w, w + nprod, ...
    const int my_iters = iters / nprod;
    int st = 0, ph = 0, cs = 0, cp = 0;
    for (int it = 0; it < my_iters; ++it) {
      const int slot = st * nprod + w;
      ptx::mbar_wait(&empty[slot], ph ^ 1);
      ptx::mbar_arrive_expect_tx(&full[slot], (uint32_t)chunk);
      const uint8_t* g = src + ((base + (size_t)((it * nprod + w) & 63) * chunk) & ~(size_t)127);
      if (mcast) {
        const uint32_t half = (uint32_t)chunk / 2;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                     ::"r"(ptx::smem_u32(smem + (size_t)slot * chunk + rank * half)), "l"(g + rank * half), "r"(half),
                       "r"(ptx::smem_u32(&full[slot])), "h"((uint16_t)3) : "memory");
      } else {
        ptx::bulk_g2s(smem + (size_t)slot * chunk, g, (uint32_t)chunk, &full[slot]);
      }
      if (++st == my_depth) { st = 0; ph ^= 1; }
      if (it >= my_depth - 1) {                              // consumer side: wait for the oldest slot and release it
        const int cslot = cs * nprod + w;
        ptx::mbar_wait(&full[cslot], cp);
        if (mcast) { ptx::mbar_arrive_remote(&empty[cslot], 0); ptx::mbar_arrive_remote(&empty[cslot], 1); } else ptx::mbar_arrive(&empty[cslot]);
        if (++cs == my_depth) { cs = 0; cp ^= 1; }
      }
    }
    for (int r = 0; r < my_depth - 1; ++r) {
      ptx::mbar_wait(&full[cs * nprod + w], cp);
      if (++cs == my_depth) { cs = 0; cp ^= 1; }
    }
    tend[w] = clock64();
  }
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    long long t1 = t0;
    for (int i = 0; i < nprod; ++i) t1 = tend[i] > t1 ? tend[i] : t1;
    out[0] = (float)((double)chunk * (iters / nprod) * nprod / (double)(t1 - t0));
  }
  if (mcast) ptx::cluster_sync_all();
}

}  // namespace
}  // namespace a3gc

extern "C" int a3gc_tc_selftest(const void* a_img, const void* b_img, float* d, int k, int n, int flags, void* stream) {
  using namespace a3gc;
  if (!a_img || !b_img || !d || k <= 0 || k % 16 != 0 || n < 16 || n > 256 || n % 16 != 0) {
    set_error("a3gc_tc_selftest: need K %% 16 == 0 and 16 <= N <= 256, N %% 16 == 0");
    return A3GC_ERR_INVALID_ARG;
  }
  // always the same large allocation, so that a wrong descriptor reads wrong data instead of faulting
  const size_t smem = 200 * 1024;
  if ((size_t)k * (128 + n) * 2 > smem) { set_error("a3gc_tc_selftest: K too large"); return A3GC_ERR_INVALID_ARG; }
  A3GC_CUDA_TRY(cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_selftest_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint16_t*>(a_img), static_cast<const uint16_t*>(b_img), d, k, n, flags & 1, (flags >> 1) & 1, (flags >> 2) & 1);
  A3GC_LAUNCH_CHECK("tc_selftest_kernel");
  return A3GC_OK;
}

// debug / tuning: see tc_mma_bench_kernel.  cycles_out is a device pointer to one float.
extern "C" int a3gc_tc_mma_bench(int n, int layout_type, int a_lbo, int a_sbo, int b_lbo, int b_sbo, int a_kstep, int b_kstep,
                                 int nk, int iters, int grid, float* cycles_out, void* stream) {
  using namespace a3gc;
  if (!cycles_out || n < 16 || n > 256 || n % 16 != 0 || nk < 1 || iters < 1 || grid < 1) {
    set_error("a3gc_tc_mma_bench: invalid argument");
    return A3GC_ERR_INVALID_ARG;
  }
  const size_t smem = 160 * 1024;
  A3GC_CUDA_TRY(cudaFuncSetAttribute(tc_mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_mma_bench_kernel<<<grid, 128, smem, static_cast<cudaStream_t>(stream)>>>(n, layout_type, a_lbo, a_sbo, b_lbo, b_sbo, a_kstep, b_kstep,
                                                                              nk, iters, cycles_out);
  A3GC_LAUNCH_CHECK("tc_mma_bench_kernel");
  return A3GC_OK;
}

// tuning aid: see tc_stream_bench_kernel.  src: device buffer of `span` bytes; out: device pointer to one float (B/cycle/CTA)
extern "C" int a3gc_tc_stream_bench(const void* src, size_t span, int chunk, int depth, int iters, int grid, int mcast, int nprod, float* out, void* stream) {
  using namespace a3gc;
  if (!src || !out || chunk < 1024 || chunk % 256 || depth < 2 || depth > 8 || iters < depth || grid < 1 || span < (size_t)chunk * 128 ||
      (size_t)chunk * depth > 200 * 1024 || (mcast && (grid % 2 || nprod != 1)) || nprod < 1 || nprod > 8 || depth % nprod) {
    set_error("a3gc_tc_stream_bench: invalid argument");
    return A3GC_ERR_INVALID_ARG;
  }
  const size_t smem = (size_t)chunk * depth;
  A3GC_CUDA_TRY(cudaFuncSetAttribute(tc_stream_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3(256, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = mcast ? 2u : 1u; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  A3GC_CUDA_TRY(cudaLaunchKernelEx(&cfg, tc_stream_bench_kernel, static_cast<const uint8_t*>(src), span, chunk, depth, iters, mcast, nprod, out));
  A3GC_LAUNCH_CHECK("tc_stream_bench_kernel");
  return A3GC_OK;
}

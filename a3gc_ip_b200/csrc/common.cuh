// Shared declarations of the a3gc_b200 library (internal; the public surface is include/a3gc_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include "../../include/a3gc_b200.h"

namespace a3gc {

constexpr int kNodes = A3GC_NODES;   // 15 graph nodes
constexpr int kNodesPad = 16;        // node dimension padded to 16 (index 15 is always zero)

// ---- error plumbing (thread-local message, no exceptions across the ABI) -------------
void set_error(const char* fmt, ...);
int64_t& launch_counter();

inline int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  return A3GC_ERR_CUDA;
}

#define A3GC_CUDA_TRY(expr)                                  \
  do {                                                       \
    cudaError_t _e = (expr);                                 \
    if (_e != cudaSuccess) return ::a3gc::cuda_fail(_e, #expr); \
  } while (0)

// check the launch that was just issued
#define A3GC_LAUNCH_CHECK(name)                              \
  do {                                                       \
    ::a3gc::launch_counter() += 1;                           \
    cudaError_t _e = cudaGetLastError();                     \
    if (_e != cudaSuccess) return ::a3gc::cuda_fail(_e, "launch " name); \
  } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }
// tanh via exp: |err| ~ 1e-7 relative, well inside the 1e-4 parity budget; saturates cleanly.
__device__ __forceinline__ float tanhf_(float x) {
  float e = __expf(-2.0f * fabsf(x));
  float r = (1.0f - e) / (1.0f + e);
  return copysignf(r, x);
}
// fast variants for the tensor-core epilogue: one MUFU.EX2 + one MUFU.RCP each, flush-to-zero forms (no denormal
// fix-up code around the MUFU), |err| ~ 2e-7 absolute
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kExpCap = 1.0995116e12f;   // 2^40: products of three capped terms stay finite in fp32
// 1 + e^-x and e^-2x, capped from above (the cap only matters where the activation is saturated to < 1e-12)
__device__ __forceinline__ float one_plus_exp_neg(float x) { return 1.0f + fminf(ex2_ftz(x * -kLog2e), kExpCap); }
__device__ __forceinline__ float exp_neg2(float x) { return fminf(ex2_ftz(x * (-2.0f * kLog2e)), kExpCap); }
__device__ __forceinline__ float fast_sigmoid(float x) { return rcp_ftz(one_plus_exp_neg(x)); }
__device__ __forceinline__ float fast_tanh(float x) { return fmaf(2.0f, rcp_ftz(1.0f + exp_neg2(x)), -1.0f); }
// bf16 path only (stated bound 5e-3): single-MUFU tanh (relative error ~2^-11) and the sigmoid derived from it
__device__ __forceinline__ float tanh_approx(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoid_approx(float x) { return fmaf(0.5f, tanh_approx(0.5f * x), 0.5f); }
__device__ __forceinline__ float apply_act(float v, int act) {
  return act == A3GC_ACT_TANH ? tanhf_(v) : (act == A3GC_ACT_RELU ? fmaxf(v, 0.0f) : v);
}

// ---- SIMT (fp32 CUDA-core) engine: simt_kernels.cu -------------------------------------
struct LayerArgs {
  int variant;
  int num_dirs;
  const a3gc_cell_params* cells;   // host array [num_dirs]
  int reverse[2];
  const float* x;
  int64_t x_stride_b, x_stride_t;
  const float* h0[2];
  const float* c0[2];
  float* y;
  int64_t y_stride_b, y_stride_t, y_ld;
  float* hT[2];
  float* cT[2];
  int64_t batch, steps;
  int f_in, hidden, out_act;
  int precision;
  // tensor-core engine only: operand images handed between kernels instead of fp32 activations
  const uint16_t* x_img;   // if non-null: the layer input already packed as [tiles][T][F/16][NP][2][128][8] (x may be null)
  uint16_t* y_img;         // if non-null: also write act(h') into the next layer's input image ...
  int y_img_f;             // ... which has this many features (direction d writes columns d*H .. d*H+H-1)
  // training mode of the tensor-core engine (LSTM family, fp32 precision): keep the tape, apply the recurrent-dropout mask
  const void* packed;      // non-null: weights already packed by tc_pack_weights (the pack kernels are skipped)
  const a3gc_tape* tape;   // null = inference
  const float* hmask;      // [D][B][T][15][H] or null
};

size_t simt_layer_workspace_bytes(int variant, int f_in, int hidden, int num_dirs);
int simt_layer_forward(const LayerArgs& a, void* ws, size_t ws_bytes, cudaStream_t stream);
int simt_gc_forward(const a3gc_gc_params* p, const float* x, float* y, int64_t frames, int f_in,
                    int f_out, int act, cudaStream_t stream);
// gc_kernels.cu: HBM-bound implementations of the AAGC graph convolution for the shapes the nets use
int gc_forward_fast(const a3gc_gc_params* p, const float* x, float* y, int64_t frames, int f_in, int f_out, int act,
                    cudaStream_t stream, int* handled);
// the raw IMU frame as the input of a net's linear_in: prepare_input (evaluate_a3gc_tp.py:64-94) and the stage
// concatenation cat(x, pos) (:168, :170) fused into the load
struct GcRawInput {
  const float* acc;        // [frames, 18]  6 IMUs x 3
  const float* ori;        // [frames, 54]  6 IMUs x 9
  const float* acc_mean; const float* acc_std;   // [18] each, or all four NULL (no --norm)
  const float* ori_mean; const float* ori_std;   // [54] each
  const float* pos;        // [frames, 15, 3] output of the previous stage, or NULL (f_in = 12)
};
int gc_forward_image(const a3gc_gc_params* p, const float* x, const GcRawInput* raw, uint16_t* img, int64_t batch, int64_t steps,
                     int f_in, int f_out, int act, int split, cudaStream_t stream);
int gc_forward_raw(const a3gc_gc_params* p, const GcRawInput* raw, float* y, int64_t frames, int f_in, int f_out, int act,
                   cudaStream_t stream);
int simt_prepare_input(const float* acc, const float* ori, const float* acc_mean, const float* acc_std,
                       const float* ori_mean, const float* ori_std, float* x, int64_t frames, int ld_x,
                       cudaStream_t stream);
int simt_concat_stage_input(const float* x, const float* pos, float* dst, int64_t frames, cudaStream_t stream);
int simt_reduced_to_full_local(const float* pose, float* out, int64_t frames, int rotsize, cudaStream_t stream);
int train_split_tf32(const float* x, float* hi, float* lo, int64_t n, cudaStream_t stream);
int train_hprev_split(const float* hp, const float* h0, const float* mask, float* hi, float* lo, int64_t batch,
                      int64_t steps, int hidden, int reverse, cudaStream_t stream);
int train_adjacency_grad(const float* dz, const float* u, int64_t records, int hidden, float* partial, int nblocks, float* dP,
                         cudaStream_t stream);
int train_split_mixed(const float* x, int64_t rows, int cols, float* hi, uint16_t* hi16, uint16_t* lo16, int64_t ld, int64_t col0,
                      cudaStream_t stream);
int train_hprev_split_mixed(const float* hp, const float* h0, const float* mask, float* hi, uint16_t* hi16, uint16_t* lo16,
                            int64_t batch, int64_t steps, int hidden, int64_t ld, int64_t col0, int reverse, cudaStream_t stream);


// training path (simt_kernels.cu): forward with a tape, reverse-time backward chain
struct TrainBwdArgs {
  int variant, num_dirs;
  const a3gc_cell_params* cells;
  int reverse[2];
  const float* dy; int64_t dy_stride_b, dy_stride_t, dy_ld;
  const float* c0[2]; const float* dhT[2]; const float* dcT[2];
  float* dh0[2]; float* dc0[2];
  int64_t batch, steps;
  int f_in, hidden, out_act;
  a3gc_tape tape; a3gc_tape_grads grads;
  const float* hmask;
};
size_t simt_train_workspace_bytes(int variant, int f_in, int hidden, int num_dirs);
int simt_train_forward(const LayerArgs& a, const a3gc_tape& tape, const float* hmask, void* ws, size_t ws_bytes, cudaStream_t stream);
int simt_train_backward(const TrainBwdArgs& a, void* ws, size_t ws_bytes, cudaStream_t stream);

// ---- tensor-core (tcgen05) engine: tc_kernels.cu ----------------------------------------
bool tc_layer_supported(int variant, int f_in, int hidden, int precision);
size_t tc_layer_workspace_bytes(int variant, int64_t batch, int64_t steps, int f_in, int hidden, int num_dirs, int precision);
size_t tc_image_bytes(int64_t batch, int64_t steps, int features, int precision);
int tc_layer_forward(const LayerArgs& a, void* ws, size_t ws_bytes, cudaStream_t stream);
size_t tc_packed_weights_bytes(int variant, int f_in, int hidden, int num_dirs, int precision);
int tc_pack_weights(int variant, int num_dirs, const a3gc_cell_params* cells, int f_in, int hidden, int precision, void* packed,
                    cudaStream_t stream);
// tc_gru_kernels.cu: graph-GRU layers on the tensor-core engine
size_t tc_gru_weights_bytes(int f_in, int hidden, int num_dirs, int precision);
int tc_gru_layer_launch(const LayerArgs& a, const uint16_t* x_img, char* weights_ws, cudaStream_t stream);
int tc_gru_pack_weights(int num_dirs, const a3gc_cell_params* cells, int f_in, int hidden, int precision, char* weights_ws, cudaStream_t stream);
int tc_gru_read_trace(unsigned long long* host_out);

}  // namespace a3gc

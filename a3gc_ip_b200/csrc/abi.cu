// extern "C" surface of liba3gc_b200.so (declared in include/a3gc_b200.h): argument validation,
// engine dispatch and the orchestration of one whole net (linear_in -> relu -> rnn1 -> rnn2 -> linear_out).
#include "common.cuh"
#include <cstring>
#include <cstdlib>
#include <vector>
#include <string>

namespace a3gc {

static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int64_t& launch_counter() { return g_launches; }

namespace {

bool variant_ok(int v) { return v >= A3GC_VARIANT_AAGC && v <= A3GC_VARIANT_GGRU; }

int check_cell(int variant, const a3gc_cell_params& c, const char* who) {
  if (variant == A3GC_VARIANT_GGRU) {
    bool ok = c.g_gcn_kernel && c.g_adjacency;
    for (int i = 0; i < 3; ++i) ok = ok && c.dense_in_w[i] && c.dense_in_b[i] && c.dense_hid_w[i];
    if (!ok) { set_error("%s: NULL G-GRU cell parameter", who); return A3GC_ERR_INVALID_ARG; }
    return A3GC_OK;
  }
  bool ok = true;
  for (int g = 0; g < 4; ++g) ok = ok && c.gcn_kernel[g] && c.gcn_bias[g];
  ok = ok && c.adjacency[0];
  if (variant != A3GC_VARIANT_AGC) for (int g = 1; g < 4; ++g) ok = ok && c.adjacency[g];
  if (variant != A3GC_VARIANT_AAGC)
    ok = ok && c.attention_w && c.attention_wq && c.attention_wh && c.attention_u && c.attention_bs && c.attention_bu;
  if (!ok) { set_error("%s: NULL LSTM-family cell parameter", who); return A3GC_ERR_INVALID_ARG; }
  return A3GC_OK;
}

// which engine runs this layer: 1 = SIMT, 2 = TC, <0 = error
int pick_engine(int engine, int variant, int f_in, int hidden, int precision) {
  const bool tc_ok = tc_layer_supported(variant, f_in, hidden, precision);
  if (engine == A3GC_ENGINE_TC) {
    if (!tc_ok) {
      set_error("tensor-core engine does not support variant=%d f_in=%d hidden=%d precision=%d", variant, f_in, hidden, precision);
      return A3GC_ERR_UNSUPPORTED;
    }
    return A3GC_ENGINE_TC;
  }
  if (engine == A3GC_ENGINE_SIMT) {
    if (precision != A3GC_PREC_FP32) { set_error("SIMT engine computes in fp32 only"); return A3GC_ERR_UNSUPPORTED; }
    return A3GC_ENGINE_SIMT;
  }
  if (engine != A3GC_ENGINE_AUTO) { set_error("bad engine %d", engine); return A3GC_ERR_INVALID_ARG; }
  if (tc_ok) return A3GC_ENGINE_TC;
  if (precision != A3GC_PREC_FP32) { set_error("bf16 precision needs the tensor-core engine (hidden multiple of 64)"); return A3GC_ERR_UNSUPPORTED; }
  return A3GC_ENGINE_SIMT;
}

size_t layer_ws(int eng, int variant, int64_t B, int64_t T, int F, int H, int nd, int precision) {
  return eng == A3GC_ENGINE_TC ? tc_layer_workspace_bytes(variant, B, T, F, H, nd, precision)
                               : simt_layer_workspace_bytes(variant, F, H, nd);
}

struct ProfRec { cudaEvent_t e0, e1; double flops; std::string label; };
bool g_prof_on = false;
std::vector<ProfRec> g_prof;

// algorithmic dense FLOPs of one layer launch: gates 2*15*(F+H)*4H per direction-frame,
// attention (minimal form) 34*H^2 + 30*H; G-GRU 2*15*(3*F*H + 4*H*H)
double layer_flops(const LayerArgs& a) {
  const double H = a.hidden, F = a.f_in;
  double per = a.variant == A3GC_VARIANT_GGRU ? 2.0 * 15 * (3 * F * H + 4 * H * H) : 2.0 * 15 * (F + H) * 4 * H;
  if (a.variant == A3GC_VARIANT_A3GC || a.variant == A3GC_VARIANT_AGC) per += 34.0 * H * H + 30.0 * H;
  return per * (double)a.batch * (double)a.steps * a.num_dirs;
}

int run_layer(int eng, const LayerArgs& a, void* ws, size_t ws_bytes, cudaStream_t s) {
  ProfRec r;
  if (g_prof_on) {
    A3GC_CUDA_TRY(cudaEventCreate(&r.e0));
    A3GC_CUDA_TRY(cudaEventCreate(&r.e1));
    A3GC_CUDA_TRY(cudaEventRecord(r.e0, s));
  }
  int rc = eng == A3GC_ENGINE_TC ? tc_layer_forward(a, ws, ws_bytes, s) : simt_layer_forward(a, ws, ws_bytes, s);
  if (g_prof_on) {
    A3GC_CUDA_TRY(cudaEventRecord(r.e1, s));
    char buf[96];
    snprintf(buf, sizeof(buf), "%s:v%d:F%d:H%d", eng == A3GC_ENGINE_TC ? "tc" : "simt", a.variant, a.f_in, a.hidden);
    r.label = buf;
    r.flops = layer_flops(a);
    g_prof.push_back(r);
  }
  return rc;
}

// empty sequence (steps == 0): the reference's time loops do not run and hand the incoming state back
// (net_aagc.py:435-441, :449-456); NULL initial state = zeros
int pass_state_through(const float* const* h0, const float* const* c0, float* const* hT, float* const* cT, int num_dirs,
                       size_t elems, bool gru, cudaStream_t s) {
  for (int d = 0; d < num_dirs; ++d) {
    float* dst[2] = {hT ? hT[d] : nullptr, (cT && !gru) ? cT[d] : nullptr};
    const float* src[2] = {h0 ? h0[d] : nullptr, (c0 && !gru) ? c0[d] : nullptr};
    for (int k = 0; k < 2; ++k) {
      if (!dst[k]) continue;
      if (src[k]) A3GC_CUDA_TRY(cudaMemcpyAsync(dst[k], src[k], elems * sizeof(float), cudaMemcpyDeviceToDevice, s));
      else A3GC_CUDA_TRY(cudaMemsetAsync(dst[k], 0, elems * sizeof(float), s));
    }
  }
  return A3GC_OK;
}

struct NetPlan {
  size_t a0, a1, a2, st, lws, img1, img2, total;   // byte offsets
  size_t lws_bytes;
  int eng1, eng2;
  bool use_img1, use_img2;   // linear_in -> rnn1 and rnn1 -> rnn2 hand-off as operand images (tensor-core engine)
};

int plan_net(int variant, int64_t B, int64_t T, int f0, int H, int precision, int engine, NetPlan* p) {
  p->eng1 = pick_engine(engine, variant, H, H, precision);
  if (p->eng1 < 0) return p->eng1;
  p->eng2 = pick_engine(engine, variant, 2 * H, H, precision);
  if (p->eng2 < 0) return p->eng2;
  const size_t frames = (size_t)B * T;
  const int img_mode = getenv("A3GC_TC_IMG") ? atoi(getenv("A3GC_TC_IMG")) : 3;   // debug: bit0 = fused linear_in image, bit1 = rnn1->rnn2 image
  // the fused linear_in -> operand image kernel covers the small input widths of the nets (f0 <= 32); wider inputs go
  // through the generic graph convolution and tc_pack_x_kernel
  p->use_img1 = p->eng1 == A3GC_ENGINE_TC && f0 <= 32 && H % 16 == 0 && (img_mode & 1);
  p->use_img2 = p->eng1 == A3GC_ENGINE_TC && p->eng2 == A3GC_ENGINE_TC && (img_mode & 2);
  size_t off = 0;
  // fp32 activations are only materialised where a consumer needs them; tensor-core layers exchange operand images
  p->a0 = off; off += p->use_img1 ? 0 : align_up(frames * kNodes * H * sizeof(float), 256);
  p->a1 = off; off += p->use_img2 ? 0 : align_up(frames * kNodes * 2 * H * sizeof(float), 256);
  p->a2 = off; off += align_up(frames * kNodes * 2 * H * sizeof(float), 256);
  p->img1 = off; off += p->use_img1 ? tc_image_bytes(B, T, H, precision) : 0;
  p->img2 = off; off += p->use_img2 ? tc_image_bytes(B, T, 2 * H, precision) : 0;
  p->st = off; off += align_up((size_t)4 * B * kNodes * H * sizeof(float), 256);   // rnn1 final (h,c) x 2 directions
  const size_t l1 = layer_ws(p->eng1, variant, B, T, H, H, 2, precision);
  const size_t l2 = layer_ws(p->eng2, variant, B, T, 2 * H, H, 2, precision);
  p->lws_bytes = l1 > l2 ? l1 : l2;
  p->lws = off; off += align_up(p->lws_bytes, 256);
  p->total = off;
  return A3GC_OK;
}

}  // namespace
}  // namespace a3gc

using namespace a3gc;

extern "C" {

int a3gc_abi_version(void) { return A3GC_ABI_VERSION; }

const char* a3gc_last_error(void) { return g_err; }

int a3gc_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return A3GC_ERR_NO_DEVICE;
  }
  return n;
}

int a3gc_profile_enable(int on) {
  for (auto& r : g_prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  g_prof.clear();
  g_prof_on = on != 0;
  return A3GC_OK;
}
int a3gc_profile_count(void) { return (int)g_prof.size(); }
int a3gc_profile_get(int index, char* label, int label_bytes, float* ms, double* flops) {
  if (index < 0 || index >= (int)g_prof.size()) { set_error("a3gc_profile_get: bad index"); return A3GC_ERR_INVALID_ARG; }
  const ProfRec& r = g_prof[index];
  if (ms) A3GC_CUDA_TRY(cudaEventElapsedTime(ms, r.e0, r.e1));
  if (flops) *flops = r.flops;
  if (label && label_bytes > 0) { strncpy(label, r.label.c_str(), label_bytes - 1); label[label_bytes - 1] = 0; }
  return A3GC_OK;
}

int64_t a3gc_launch_count(void) { return g_launches; }
void a3gc_reset_launch_count(void) { g_launches = 0; }

int a3gc_gc_forward(const a3gc_gc_params* p, const float* x, float* y, int64_t frames, int f_in,
                    int f_out, int act, void* stream) {
  if (!p || !p->gcn_kernel || !p->adj || !p->gcn_bias || (frames > 0 && (!x || !y)) || frames < 0 || f_in <= 0 || f_out <= 0 ||
      act < A3GC_ACT_LINEAR || act > A3GC_ACT_RELU) {
    set_error("a3gc_gc_forward: invalid argument");
    return A3GC_ERR_INVALID_ARG;
  }
  return simt_gc_forward(p, x, y, frames, f_in, f_out, act, static_cast<cudaStream_t>(stream));
}

size_t a3gc_layer_workspace_bytes(int variant, int64_t batch, int64_t steps, int f_in, int hidden,
                                  int num_dirs, int precision, int engine) {
  if (!variant_ok(variant) || num_dirs < 1 || num_dirs > 2 || f_in <= 0 || hidden <= 0) return 0;
  int eng = pick_engine(engine, variant, f_in, hidden, precision);
  if (eng < 0) return 0;
  return layer_ws(eng, variant, batch, steps, f_in, hidden, num_dirs, precision);
}

int a3gc_layer_forward(int variant, int num_dirs, const a3gc_cell_params* cells, const int* reverse,
                       const float* x, int64_t x_stride_b, int64_t x_stride_t,
                       const float* const* h0, const float* const* c0,
                       float* y, int64_t y_stride_b, int64_t y_stride_t, int64_t y_ld,
                       float* const* hT, float* const* cT,
                       int64_t batch, int64_t steps, int f_in, int hidden, int out_act,
                       int precision, int engine, void* workspace, size_t workspace_bytes, void* stream) {
  if (!variant_ok(variant) || num_dirs < 1 || num_dirs > 2 || !cells || !reverse || batch < 0 || steps < 0 ||
      f_in <= 0 || hidden <= 0 || out_act < A3GC_ACT_LINEAR || out_act > A3GC_ACT_TANH) {
    set_error("a3gc_layer_forward: invalid argument");
    return A3GC_ERR_INVALID_ARG;
  }
  if (batch > 0 && steps > 0 && (!x || !y)) { set_error("a3gc_layer_forward: NULL x / y"); return A3GC_ERR_INVALID_ARG; }
  for (int d = 0; d < num_dirs; ++d) {
    int rc = check_cell(variant, cells[d], "a3gc_layer_forward");
    if (rc) return rc;
  }
  if (batch == 0) return A3GC_OK;
  if (steps == 0)
    return pass_state_through(h0, c0, hT, cT, num_dirs, (size_t)batch * kNodes * hidden, variant == A3GC_VARIANT_GGRU, static_cast<cudaStream_t>(stream));
  int eng = pick_engine(engine, variant, f_in, hidden, precision);
  if (eng < 0) return eng;
  LayerArgs a;
  memset(&a, 0, sizeof(a));
  a.variant = variant; a.num_dirs = num_dirs; a.cells = cells;
  for (int d = 0; d < num_dirs; ++d) {
    a.reverse[d] = reverse[d];
    a.h0[d] = h0 ? h0[d] : nullptr;
    a.c0[d] = c0 ? c0[d] : nullptr;
    a.hT[d] = hT ? hT[d] : nullptr;
    a.cT[d] = cT ? cT[d] : nullptr;
  }
  a.x = x; a.x_stride_b = x_stride_b; a.x_stride_t = x_stride_t;
  a.y = y; a.y_stride_b = y_stride_b; a.y_stride_t = y_stride_t; a.y_ld = y_ld;
  a.batch = batch; a.steps = steps; a.f_in = f_in; a.hidden = hidden; a.out_act = out_act; a.precision = precision;
  return run_layer(eng, a, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

size_t a3gc_packed_weights_bytes(int variant, int f_in, int hidden, int num_dirs, int precision, int engine) {
  if (!variant_ok(variant) || num_dirs < 1 || num_dirs > 2 || f_in <= 0 || hidden <= 0) return 0;
  if (pick_engine(engine, variant, f_in, hidden, precision) != A3GC_ENGINE_TC) return 0;
  return tc_packed_weights_bytes(variant, f_in, hidden, num_dirs, precision);
}

int a3gc_pack_weights(int variant, int num_dirs, const a3gc_cell_params* cells, int f_in, int hidden, int precision,
                      int engine, void* packed, size_t packed_bytes, void* stream) {
  if (!variant_ok(variant) || num_dirs < 1 || num_dirs > 2 || !cells || f_in <= 0 || hidden <= 0 || !packed) {
    set_error("a3gc_pack_weights: invalid argument");
    return A3GC_ERR_INVALID_ARG;
  }
  for (int d = 0; d < num_dirs; ++d) {
    int rc = check_cell(variant, cells[d], "a3gc_pack_weights");
    if (rc) return rc;
  }
  const size_t need = a3gc_packed_weights_bytes(variant, f_in, hidden, num_dirs, precision, engine);
  if (need == 0) { set_error("a3gc_pack_weights: this layer does not run on the tensor-core engine (nothing to pack)"); return A3GC_ERR_UNSUPPORTED; }
  if (packed_bytes < need) { set_error("a3gc_pack_weights: buffer too small (%zu < %zu bytes)", packed_bytes, need); return A3GC_ERR_WORKSPACE; }
  return tc_pack_weights(variant, num_dirs, cells, f_in, hidden, precision, packed, static_cast<cudaStream_t>(stream));
}

size_t a3gc_net_workspace_bytes(int variant, int64_t batch, int64_t steps, int f0, int hidden,
                                int f_out, int precision, int engine) {
  (void)f_out;
  if (!variant_ok(variant) || batch < 0 || steps < 0 || f0 <= 0 || hidden <= 0) return 0;
  NetPlan p;
  if (plan_net(variant, batch, steps, f0, hidden, precision, engine, &p) != A3GC_OK) return 0;
  return p.total;
}

static int net_forward_impl(int variant, const a3gc_net_params* net, const float* x, const GcRawInput* raw,
                     const float* const* h0, const float* const* c0, float* y,
                     float* const* hT, float* const* cT,
                     int64_t batch, int64_t steps, int f0, int hidden, int f_out,
                     int precision, int engine, void* workspace, size_t workspace_bytes, void* stream) {
  if (!variant_ok(variant) || !net || batch < 0 || steps < 0 || f0 <= 0 || hidden <= 0 || f_out <= 0) {
    set_error("a3gc_net_forward: invalid argument");
    return A3GC_ERR_INVALID_ARG;
  }
  if (batch * steps > 0 && ((!x && !raw) || !y)) { set_error("a3gc_net_forward: NULL x / y"); return A3GC_ERR_INVALID_ARG; }
  for (int l = 0; l < 2; ++l)
    for (int d = 0; d < 2; ++d) {
      int rc = check_cell(variant, net->rnn[l][d], "a3gc_net_forward");
      if (rc) return rc;
    }
  if (batch == 0) return A3GC_OK;
  if (steps == 0)   // rnn2 is seeded with rnn1's final state = the incoming state
    return pass_state_through(h0, c0, hT, cT, 2, (size_t)batch * kNodes * hidden, variant == A3GC_VARIANT_GGRU, static_cast<cudaStream_t>(stream));
  NetPlan p;
  int rc = plan_net(variant, batch, steps, f0, hidden, precision, engine, &p);
  if (rc) return rc;
  if (workspace_bytes < p.total || !workspace) {
    set_error("a3gc_net_forward: workspace too small (%zu < %zu bytes)", workspace_bytes, p.total);
    return A3GC_ERR_WORKSPACE;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  char* ws = static_cast<char*>(workspace);
  float* a0 = reinterpret_cast<float*>(ws + p.a0);
  float* a1 = reinterpret_cast<float*>(ws + p.a1);
  float* a2 = reinterpret_cast<float*>(ws + p.a2);
  float* st = reinterpret_cast<float*>(ws + p.st);
  const int H = hidden;
  const int64_t frames = batch * steps;
  const size_t state_elems = (size_t)batch * kNodes * H;

  const bool tc1 = p.use_img1;
  uint16_t* img1 = p.use_img1 ? reinterpret_cast<uint16_t*>(ws + p.img1) : nullptr;
  uint16_t* img2 = p.use_img2 ? reinterpret_cast<uint16_t*>(ws + p.img2) : nullptr;
  // linear_in + relu  (net_aagc.py:640-641); on the tensor-core path written straight into rnn1's operand image
  if (tc1) rc = gc_forward_image(&net->linear_in, x, raw, img1, batch, steps, f0, H, A3GC_ACT_RELU, precision == A3GC_PREC_FP32 ? 1 : 0, s);
  else if (raw) rc = gc_forward_raw(&net->linear_in, raw, a0, frames, f0, H, A3GC_ACT_RELU, s);
  else rc = simt_gc_forward(&net->linear_in, x, a0, frames, f0, H, A3GC_ACT_RELU, s);
  if (rc) return rc;

  const int rev[2] = {0, 1};
  float* h1[2] = {st, st + state_elems};
  float* c1[2] = {st + 2 * state_elems, st + 3 * state_elems};
  const bool gru = variant == A3GC_VARIANT_GGRU;

  LayerArgs a;
  memset(&a, 0, sizeof(a));
  a.variant = variant; a.num_dirs = 2; a.cells = net->rnn[0];
  a.reverse[0] = rev[0]; a.reverse[1] = rev[1];
  for (int d = 0; d < 2; ++d) {
    a.h0[d] = h0 ? h0[d] : nullptr;
    a.c0[d] = (c0 && !gru) ? c0[d] : nullptr;
    a.hT[d] = h1[d];
    a.cT[d] = gru ? nullptr : c1[d];
  }
  a.x = tc1 ? nullptr : a0; a.x_stride_b = (int64_t)steps * kNodes * H; a.x_stride_t = (int64_t)kNodes * H;
  a.x_img = img1;
  a.y = img2 ? nullptr : a1; a.y_stride_b = (int64_t)steps * kNodes * 2 * H; a.y_stride_t = (int64_t)kNodes * 2 * H; a.y_ld = 2 * H;
  a.y_img = img2; a.y_img_f = 2 * H;
  a.batch = batch; a.steps = steps; a.f_in = H; a.hidden = H;
  a.out_act = A3GC_ACT_TANH;   // activation_fn='tanh' for both recurrent layers (net_aagc.py:629-630)
  a.precision = precision;
  a.packed = p.eng1 == A3GC_ENGINE_TC ? net->packed_rnn[0] : nullptr;
  rc = run_layer(p.eng1, a, ws + p.lws, p.lws_bytes, s);
  if (rc) return rc;

  // rnn2, seeded with rnn1's final state (net_aagc.py:642-643)
  a.cells = net->rnn[1];
  a.packed = p.eng2 == A3GC_ENGINE_TC ? net->packed_rnn[1] : nullptr;
  for (int d = 0; d < 2; ++d) {
    a.h0[d] = h1[d];
    a.c0[d] = gru ? nullptr : c1[d];
    a.hT[d] = hT ? hT[d] : nullptr;
    a.cT[d] = (cT && !gru) ? cT[d] : nullptr;
  }
  a.x = img2 ? nullptr : a1; a.x_stride_b = (int64_t)steps * kNodes * 2 * H; a.x_stride_t = (int64_t)kNodes * 2 * H;
  a.x_img = img2;
  a.y = a2; a.y_img = nullptr; a.y_img_f = 0;
  a.f_in = 2 * H;
  rc = run_layer(p.eng2, a, ws + p.lws, p.lws_bytes, s);
  if (rc) return rc;

  // linear_out (net_aagc.py:644)
  return simt_gc_forward(&net->linear_out, a2, y, frames, 2 * H, f_out, A3GC_ACT_LINEAR, s);
}

int a3gc_net_forward(int variant, const a3gc_net_params* net, const float* x,
                     const float* const* h0, const float* const* c0, float* y,
                     float* const* hT, float* const* cT,
                     int64_t batch, int64_t steps, int f0, int hidden, int f_out,
                     int precision, int engine, void* workspace, size_t workspace_bytes, void* stream) {
  return net_forward_impl(variant, net, x, nullptr, h0, c0, y, hT, cT, batch, steps, f0, hidden, f_out, precision, engine,
                          workspace, workspace_bytes, stream);
}

int a3gc_net_forward_raw(int variant, const a3gc_net_params* net, const float* acc, const float* ori,
                         const float* acc_mean, const float* acc_std, const float* ori_mean, const float* ori_std,
                         const float* pos,
                         const float* const* h0, const float* const* c0, float* y,
                         float* const* hT, float* const* cT,
                         int64_t batch, int64_t steps, int hidden, int f_out,
                         int precision, int engine, void* workspace, size_t workspace_bytes, void* stream) {
  if (batch * steps > 0 && (!acc || !ori)) { set_error("a3gc_net_forward_raw: NULL acc / ori"); return A3GC_ERR_INVALID_ARG; }
  const bool norm = acc_mean || acc_std || ori_mean || ori_std;
  if (norm && !(acc_mean && acc_std && ori_mean && ori_std)) {
    set_error("a3gc_net_forward_raw: the four normalisation vectors must be given together");
    return A3GC_ERR_INVALID_ARG;
  }
  GcRawInput raw{acc, ori, acc_mean, acc_std, ori_mean, ori_std, pos};
  return net_forward_impl(variant, net, nullptr, &raw, h0, c0, y, hT, cT, batch, steps, pos ? 15 : 12, hidden, f_out, precision,
                          engine, workspace, workspace_bytes, stream);
}

size_t a3gc_layer_train_workspace_bytes(int variant, int64_t batch, int64_t steps, int f_in, int hidden, int num_dirs, int engine) {
  if (!variant_ok(variant) || num_dirs < 1 || num_dirs > 2 || f_in <= 0 || hidden <= 0 || batch < 0 || steps < 0) return 0;
  size_t b = simt_train_workspace_bytes(variant, f_in, hidden, num_dirs);        // backward (and the CUDA-core forward)
  if (variant != A3GC_VARIANT_GGRU && engine != A3GC_ENGINE_SIMT && tc_layer_supported(variant, f_in, hidden, A3GC_PREC_FP32)) {
    const size_t t = tc_layer_workspace_bytes(variant, batch, steps, f_in, hidden, num_dirs, A3GC_PREC_FP32);
    if (t > b) b = t;
  }
  return b;
}

static int check_tape(int variant, const a3gc_tape* t, const char* who) {
  const bool att = variant == A3GC_VARIANT_A3GC || variant == A3GC_VARIANT_AGC;
  if (!t || !t->gates || !t->c || !t->hh || !t->hp || (att && (!t->e || !t->a || !t->q || !t->s))) {
    set_error("%s: incomplete tape", who);
    return A3GC_ERR_INVALID_ARG;
  }
  return A3GC_OK;
}

int a3gc_layer_train_forward(int variant, int num_dirs, const a3gc_cell_params* cells, const int* reverse,
                             const float* x, int64_t x_stride_b, int64_t x_stride_t,
                             const float* const* h0, const float* const* c0,
                             float* y, int64_t y_stride_b, int64_t y_stride_t, int64_t y_ld,
                             float* const* hT, float* const* cT,
                             int64_t batch, int64_t steps, int f_in, int hidden, int out_act,
                             const a3gc_tape* tape, const float* hmask, int engine,
                             void* workspace, size_t workspace_bytes, void* stream) {
  if (!variant_ok(variant)) { set_error("a3gc_layer_train_forward: bad variant %d", variant); return A3GC_ERR_INVALID_ARG; }
  if (num_dirs < 1 || num_dirs > 2 || !cells || !reverse || batch < 0 || steps < 0 || f_in <= 0 || hidden <= 0 ||
      out_act < A3GC_ACT_LINEAR || out_act > A3GC_ACT_TANH || (batch * steps > 0 && (!x || !y))) {
    set_error("a3gc_layer_train_forward: invalid argument");
    return A3GC_ERR_INVALID_ARG;
  }
  for (int d = 0; d < num_dirs; ++d) { int rc = check_cell(variant, cells[d], "a3gc_layer_train_forward"); if (rc) return rc; }
  if (batch == 0 || steps == 0) return A3GC_OK;
  int rc = check_tape(variant, tape, "a3gc_layer_train_forward");
  if (rc) return rc;
  LayerArgs a;
  memset(&a, 0, sizeof(a));
  a.variant = variant; a.num_dirs = num_dirs; a.cells = cells;
  for (int d = 0; d < num_dirs; ++d) {
    a.reverse[d] = reverse[d];
    a.h0[d] = h0 ? h0[d] : nullptr; a.c0[d] = c0 ? c0[d] : nullptr;
    a.hT[d] = hT ? hT[d] : nullptr; a.cT[d] = cT ? cT[d] : nullptr;
  }
  a.x = x; a.x_stride_b = x_stride_b; a.x_stride_t = x_stride_t;
  a.y = y; a.y_stride_b = y_stride_b; a.y_stride_t = y_stride_t; a.y_ld = y_ld;
  a.batch = batch; a.steps = steps; a.f_in = f_in; a.hidden = hidden; a.out_act = out_act; a.precision = A3GC_PREC_FP32;
  if (engine < A3GC_ENGINE_AUTO || engine > A3GC_ENGINE_TC) { set_error("a3gc_layer_train_forward: bad engine %d", engine); return A3GC_ERR_INVALID_ARG; }
  const bool tc_ok = variant != A3GC_VARIANT_GGRU && tc_layer_supported(variant, f_in, hidden, A3GC_PREC_FP32) && x_stride_t == (int64_t)kNodes * f_in;
  if (engine == A3GC_ENGINE_TC && !tc_ok) {
    set_error("a3gc_layer_train_forward: tensor-core engine does not support variant=%d f_in=%d hidden=%d", variant, f_in, hidden);
    return A3GC_ERR_UNSUPPORTED;
  }
  if (engine != A3GC_ENGINE_SIMT && tc_ok) {
    a.tape = tape; a.hmask = hmask;                  // tcgen05 engine in training mode (fp32-parity split, tape, masked state image)
    return run_layer(A3GC_ENGINE_TC, a, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
  }
  return simt_train_forward(a, *tape, hmask, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int a3gc_layer_backward(int variant, int num_dirs, const a3gc_cell_params* cells, const int* reverse,
                        const float* dy, int64_t dy_stride_b, int64_t dy_stride_t, int64_t dy_ld,
                        const float* const* c0, const float* const* dhT, const float* const* dcT,
                        float* const* dh0, float* const* dc0,
                        int64_t batch, int64_t steps, int f_in, int hidden, int out_act,
                        const a3gc_tape* tape, const a3gc_tape_grads* grads, const float* hmask,
                        void* workspace, size_t workspace_bytes, void* stream) {
  if (!variant_ok(variant)) { set_error("a3gc_layer_backward: bad variant %d", variant); return A3GC_ERR_INVALID_ARG; }
  if (num_dirs < 1 || num_dirs > 2 || !cells || !reverse || batch < 0 || steps < 0 || f_in <= 0 || hidden <= 0 ||
      out_act < A3GC_ACT_LINEAR || out_act > A3GC_ACT_TANH || (batch * steps > 0 && !dy)) {
    set_error("a3gc_layer_backward: invalid argument");
    return A3GC_ERR_INVALID_ARG;
  }
  for (int d = 0; d < num_dirs; ++d) { int rc = check_cell(variant, cells[d], "a3gc_layer_backward"); if (rc) return rc; }
  if (batch == 0 || steps == 0) return A3GC_OK;
  int rc = check_tape(variant, tape, "a3gc_layer_backward");
  if (rc) return rc;
  const bool att = variant == A3GC_VARIANT_A3GC || variant == A3GC_VARIANT_AGC;
  if (!grads || !grads->dzm || (att && (!grads->dep || !grads->dqs || !grads->dqp || !grads->dap)) ||
      (variant == A3GC_VARIANT_GGRU && (!grads->dep || !grads->dqs))) {
    set_error("a3gc_layer_backward: incomplete gradient tape");
    return A3GC_ERR_INVALID_ARG;
  }
  TrainBwdArgs a;
  memset(&a, 0, sizeof(a));
  a.variant = variant; a.num_dirs = num_dirs; a.cells = cells;
  for (int d = 0; d < num_dirs; ++d) {
    a.reverse[d] = reverse[d];
    a.c0[d] = c0 ? c0[d] : nullptr; a.dhT[d] = dhT ? dhT[d] : nullptr; a.dcT[d] = dcT ? dcT[d] : nullptr;
    a.dh0[d] = dh0 ? dh0[d] : nullptr; a.dc0[d] = dc0 ? dc0[d] : nullptr;
  }
  a.dy = dy; a.dy_stride_b = dy_stride_b; a.dy_stride_t = dy_stride_t; a.dy_ld = dy_ld;
  a.batch = batch; a.steps = steps; a.f_in = f_in; a.hidden = hidden; a.out_act = out_act;
  a.tape = *tape; a.grads = *grads; a.hmask = hmask;
  return simt_train_backward(a, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int a3gc_prepare_input(const float* acc, const float* ori, const float* acc_mean, const float* acc_std,
                       const float* ori_mean, const float* ori_std, float* x, int64_t frames, int ld_x,
                       void* stream) {
  if (frames < 0 || ld_x < 12 || (frames > 0 && (!acc || !ori || !x)) || ((acc_mean == nullptr) != (acc_std == nullptr)) ||
      ((ori_mean == nullptr) != (ori_std == nullptr)) || ((acc_mean == nullptr) != (ori_mean == nullptr))) {
    set_error("a3gc_prepare_input: invalid argument");
    return A3GC_ERR_INVALID_ARG;
  }
  return simt_prepare_input(acc, ori, acc_mean, acc_std, ori_mean, ori_std, x, frames, ld_x, static_cast<cudaStream_t>(stream));
}

int a3gc_reduced_to_full_local(const float* pose, float* out, int64_t frames, int rotsize, void* stream) {
  if (frames < 0 || (rotsize != 9 && rotsize != 6) || (frames > 0 && (!pose || !out))) {
    set_error("a3gc_reduced_to_full_local: invalid argument (rotsize must be 9 or 6)");
    return A3GC_ERR_INVALID_ARG;
  }
  return simt_reduced_to_full_local(pose, out, frames, rotsize, static_cast<cudaStream_t>(stream));
}

int a3gc_train_split_tf32(const float* x, float* hi, float* lo, int64_t n, void* stream) {
  if (n < 0 || (n > 0 && (!x || !hi || !lo))) { set_error("a3gc_train_split_tf32: invalid argument"); return A3GC_ERR_INVALID_ARG; }
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(lo)) & 15) {
    set_error("a3gc_train_split_tf32: buffers must be 16-byte aligned");
    return A3GC_ERR_INVALID_ARG;
  }
  return train_split_tf32(x, hi, lo, n, static_cast<cudaStream_t>(stream));
}

int a3gc_train_hprev_split(const float* hp, const float* h0, const float* mask, float* hi, float* lo, int64_t batch,
                           int64_t steps, int hidden, int reverse, void* stream) {
  if (batch < 0 || steps < 0 || hidden <= 0 || hidden % 4 != 0 || (batch * steps > 0 && (!hp || !hi || !lo))) {
    set_error("a3gc_train_hprev_split: invalid argument (hidden must be a positive multiple of 4)");
    return A3GC_ERR_INVALID_ARG;
  }
  return train_hprev_split(hp, h0, mask, hi, lo, batch, steps, hidden, reverse, static_cast<cudaStream_t>(stream));
}

int a3gc_train_split_mixed(const float* x, int64_t rows, int cols, float* hi, uint16_t* hi16, uint16_t* lo16, int64_t ld, int64_t col0,
                           void* stream) {
  if (rows < 0 || cols < 0 || cols % 4 != 0 || ld % 4 != 0 || col0 % 4 != 0 || col0 < 0 || col0 + cols > ld ||
      (rows * cols > 0 && (!x || !hi || !hi16 || !lo16))) {
    set_error("a3gc_train_split_mixed: invalid argument (cols, ld and col0 must be multiples of 4, col0 + cols <= ld)");
    return A3GC_ERR_INVALID_ARG;
  }
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(hi)) & 15 || (reinterpret_cast<uintptr_t>(hi16) | reinterpret_cast<uintptr_t>(lo16)) & 7) {
    set_error("a3gc_train_split_mixed: x / hi must be 16-byte aligned, hi16 / lo16 8-byte aligned");
    return A3GC_ERR_INVALID_ARG;
  }
  return train_split_mixed(x, rows, cols, hi, hi16, lo16, ld, col0, static_cast<cudaStream_t>(stream));
}

int a3gc_train_hprev_split_mixed(const float* hp, const float* h0, const float* mask, float* hi, uint16_t* hi16, uint16_t* lo16,
                                 int64_t batch, int64_t steps, int hidden, int64_t ld, int64_t col0, int reverse, void* stream) {
  if (batch < 0 || steps < 0 || hidden <= 0 || hidden % 4 != 0 || ld % 4 != 0 || col0 % 4 != 0 || col0 < 0 || col0 + hidden > ld ||
      (batch * steps > 0 && (!hp || !hi || !hi16 || !lo16))) {
    set_error("a3gc_train_hprev_split_mixed: invalid argument (hidden, ld and col0 must be multiples of 4, col0 + hidden <= ld)");
    return A3GC_ERR_INVALID_ARG;
  }
  return train_hprev_split_mixed(hp, h0, mask, hi, hi16, lo16, batch, steps, hidden, ld, col0, reverse, static_cast<cudaStream_t>(stream));
}

int a3gc_train_adjacency_grad(const float* dz, const float* u, int64_t records, int hidden, float* partial, int nblocks, float* dP,
                              void* stream) {
  if (records < 0 || hidden <= 0 || hidden % 16 != 0 || nblocks <= 0 || !dz || !u || !partial || !dP) {
    set_error("a3gc_train_adjacency_grad: invalid argument (hidden must be a positive multiple of 16, nblocks > 0)");
    return A3GC_ERR_INVALID_ARG;
  }
  if ((reinterpret_cast<uintptr_t>(dz) | reinterpret_cast<uintptr_t>(u)) & 15) {
    set_error("a3gc_train_adjacency_grad: dz / u must be 16-byte aligned");
    return A3GC_ERR_INVALID_ARG;
  }
  return train_adjacency_grad(dz, u, records, hidden, partial, nblocks, dP, static_cast<cudaStream_t>(stream));
}

int a3gc_concat_stage_input(const float* x, const float* pos, float* dst, int64_t frames, void* stream) {
  if (frames < 0 || (frames > 0 && (!x || !pos || !dst))) { set_error("a3gc_concat_stage_input: invalid argument"); return A3GC_ERR_INVALID_ARG; }
  return simt_concat_stage_input(x, pos, dst, frames, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

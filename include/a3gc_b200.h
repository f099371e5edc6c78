/*
 * a3gc_b200.h -- C ABI of the B200-native A3GC-IP hot path.
 *
 * Drop-in boundary for the recurrent adaptive-graph-convolution path of trikpachu/A3GC-IP
 * (reference file net_aagc.py:40-695, chained as in evaluate_a3gc_tp.py:164-172).  The reference
 * is pure PyTorch, so the "FFI" a maintainer binds is a ctypes stub: the
 * drop-in torch.nn.Modules in a3gc_ip_b200/net_aagc.py hand raw device pointers of their
 * parameters and activations to these entry points.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a CUDA device pointer to fp32 data unless
 *     stated otherwise; the caller (torch) owns all memory, including the workspace;
 *   - the library never allocates persistent device memory and never frees caller memory;
 *   - every entry point returns 0 on success, a negative a3gc_status on error, and never throws;
 *     a3gc_last_error() returns a thread-local description of the last failure;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises;
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 *   - the graph always has A3GC_NODES = 15 nodes (net_aagc.py:142,233,319 assert it).
 */
#ifndef A3GC_B200_H_
#define A3GC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define A3GC_ABI_VERSION 3
#define A3GC_NODES 15

/* Cell family.  Values are part of the ABI. */
typedef enum a3gc_variant {
  A3GC_VARIANT_AAGC = 0, /* AAGC_LSTM_cell  net_aagc.py:68-126  (4 learnable adjacencies, no attention) */
  A3GC_VARIANT_A3GC = 1, /* A3GC_LSTM_cell  net_aagc.py:128-217 (4 learnable adjacencies + joint attention) */
  A3GC_VARIANT_AGC = 2,  /* AGC_LSTM_cell   net_aagc.py:219-303 (1 frozen adjacency, applied transposed, + attention) */
  A3GC_VARIANT_GGRU = 3  /* G_GRU_cell      net_aagc.py:305-368 */
} a3gc_variant;

typedef enum a3gc_activation {
  A3GC_ACT_LINEAR = 0, /* 'linear' (net_aagc.py:46) */
  A3GC_ACT_TANH = 1,   /* 'tanh'   (net_aagc.py:48) */
  A3GC_ACT_RELU = 2    /* the torch.relu the nets apply after linear_in (net_aagc.py:641) */
} a3gc_activation;

/* Arithmetic of the gate / attention contractions. */
typedef enum a3gc_precision {
  A3GC_PREC_FP32 = 0, /* fp32-accurate (parity bar 1e-4 rel vs the reference's CPU fp32 forward) */
  A3GC_PREC_BF16 = 1  /* bf16 operands, fp32 accumulate and state (stated looser bound) */
} a3gc_precision;

/* Which kernels run the recurrent layers. */
typedef enum a3gc_engine {
  A3GC_ENGINE_AUTO = 0, /* tensor-core engine when the shape allows it, else SIMT */
  A3GC_ENGINE_SIMT = 1, /* fp32 CUDA-core kernels, any shape */
  A3GC_ENGINE_TC = 2    /* tcgen05 / TMEM kernels (hidden size multiple of 64); error if unsupported */
} a3gc_engine;

typedef enum a3gc_status {
  A3GC_OK = 0,
  A3GC_ERR_INVALID_ARG = -1,
  A3GC_ERR_UNSUPPORTED = -2,
  A3GC_ERR_WORKSPACE = -3,
  A3GC_ERR_NO_DEVICE = -4,
  A3GC_ERR_CUDA = -5
} a3gc_status;

/* Parameters of one AAGC graph convolution (class AAGC, net_aagc.py:55-57). */
typedef struct a3gc_gc_params {
  const float* gcn_kernel; /* [f_out, f_in] */
  const float* adj;        /* [15, 15], used as adj @ x (net_aagc.py:63) */
  const float* gcn_bias;   /* [f_out] */
} a3gc_gc_params;

/*
 * Parameters of one recurrent cell, exactly the tensors of the reference's state_dict, in the
 * reference's storage layout (no repacking by the caller).  Gate order is i, f, c, o.
 *   LSTM family (AAGC / A3GC / AGC): gcn_kernel[g] [H, F+H] (columns: x features first, then h --
 *   the cat order of net_aagc.py:182), gcn_bias[g] [H], adjacency[g] [15,15].
 *   AGC stores ONE frozen `adjacency` (net_aagc.py:238): pass it in adjacency[0]; [1..3] are ignored.
 *   attention_* (A3GC / AGC only, NULL for AAGC): w, wq, wh [H,H]; u [1,H]; bs [H]; bu [15].
 *   G-GRU: g_gcn_kernel [H,H], g_adjacency [15,15] (used transposed, net_aagc.py:348),
 *   dense_in_w[r,u,c] [H,F], dense_in_b[r,u,c] [H], dense_hid_w[r,u,c] [H,H].  The frozen,
 *   unused `a` buffer of G_GRU_cell (net_aagc.py:324) never crosses the ABI.
 */
typedef struct a3gc_cell_params {
  const float* gcn_kernel[4];
  const float* adjacency[4];
  const float* gcn_bias[4];
  const float* attention_w;
  const float* attention_wq;
  const float* attention_wh;
  const float* attention_u;
  const float* attention_bs;
  const float* attention_bu;
  const float* g_gcn_kernel;
  const float* g_adjacency;
  const float* dense_in_w[3];
  const float* dense_in_b[3];
  const float* dense_hid_w[3];
} a3gc_cell_params;

/* One whole net: linear_in -> relu -> rnn1 (2 directions) -> rnn2 (2 directions) -> linear_out
 * (A3GC_net & co., net_aagc.py:595-695).  rnn[layer][direction]; direction 1 is the reverse layer. */
typedef struct a3gc_net_params {
  a3gc_gc_params linear_in;
  a3gc_cell_params rnn[2][2];
  a3gc_gc_params linear_out;
  /* Optional (NULL = repack on every call, the stateless default): the output of a3gc_pack_weights for rnn1 / rnn2
   * (both directions), packed from the CURRENT values of rnn[l][*] at the same precision.  The library only reads it;
   * the caller owns the buffer and re-packs after any parameter update (cached packed weights). */
  const void* packed_rnn[2];
} a3gc_net_params;

/* Version of this ABI (A3GC_ABI_VERSION of the library that was built). */
int a3gc_abi_version(void);

/* Thread-local text of the last error returned on this thread ("" if none). */
const char* a3gc_last_error(void);

/* Number of CUDA devices visible to the library; negative a3gc_status if the runtime fails. */
int a3gc_device_count(void);

/*
 * AAGC.forward (net_aagc.py:61-66):  y = act((adj @ x) @ W^T + b), eval mode.
 *   x [frames, 15, f_in] contiguous, y [frames, 15, f_out] contiguous; frames = B*T.
 */
int a3gc_gc_forward(const a3gc_gc_params* p, const float* x, float* y, int64_t frames, int f_in,
                    int f_out, int act, void* stream);

/* Bytes of caller-provided workspace a3gc_layer_forward needs (0 is possible). */
size_t a3gc_layer_workspace_bytes(int variant, int64_t batch, int64_t steps, int f_in, int hidden,
                                  int num_dirs, int precision, int engine);

/*
 * Time loop of num_dirs (1 or 2) recurrent layer directions over the same input, run concurrently:
 * A3GC_LSTM / ReverseA3GC_LSTM / BiA3GC_LSTM .forward (net_aagc.py:435-441, :449-456, :469-480) and
 * their AAGC / AGC / G_GRU twins.
 *   cells[d], reverse[d]   parameters and walking order of direction d (reverse: t = T-1 .. 0,
 *                          outputs stored at their own t, final state = state after t = 0);
 *   x      element (b, t, n, k) at x[b*x_stride_b + t*x_stride_t + n*f_in + k]   (any of [B,T,..] / [T,B,..]);
 *   h0[d], c0[d]           initial state [B, 15, H] contiguous, NULL = zeros (c0 ignored for G-GRU);
 *   y      element (b, t, n, j) of direction d at y[b*y_stride_b + t*y_stride_t + n*y_ld + d*H + j];
 *          holds out_act(h') for the LSTM family, h' for G-GRU (its activation_fn is unused, :368);
 *   hT[d], cT[d]           final state [B, 15, H] contiguous (may be NULL to skip; cT ignored for G-GRU).
 */
int a3gc_layer_forward(int variant, int num_dirs, const a3gc_cell_params* cells, const int* reverse,
                       const float* x, int64_t x_stride_b, int64_t x_stride_t,
                       const float* const* h0, const float* const* c0,
                       float* y, int64_t y_stride_b, int64_t y_stride_t, int64_t y_ld,
                       float* const* hT, float* const* cT,
                       int64_t batch, int64_t steps, int f_in, int hidden, int out_act,
                       int precision, int engine, void* workspace, size_t workspace_bytes, void* stream);

/*
 * Packed weights of one (bi)layer for the tensor-core engine: operand images of the gate / attention (G-GRU: input and
 * fused message) weights, mix matrices and biases of `num_dirs` directions, in the form the layer kernel streams.
 * a3gc_packed_weights_bytes returns 0 when the layer would not run on the tensor-core engine (shape / precision / engine),
 * i.e. there is nothing to cache.  a3gc_pack_weights enqueues the packing kernels on `stream`.
 */
size_t a3gc_packed_weights_bytes(int variant, int f_in, int hidden, int num_dirs, int precision, int engine);
int a3gc_pack_weights(int variant, int num_dirs, const a3gc_cell_params* cells, int f_in, int hidden, int precision,
                      int engine, void* packed, size_t packed_bytes, void* stream);

/* Bytes of workspace a3gc_net_forward needs for this shape. */
size_t a3gc_net_workspace_bytes(int variant, int64_t batch, int64_t steps, int f0, int hidden,
                                int f_out, int precision, int engine);

/*
 * {AAGC,A3GC,AGC,G_GRU}_net.forward (net_aagc.py:607-619, :633-645, :659-671, :685-695), eval mode:
 * linear_in -> relu -> rnn1 -> rnn2 (seeded with rnn1's final state) -> linear_out.
 *   x [B, T, 15, f0] contiguous;  y [B, T, 15, f_out] contiguous;
 *   h0[d], c0[d]  initial state of rnn1 direction d ([B,15,H]; NULL = zeros; c0 unused for G-GRU);
 *   hT[d], cT[d]  final state of rnn2 direction d (may be NULL).
 */
int a3gc_net_forward(int variant, const a3gc_net_params* net, const float* x,
                     const float* const* h0, const float* const* c0, float* y,
                     float* const* hT, float* const* cT,
                     int64_t batch, int64_t steps, int f0, int hidden, int f_out,
                     int precision, int engine, void* workspace, size_t workspace_bytes, void* stream);

/*
 * The same net fed with the RAW IMU frame: prepare_input (evaluate_a3gc_tp.py:64-94: (v - mean) / std per channel when the
 * four statistics vectors are given (--norm), the 6th IMU dropped, IMU j = (acc[3j..3j+2], ori[9j..9j+8]) scattered onto
 * node [3,4,13,14,10][j], all other nodes zero) and, when pos != NULL, the stage concatenation
 * torch.cat((x, pos.view(B,T,15,3)), dim=-1) (evaluate_a3gc_tp.py:168, :170) are fused into the load of linear_in, so
 * the [B,T,15,12] / [B,T,15,15] tensors are never materialised (288 B of input per frame instead of 720).
 *   acc [B, T, 18], ori [B, T, 54] contiguous;  acc_mean/acc_std [18], ori_mean/ori_std [54] or all NULL;
 *   pos [B, T, 15, 3] (previous stage's output) or NULL;  units_in of the net is 15 with pos, 12 without.
 * Workspace: a3gc_net_workspace_bytes with f0 = 12 or 15.
 */
int a3gc_net_forward_raw(int variant, const a3gc_net_params* net, const float* acc, const float* ori,
                         const float* acc_mean, const float* acc_std, const float* ori_mean, const float* ori_std,
                         const float* pos,
                         const float* const* h0, const float* const* c0, float* y,
                         float* const* hT, float* const* cT,
                         int64_t batch, int64_t steps, int hidden, int f_out,
                         int precision, int engine, void* workspace, size_t workspace_bytes, void* stream);

/*
 * ---- Training path (BPTT) of the LSTM family: train_*_tp.py:74-84 calls model.forward in train mode, then
 * loss.backward().  The forward keeps a tape of per-step intermediates; a3gc_layer_backward walks the
 * recurrence in reverse and leaves, per step, the operands of the weight / input gradient contractions,
 * which are plain batched GEMMs over all (t, b) the caller forms afterwards.
 *
 * All tape arrays are caller-allocated fp32.  Unit-major arrays hold one record per (direction d, time t,
 * sequence b) at index (d*T + t)*B + b, each unit row = the 15 nodes padded to 16 (slot 15 = 0); node-major
 * arrays are [D][B][T][15][.] like the layer input.
 * G-GRU (net_aagc.py:343-368) reuses the structs with this meaning: tape.gates = (r, u, c, Wch msg), tape.c = h Wg^T before
 * the node mix, tape.hh = msg, tape.hp = h'; grads.dzm = (dzr, dzu, dzc, dzc r) node-major [D][B][T][15][4H], grads.dep = dmsg
 * (unit-major [D][T][B][H][16]), grads.dqs = dM = P^T dmsg node-major [D][B][T][15][H]; the other fields are unused, and
 * a3gc_layer_backward's c0 argument carries the forward's initial h (the G-GRU state is a single tensor).
 */
typedef struct a3gc_tape {
  float* gates; /* [D][T][B][4][H][16] activated i, f, c~, o; OVERWRITTEN by the backward with dz (gate pre-activation grads) */
  float* u;     /* [D][T][B][4][H][16] pre-mix accumulators S W_g^T (operand of the adjacency gradients); may be NULL */
  float* c;     /* [D][T][B][H][16]    c'_t */
  float* hh;    /* [D][T][B][H][16]    hy_t before attention */
  float* e;     /* [D][T][B][H][16]    e_t = tanh(Wh hy + Wq q + bs)        (A3GC / AGC) */
  float* hp;    /* [D][B][T][15][H]    h'_t (the carried state; y_t = act(h'_t)), node-major like x */
  float* a;     /* [D][T][B][16]       sigmoid attention weights           (A3GC / AGC) */
  float* q;     /* [D][T][B][H]        q_t                                  (A3GC / AGC) */
  float* s;     /* [D][T][B][H]        node sums of hy_t                    (A3GC / AGC) */
} a3gc_tape;

typedef struct a3gc_tape_grads {
  float* dzm;   /* [D][B][T][15][4H]   P_g^T dz_g (column g*H + j), rows as in x: dW = dzm^T [x | h_prev], dX = dzm W_x */
  float* dep;   /* [D][T][B][H][16]    grads of the tanh pre-activation of e_t (dWh, dbs)   (A3GC / AGC) */
  float* dqs;   /* [D][T][B][H]        node sums of dep (dWq)                                */
  float* dqp;   /* [D][T][B][H]        grads of the relu pre-activation of q_t (dWa)         */
  float* dap;   /* [D][T][B][16]       grads of the sigmoid pre-activation of a_t (du, dbu)  */
  /* ABI 3, LSTM family, optional (NULL = dzm is plain fp32): dzm in the "mixed" form of the hoisted GEMMs, written by the
     backward itself -- dzm then holds the TF32-exact head, these two bf16(head) and bf16(value - head), same shape.
     The graph-GRU backward ignores them. */
  uint16_t* dzm_hi16;
  uint16_t* dzm_lo16;
} a3gc_tape_grads;

/* Workspace for a3gc_layer_train_forward AND a3gc_layer_backward of this shape (the larger of the two). */
size_t a3gc_layer_train_workspace_bytes(int variant, int64_t batch, int64_t steps, int f_in, int hidden, int num_dirs, int engine);

/* a3gc_layer_forward in training mode: same arguments plus the tape and an optional recurrent-dropout mask hmask
 * [D][B][T][15][H] (0 or 1/(1-p), net_aagc.py:181; NULL = no dropout).  Input dropout (:180) is applied by the caller
 * to x.  engine: A3GC_ENGINE_AUTO / TC run the tcgen05 kernel in training mode (fp32-parity operand split; hidden in
 * {64,128,256}, f_in % 16 == 0, batch-major x), A3GC_ENGINE_SIMT the CUDA-core kernel (any shape). */
int a3gc_layer_train_forward(int variant, int num_dirs, const a3gc_cell_params* cells, const int* reverse,
                             const float* x, int64_t x_stride_b, int64_t x_stride_t,
                             const float* const* h0, const float* const* c0,
                             float* y, int64_t y_stride_b, int64_t y_stride_t, int64_t y_ld,
                             float* const* hT, float* const* cT,
                             int64_t batch, int64_t steps, int f_in, int hidden, int out_act,
                             const a3gc_tape* tape, const float* hmask, int engine,
                             void* workspace, size_t workspace_bytes, void* stream);

/* Reverse-time chain.  dy: gradient of y (same addressing as y); c0[d]: the forward's initial cell state;
 * dhT/dcT[d]: gradients of the final state (NULL = 0); dh0/dc0[d]: receive the gradients of the initial
 * state (NULL = skip). */
int a3gc_layer_backward(int variant, int num_dirs, const a3gc_cell_params* cells, const int* reverse,
                        const float* dy, int64_t dy_stride_b, int64_t dy_stride_t, int64_t dy_ld,
                        const float* const* c0, const float* const* dhT, const float* const* dcT,
                        float* const* dh0, float* const* dc0,
                        int64_t batch, int64_t steps, int f_in, int hidden, int out_act,
                        const a3gc_tape* tape, const a3gc_tape_grads* grads, const float* hmask,
                        void* workspace, size_t workspace_bytes, void* stream);

/*
 * prepare_input (evaluate_a3gc_tp.py:64-94) on the device: per frame normalise the raw 18-d
 * acceleration and 54-d orientation channels ((v - mean) / std; pass NULL mean/std to skip, i.e.
 * --norm off), drop the 6th IMU and scatter (acc_i, ori_i) of IMU i onto node {3,4,13,14,10}[i].
 *   acc [frames, 18], ori [frames, 54] -> x [frames, 15, ld_x] (columns 0..11 written; other nodes' 0..11 zeroed).
 */
int a3gc_prepare_input(const float* acc, const float* ori, const float* acc_mean, const float* acc_std,
                       const float* ori_mean, const float* ori_std, float* x, int64_t frames, int ld_x,
                       void* stream);

/*
 * Stage chaining of evaluate_a3gc_tp.py:168,170:  dst[f, n, 0..12) = x[f, n, 0..12),
 * dst[f, n, 12..15) = pos[f, n, 0..3)   (torch.cat((x, pos), dim=-1)).
 */
int a3gc_concat_stage_input(const float* x, const float* pos, float* dst, int64_t frames, void* stream);

/*
 * Reduced-global -> full-local pose post-step of forward_offline (PoseNet3._reduced_glb_to_full_local_mat /
 * _reduced_glb_6d_to_full_local_mat, net_aagc.py:788-800, 825-829): scatter the 15 predicted global joint rotations
 * into the 24 SMPL joints by joint_set.reduced (config.py:29), identity elsewhere; inverse kinematics along the SMPL
 * tree, R_local[i] = R_global[parent[i]]^T R_global[i] (articulate/math/spatial.py:115-123, 197-221); identity on
 * joint_set.ignored (config.py:30).  The parent table is the published SMPL kinematic tree (the reference reads it
 * from the SMPL model file, articulate/model.py:37, which it does not ship).
 *   pose [frames, 15, rotsize] (rotsize 9: row-major 3x3; 6: 6D rotation, articulate/math/angular.py:167-182)
 *   out  [frames, 24, 3, 3]
 */
int a3gc_reduced_to_full_local(const float* pose, float* out, int64_t frames, int rotsize, void* stream);

/*
 * Operand preparation of the hoisted weight / input gradient GEMMs of the training step (the autograd mm nodes of
 * train_a3gc_tp.py:77, loss.backward()).  They run as three TF32 tensor-core passes hi*hi + lo*hi + hi*lo:
 * a3gc_train_split_tf32 cuts x[n] into hi (exactly representable in TF32, round-to-nearest) and lo = x - hi.
 * a3gc_train_hprev_split builds the h half of S = [x | h_prev] of one direction directly in split form:
 * h_prev(b, t) = hp(b, t -/+ 1) (h0, or zeros if null, at the direction's first step) times mask(b, t) if given;
 * hp, mask, hi, lo are [batch, steps, 15, hidden], h0 is [batch, 15, hidden].  All buffers 16-byte aligned.
 */
int a3gc_train_split_tf32(const float* x, float* hi, float* lo, int64_t n, void* stream);
int a3gc_train_hprev_split(const float* hp, const float* h0, const float* mask, float* hi, float* lo, int64_t batch,
                           int64_t steps, int hidden, int reverse, void* stream);

/*
 * The same builders in "mixed" form (ABI 3): the hi*hi product stays TF32, the two correction products run on bf16 copies
 * of their operands (a correction is 2^-11 of the result, bf16 rounding 2^-9 of that).  One pass writes hi (fp32, exactly
 * representable in TF32), hi16 = bf16(hi) and lo16 = bf16(x - hi) into row-major buffers of row length ld at column col0,
 * so that x (cols = F, col0 = 0) and h_prev (col0 = F) form one S = [x | h_prev] operand of row length F + H:
 * dW = dzm^T S is then one GEMM per precision instead of two.  x is [rows, cols] contiguous; cols, ld, col0 multiples of 4.
 */
int a3gc_train_split_mixed(const float* x, int64_t rows, int cols, float* hi, uint16_t* hi16, uint16_t* lo16, int64_t ld, int64_t col0,
                           void* stream);
int a3gc_train_hprev_split_mixed(const float* hp, const float* h0, const float* mask, float* hi, uint16_t* hi16, uint16_t* lo16,
                                 int64_t batch, int64_t steps, int hidden, int64_t ld, int64_t col0, int reverse, void* stream);

/*
 * Adjacency gradients of one direction of an A3GC / AAGC layer (ABI 3; the bmm + sum autograd nodes of loss.backward(),
 * train_a3gc_tp.py:83): dP[g][m][n] = sum over records r and units j of dz[r][g][j][m] * u[r][g][j][n], with dz the gate
 * pre-activation gradients a3gc_layer_backward left in tape.gates and u = tape.u, both [records][4][hidden][16].
 * One pass over the two arrays; partial is scratch of nblocks * 1024 floats (nblocks: a few per SM), dP is [4][16][16]
 * (rows / columns 15 are padding).  The sum order is fixed: the result is deterministic.
 */
int a3gc_train_adjacency_grad(const float* dz, const float* u, int64_t records, int hidden, float* partial, int nblocks, float* dP,
                              void* stream);

/*
 * Optional per-launch timing of the recurrent-layer kernels (used by bench.py for the roofline):
 * while enabled, every layer launch is bracketed by CUDA events on the launching stream.
 * a3gc_profile_get must be called after the stream has been synchronised; it returns the launch's
 * duration in ms, its algorithmic dense FLOPs (2*M*N*K of the gate + attention contractions, the
 * SURVEY.md 8d convention), and a short label "<engine>:<variant>:F<f_in>:H<hidden>".
 */
int a3gc_profile_enable(int on);            /* on != 0: start a fresh recording; 0: stop */
int a3gc_profile_count(void);
int a3gc_profile_get(int index, char* label, int label_bytes, float* ms, double* flops);

/*
 * Self-test of the tcgen05 building blocks (bulk copy + mbarrier, UMMA descriptors, TMEM read-back):
 * D[128, n] (fp32) = A[128, k] * B[n, k]^T with 16-bit operands given as K-major no-swizzle images
 * [k/8][rows][8].  flags bit 0: operands are bf16 (else fp16); bit 1: swap the LBO / SBO descriptor
 * fields (diagnostic).  One CTA; k %% 16 == 0, 16 <= n <= 256, n %% 16 == 0.
 */
int a3gc_tc_selftest(const void* a_img, const void* b_img, float* d, int k, int n, int flags, void* stream);

/* Tuning aid: cycles per tcgen05.mma (M=128, N=n, K=16, fp16) for a shared-memory operand layout given by the
 * descriptor fields (layout_type 0 = no swizzle, 2 = 128-byte swizzle; LBO / SBO / per-K-step start offsets in
 * bytes); `iters` back-to-back MMAs cycling over nk K-steps, `grid` CTAs; cycles_out: device pointer to one float. */
int a3gc_tc_mma_bench(int n, int layout_type, int a_lbo, int a_sbo, int b_lbo, int b_sbo, int a_kstep, int b_kstep,
                      int nk, int iters, int grid, float* cycles_out, void* stream);

/* Tuning aid: L2 -> shared-memory bulk-copy stream rate (bytes per cycle per CTA) with `grid` CTAs each pulling `iters`
 * chunks of `chunk` bytes from a `span`-byte buffer through a `depth`-slot ring; `nprod` producer threads (one per warp,
 * each owning every nprod-th slot); mcast != 0: 2-CTA clusters, each CTA fetches half of every chunk and multicasts it to
 * both.  out: device pointer to one float. */
int a3gc_tc_stream_bench(const void* src, size_t span, int chunk, int depth, int iters, int grid, int mcast, int nprod, float* out, void* stream);

/* Debug: per-phase clock64 timeline of CTA (0,0) of the last tensor-core layer launch made with the
 * environment variable A3GC_TC_TRACE set; host_out receives [2 roles][16 steps][16 slots] uint64. */
int a3gc_debug_read_tc_trace(unsigned long long* host_out);

/* Tuning aid: clusters of `cluster_size` CTAs of the tensor-core layer kernel (smem_bytes dynamic shared memory per CTA)
 * the device can hold at once (cudaOccupancyMaxActiveClusters); negative a3gc_status on error. */
int a3gc_debug_max_active_clusters(int cluster_size, int smem_bytes);

/* Number of kernels this library has launched on the calling thread since the last reset. */
int64_t a3gc_launch_count(void);
void a3gc_reset_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* A3GC_B200_H_ */

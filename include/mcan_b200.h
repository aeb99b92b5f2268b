/*
 * mcan_b200.h -- C ABI of the B200-native MCAN co-attention hot path.
 *
 * The reference (Originofamonia/mcan-vqa) is pure Python/PyTorch and has no FFI of its
 * own; every arithmetic step of its hot path is an ATen call issued from
 *   core/model/mca.py      (MHAtt 18-78, FFN 85-98, SA 105-127, SGA 134-164, MCA_ED 171-186)
 *   core/model/net_utils.py (FC 11-34, MLP 37-45, LayerNorm 48-60)
 *   core/model/net.py      (AttFlat 20-55).
 * This header is the boundary a maintainer of the reference binds to instead of those
 * ATen calls (ctypes stub: see INTEGRATION.md).  Each entry point names the reference
 * lines it replaces.
 *
 * Conventions
 *   - plain C: raw device pointers, sizes, strides in ELEMENTS, POD arg structs.
 *   - the caller owns every buffer (inputs, outputs, workspaces); nothing is allocated
 *     here except a process-wide cache of TMA descriptors.
 *   - every call only ENQUEUES work on `stream` (a cudaStream_t passed as void*); it never
 *     synchronises.  Safe to capture into a CUDA graph.
 *   - return 0 on success, <0 on error; mcan_last_error() returns a thread-local message.
 *   - sm_100a only.  There is no CPU path: without a B200 the calls fail with an error.
 *   - "bf16" pointers are `void*` to __nv_bfloat16 data; activations are row-major
 *     [rows, features] with rows = batch * sequence.
 *   - dropout masks are never stored: forward and backward regenerate them from
 *     (seed, element index) with the hash in csrc/common.cuh.  The effective seed is
 *     dropout_seed ^ *dropout_seed_dev when the device pointer is given, so a captured CUDA
 *     graph draws a fresh mask on every replay by updating one device word.
 */
#ifndef MCAN_B200_H_
#define MCAN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCAN_B200_ABI_VERSION 1
#define MCAN_MAX_GEMM_SEGMENTS 3

/* -- library ----------------------------------------------------------------------- */
int mcan_version(void);
const char* mcan_last_error(void);
/* number of SMs of the current device (persistent kernels size their grids with it) */
int mcan_num_sms(void);
/* Persistent kernels use at most `sms` SMs (0 = all).  Data-parallel training leaves a few SMs to
 * the NCCL all-reduce that overlaps the backward pass: the statically scheduled GEMM would
 * otherwise wait a whole extra wave for the SMs NCCL occupies. */
int mcan_set_sm_limit(int sms);
/* Tile schedule of the persistent GEMM.  0 (default): static round robin -- no atomics on the
 * critical path, best when the GEMM owns the GPU.  1: dynamic -- CTAs claim work units from a
 * global counter, so an SM that a co-running kernel (the overlapped NCCL all-reduce of
 * data-parallel training) holds simply claims fewer units instead of stalling a whole wave. */
int mcan_set_gemm_schedule(int dynamic);
/* Programmatic dependent launch between consecutive kernels of the library (default on): the next
 * kernel's launch latency and prologue overlap the previous kernel's tail; every kernel waits for
 * its predecessors' completion (griddepcontrol.wait) before touching global memory. */
int mcan_set_pdl(int enabled);
/* Attention implementation for the image-side problems (head_dim 64, 49..128 query rows, 33..128 keys; everything
 * else always runs on the mma.sync kernels).  1 (default): tcgen05 / TMEM kernels (csrc/attention_tc.cu);
 * 0: mma.sync kernels (csrc/attention.cu).  Same results contract (dropout mask, masking semantics); the
 * environment variable MCAN_ATTN_TC=0 selects 0 at load time. */
int mcan_set_attn_impl(int tcgen05);

/* -- G1/G2/G3: tcgen05 GEMM with fused epilogue --------------------------------------
 * D[M,N] = epilogue( sum_{s<num_seg} A_s[M,K] * B_s[N,K]^T )         (fp32 accumulate in TMEM)
 *
 * Replaces nn.Linear forward (mca.py:33,40,47,61; net_utils.py:26,45; net.py:39,53) and the
 * autograd dgrad / wgrad GEMMs of the same layers.
 *
 *   a_layout 0: A_s is stored [M,K] row-major (lda >= K)     -- activations, forward/dgrad
 *   a_layout 1: A_s is stored [K,M] row-major (lda >= M)     -- wgrad: A = dY^T, stored as dY
 *   b_layout 0: B_s is stored [N,K] row-major (ldb >= K)     -- nn.Linear weight (out,in)
 *   b_layout 1: B_s is stored [K,N] row-major (ldb >= N)     -- dgrad: B = W^T stored as W;
 *                                                               wgrad: B = X^T stored as X
 * num_seg > 1 sums several operand pairs into one accumulator (split-precision
 * "bf16x3": (A_hi,B_hi) + (A_hi,B_lo) + (A_lo,B_hi)).
 *
 * Epilogue, applied in this order to v = acc:
 *   v += bias[n]                         (bias != NULL)
 *   v  = max(v, 0)                       (relu)
 *   v  = keep(m*N+n) ? v/(1-p) : 0       (dropout_p > 0)
 *   v  = gate[m,n] > 0 ? v*gate_scale : 0  (gate != NULL; backward through ReLU+dropout)
 *   v += resid[m,n]                      (resid != NULL, fp32)
 *   colsum[n] += sum_m v                 (colsum != NULL, fp32 atomics into a zeroed / running buffer: the
 *                                         bias gradient of the layer whose input gradient this GEMM computes;
 *                                         only with a bf16-only output and no residual / accumulate)
 *   out_f32[m,n] = v / out_bf16[m,n] = bf16(v) / out_bf16_lo[m,n] = bf16(v - bf16(v))
 * accumulate != 0: out_f32[m,n] += v with fp32 atomics (required for split_k > 1).  Only out_f32
 * may be set and no ReLU: the remaining stages are linear in the accumulator, so with K splits every
 * split scales its partial sum (dropout, gate) and split 0 alone adds bias and resid --
 * out = resid + keep/(1-p) * (sum_s acc_s + bias) accumulates correctly into a ZEROED out_f32.
 * This is how the short-M (question-side) GEMMs with K >= 2048 use all SMs.
 * Alignment: operand base pointers 16 B, leading dimensions multiples of 8 elements.
 */
typedef struct mcan_gemm_args {
    const void* a[MCAN_MAX_GEMM_SEGMENTS];
    const void* b[MCAN_MAX_GEMM_SEGMENTS];
    int32_t num_seg;
    int32_t a_layout;
    int32_t b_layout;
    int64_t m, n, k;
    int64_t lda, ldb;

    const float* bias;
    int32_t relu;
    float dropout_p;
    uint32_t dropout_seed;
    const uint32_t* dropout_seed_dev; /* optional device word XOR-ed into dropout_seed (CUDA graphs) */
    const void* gate;
    int64_t ldg;
    float gate_scale;
    const float* resid;
    int64_t ldr;

    float* out_f32;
    int64_t ldo_f32;
    void* out_bf16;
    void* out_bf16_lo;
    int64_t ldo_bf16;
    float* colsum; /* optional fp32 [N]: += column sums of the epilogue output */

    int32_t accumulate;
    int32_t split_k; /* 0 = choose automatically (only when accumulate != 0) */
    int32_t block_n; /* 0 = choose automatically, else 128 or 256 */
    int32_t cta_group; /* 0 = choose automatically, 1 = one CTA per tile, 2 = CTA pair (256-row tiles) */
    void* stream;
} mcan_gemm_args;

int mcan_gemm(const mcan_gemm_args* args);

/* Grouped weight-gradient GEMM: up to MCAN_MAX_GEMM_GROUPS independent problems that share the contraction length k,
 *   out_g[M_g, N_g] += A_g^T B_g,   A_g = bf16 [k, M_g] (lda), B_g = bf16 [k, N_g] (ldb), out_g fp32 (ldo),
 * as ONE persistent launch (the wgrads dW = dY^T X of one MCAN layer: mca.py:33-61, net_utils.py:26,45 backward).
 * accumulate != 0: same arithmetic as one mcan_gemm(a_layout=1, b_layout=1, accumulate=1) per group -- fp32 atomics
 * into out_g, which the caller zero-initialises; split_k: 0 = choose automatically.
 * accumulate == 0: out_g = A_g^T B_g, plain stores by exactly one work unit per element (K is not split): the
 * outputs need no initialisation -- no zero-fill pass and no read-modify-write of the gradients. */
#define MCAN_MAX_GEMM_GROUPS 8
typedef struct mcan_gemm_group {
    const void* a;
    const void* b;
    int64_t m, n, lda, ldb;
    void* out;            /* fp32 [M, N], or bf16 when mcan_gemm_grouped_args.out_bf16 != 0 */
    int64_t ldo;
} mcan_gemm_group;

typedef struct mcan_gemm_grouped_args {
    mcan_gemm_group g[MCAN_MAX_GEMM_GROUPS];
    int32_t num_groups;
    int32_t split_k;
    int32_t accumulate;
    int32_t out_bf16;     /* != 0: the outputs are bf16 (accumulate must be 0): gradients produced directly in the
                           * buffer a bf16 gradient all-reduce works on */
    int64_t k;
    void* stream;
} mcan_gemm_grouped_args;

int mcan_gemm_grouped(const mcan_gemm_grouped_args* args);

/* -- G2: GEMM with residual add + MCAN LayerNorm fused into the epilogue ----------------------------
 * Replaces, for the sub-layer outputs of SA / SGA (mca.py:119-125, 152-162), the chain
 *   GEMM[+bias, dropout, +resid] -> s -> LayerNorm(s) (net_utils.py:56-60):
 *   s[m,:] = resid[m,:] + dropout(A[m,:] W^T + bias)          A: bf16 [M,K] (lda), W: bf16 [N,K] (ldb)
 *   y[m,:] = ln_a2 * (s - mean) / (std_unbiased + eps) + ln_b2
 * One thread-block cluster owns 256 complete rows (N = 512: one CTA pair; N = 1024: two pairs that share the A
 * tile by TMA multicast and exchange the row statistics through distributed shared memory), so N must be 512
 * or 1024.  Outputs (all row-major with leading dimension N): s_f32 (optional; the LayerNorm input, saved for the
 * backward kernel), y_f32 / y_bf16 (at least one), mean / sigma [M] (optional).  Dropout uses the same
 * (seed, m*N+n) hash as mcan_gemm, so the backward kernels regenerate the mask. */
typedef struct mcan_gemm_ln_args {
    const void* a;
    const void* b;
    int64_t m, n, k;
    int64_t lda, ldb;
    const float* bias;
    float dropout_p;
    uint32_t dropout_seed;
    const uint32_t* dropout_seed_dev;
    const float* resid;
    int64_t ldr;
    const float* ln_a2;
    const float* ln_b2;
    float eps;
    float* s_f32;
    float* y_f32;
    void* y_bf16;
    float* mean;
    float* sigma;
    void* stream;
} mcan_gemm_ln_args;

int mcan_gemm_ln(const mcan_gemm_ln_args* args);

/* -- A1: fused masked-softmax attention, one CTA per (batch, head) -------------------
 * Replaces MHAtt.att (mca.py:65-78) plus the head split/merge transposes (mca.py:33-59):
 *   P = softmax(masked_fill(Q K^T * scale, key_mask, -1e9)); P = dropout(P); O = P V
 * q/k/v/out address row (b*S + s) and columns [h*head_dim, (h+1)*head_dim) of bf16
 * matrices with leading dimensions ldq/ldk/ldv/ldo, so a fused [rows,3H] QKV buffer or a
 * cross-layer K/V buffer can be used in place.  key_mask: uint8 [batch, sk], 1 = masked,
 * may be NULL.  A fully masked row yields the uniform 1/sk distribution, like the
 * reference.  sq, sk <= 128; head_dim in {64, 128}.
 */
typedef struct mcan_attn_args {
    const void* q;
    const void* k;
    const void* v;
    int64_t ldq, ldk, ldv;
    const uint8_t* key_mask;
    void* out;
    int64_t ldo;
    int32_t batch, heads, sq, sk, head_dim;
    float scale;
    float dropout_p;
    uint32_t dropout_seed;
    const uint32_t* dropout_seed_dev; /* optional device word XOR-ed into dropout_seed */
    /* split precision (forward only): low-order bf16 halves of q/k/v and of the output, same
     * addressing as the hi pointers; all four or none.  Products become hi*hi + hi*lo + lo*hi. */
    const void* q_lo;
    const void* k_lo;
    const void* v_lo;
    void* out_lo;
    void* stream;
} mcan_attn_args;

int mcan_attn_fwd(const mcan_attn_args* args);

/* backward of the above: recomputes P from Q,K (no probabilities are saved) and the
 * dropout mask from the seed.  dq/dk/dv are bf16 with the same addressing as q/k/v. */
typedef struct mcan_attn_bwd_args {
    mcan_attn_args fwd; /* same q,k,v,key_mask,shape,scale,dropout as the forward call; out unused */
    const void* dout;   /* bf16, [batch*sq, heads*head_dim] */
    int64_t lddo;
    void* dq;
    void* dk;
    void* dv;
    int64_t lddq, lddk, lddv;
    /* optional fp32 [heads * head_dim] each: += column sums over all rows of dq / dk / dv as stored (the bias
     * gradients of linear_q / linear_k / linear_v, mca.py:33-55), fp32 atomics on zero-initialised buffers */
    float* dbq;
    float* dbk;
    float* dbv;
} mcan_attn_bwd_args;

int mcan_attn_bwd(const mcan_attn_bwd_args* args);

/* -- L1: MCAN LayerNorm (net_utils.py:48-60) -----------------------------------------
 * y = a_2 * (x - mean) / (std_unbiased + eps) + b_2 over the last dimension (h).
 * NOT torch.nn.functional.layer_norm: std uses N-1 and eps is added to std.
 * Writes y as fp32 and (optionally) bf16 / bf16-lo copies for the next GEMM, and the row
 * statistics mean[rows], sigma[rows] (sigma = unbiased std, without eps) for backward.
 */
int mcan_layernorm_fwd(const float* x, int64_t rows, int64_t h, const float* a2, const float* b2,
                       float eps, float* y_f32, void* y_bf16, void* y_bf16_lo, float* mean,
                       float* sigma, void* stream);

/* Same, for a row that is the SUM of two inputs (proj_norm(lang_feat + img_feat), net.py:125-126):
 * s = x + x2 is normalised; s_out (optional) receives s for the backward pass. */
int mcan_layernorm_add_fwd(const float* x, const float* x2, float* s_out, int64_t rows, int64_t h,
                           const float* a2, const float* b2, float eps, float* y_f32, void* y_bf16,
                           void* y_bf16_lo, float* mean, float* sigma, void* stream);

/* dx = LayerNorm backward, plus:
 *   dx_bf16 (optional) = bf16( dropout_mask(seed)[r,c] * dx / (1-p) )  -- the gradient that flows
 *       into the GEMM whose epilogue produced x = resid + dropout(gemm), ready as a GEMM operand;
 *   da2[h] += sum_r dy*c/s, db2[h] += sum_r dy, dbias[h] += sum_r dx_bf16 value (optional) -- atomics.
 */
int mcan_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* sigma,
                       const float* a2, float eps, int64_t rows, int64_t h, float* dx_f32,
                       void* dx_bf16, float dropout_p, uint32_t dropout_seed,
                       const uint32_t* dropout_seed_dev, float* da2, float* db2, float* dbias,
                       void* stream);

/* -- F1: AttFlat pooling (net.py:38-55) ----------------------------------------------
 * Input hmid = dropout(relu(x W1^T + b1)) comes from mcan_gemm (bf16 [batch*s, mlp]).
 *   logit[b,s,g] = hmid[b,s,:] . w2[g,:] + b2[g];  masked_fill(mask, -1e9)
 *   att_w = softmax over s;  pooled[b, g*h : (g+1)*h] = sum_s att_w[b,s,g] * x[b,s,:]
 * pooled is written as fp32 (the backward reads it) and bf16 (operand of linear_merge).  Grid = (slices, batch):
 * a sample is worked on by up to 8 CTAs (column slices of the pooled sums) so that a batch of 64 fills 148 SMs.
 * hmid_lo (optional): low-order bf16 half of hmid (split precision).
 */
int mcan_attflat_pool_fwd(const void* hmid, const void* hmid_lo, const float* w2, const float* b2,
                          const uint8_t* mask, const float* x, int32_t batch, int32_t s, int32_t h,
                          int32_t mlp, int32_t glimpses, float* att_w, float* pooled_f32,
                          void* pooled_bf16, void* stream);

/* backward: dpooled fp32 [batch, g*h], pooled = the forward's fp32 output ->
 *   dx[b,s,:]  = sum_g att_w[b,s,g] * dpooled[b,g,:]                       (fp32, overwritten)
 *   dlogit     = softmax backward of d att_w[b,s,g] = dpooled[b,g,:] . x[b,s,:]; 0 where masked
 *   dhmid      = bf16( (hmid > 0) * gate_scale * sum_g dlogit[.,g] * w2[g,:] )  (through ReLU+dropout)
 *   dw2 += dlogit^T hmid, db2 += sum dlogit                                 (fp32 atomics)
 */
int mcan_attflat_pool_bwd(const float* dpooled, const float* pooled, const void* hmid, const float* w2,
                          const uint8_t* mask, const float* x, const float* att_w, int32_t batch,
                          int32_t s, int32_t h, int32_t mlp, int32_t glimpses, float gate_scale,
                          float* dx, void* dhmid, float* dw2, float* db2, void* stream);

/* -- the two ends of the path ---------------------------------------
 * Image-side input (net.py:100,107,135-137): one pass over the fp32 region features x [rows, cols]
 * writes the bf16 operand of img_feat_linear (hi [+ lo], leading dimension ld >= cols) AND
 * mask[row] = 1 iff every feature of the row is zero  (== (sum(|x|, -1) == 0) of make_mask). */
int mcan_rowmask_cast(const float* x, int64_t rows, int64_t cols, void* hi, void* lo, int64_t ld,
                      uint8_t* mask, void* stream);

/* Output head (net.py:127-129, exec.py:67,178): probs = sigmoid(logits) (logits fp32 [rows, ld], probs
 * contiguous [rows, cols]); with target != NULL also *loss = BCELoss(reduction='sum')(probs, target),
 * torch semantics (log terms clamped at -100), summed in a fixed order (bit-reproducible).
 * workspace: >= 1025 floats of device memory, zero before the first call (the kernel leaves it ready for
 * the next call); needed only with loss. */
int mcan_sigmoid_bce_fwd(const float* logits, int64_t ld, const float* target, int32_t rows, int32_t cols,
                         float* probs, float* loss, float* workspace, void* stream);

/* Backward into the bf16 operand dz [rows, ld] (pad columns zeroed) of the proj wgrad / dgrad GEMMs, and
 * dbias[c] += sum_r dz[r, c] (optional; fp32 atomics on a zeroed buffer).
 *   target != NULL: dz = g (p - t) / max(p (1-p), 1e-12) * p (1-p),  g = *gscale_dev (1 if NULL)  -- BCE(sum) + sigmoid
 *   gout   != NULL: dz = gout * p (1-p)                                                        -- sigmoid only
 * Exactly one of target / gout is given. */
int mcan_sigmoid_bce_bwd(const float* probs, const float* target, const float* gout, const float* gscale_dev,
                         int32_t rows, int32_t cols, void* dz_bf16, int64_t ld, float* dbias, void* stream);

/* -- question encoder (net.py:66-78, 96-104): embedding + single-layer LSTM ---------------------
 * Row layout of every per-token buffer below: row(b, s) = b * (steps + 1) + s, one spare slot per sample
 * (R = batch * (steps + 1) rows), so that hbuf[b, s] = h_{s-1} sits at the row of dA[b, s = t] and x[b, s = t] and the
 * weight gradients dW_hh = dA^T hbuf, dW_ih = dA^T x are plain GEMMs over R rows.
 *
 * mcan_embed_gather: x[row(b, t), :embed] = bf16(table[tokens[b, t]]) (pad columns and slot `steps` zero);
 *   x_bf16_lo (optional, same layout) = bf16(value - x), for a split-precision input projection;
 *   mask[b * steps + t] = (tokens[b, t] == 0)  -- make_mask(ques_ix), net.py:99,135-137.  mask may be NULL.
 * mcan_embed_scatter_add: dtable[tokens[b, t], :] += dx[row(b, t), :embed] (fp32 atomics, dtable zero-initialised). */
int mcan_embed_gather(const int64_t* tokens, const float* table, int32_t vocab, int32_t embed, int32_t batch,
                      int32_t steps, void* x_bf16, void* x_bf16_lo, int32_t ldx, uint8_t* mask, void* stream);
int mcan_embed_scatter_add(const int64_t* tokens, const float* dx, int32_t lddx, int32_t vocab, int32_t embed,
                           int32_t batch, int32_t steps, float* dtable, void* stream);

/* nn.LSTM(num_layers=1, batch_first=True), zero initial state, gate order (i, f, g, o), as ONE persistent kernel per
 * direction of differentiation (all time steps; W_hh slices resident in registers; grid barrier between steps).
 * forward : xw = x W_ih^T + b_ih (fp32 [R, 4H], from mcan_gemm) -> hbuf (bf16 [R, H], slot s = h_{s-1}), h_out (fp32
 *           [batch * steps, H], row b * steps + t: the module output), and for training cbuf (fp32 [R, H]) and gates
 *           (fp32 [R, 4H], activated) -- both NULL for inference.
 * backward: dout (fp32 [batch * steps, H]) + gates, cbuf, w_hh -> da (bf16 [R, 4H]: gradients of the gate
 *           pre-activations, slot `steps` zero), from which dW_ih, dW_hh, the bias gradients and dx follow as GEMMs.
 * batch <= 64 per launch (the host splits larger batches); hidden in {128, 256, 512, 1024}; needs >= 128 SMs.
 * barrier: 2 words of device memory, zero before the first launch (the kernels leave them zero). */
typedef struct mcan_lstm_args {
    const float* xw;
    const void* w_hh;     /* bf16 [4H, H] */
    const float* b_hh;    /* fp32 [4H] */
    int32_t batch, steps, hidden;
    void* hbuf;
    float* h_out;
    float* cbuf;
    float* gates;
    const float* dout;
    void* da;
    uint32_t* barrier;
    void* stream;
} mcan_lstm_args;

int mcan_lstm_fwd(const mcan_lstm_args* args);
int mcan_lstm_bwd(const mcan_lstm_args* args);

/* -- small memory-bound helpers -------------------------------------------------------- */
/* hi = bf16(x); lo (optional) = bf16(x - hi).  n elements. */
int mcan_cast_bf16(const float* x, int64_t n, void* hi, void* lo, void* stream);
/* Multi-tensor refresh of GEMM-operand copies in ONE launch.  seg_table_dev: device array of
 * num_segments entries {const float* src; void* dst; int64 n; int64 first_chunk; void* dst_lo},
 * segments laid out back to back in 4096-element chunks (first_chunk = running chunk index; bit 62
 * set means dst is fp32 (plain copy, used to concatenate biases), otherwise dst is bf16 and dst_lo,
 * if not NULL, receives bf16(x - bf16(x)) for the split-precision mode). */
int mcan_cast_multi(const void* seg_table_dev, int32_t num_segments, int64_t total_chunks, void* stream);
/* out = bf16(act > 0 ? dy * scale : 0): gradient through FC's ReLU + dropout (net_utils.py:28-32)
 * from the saved bf16 activation; n contiguous elements. */
int mcan_gate_bf16(const float* dy, const void* act, float scale, void* out, int64_t n, void* stream);
/* out[c] += sum_r x[r,c]   (x bf16 [rows, cols] with leading dimension ld; fp32 atomics) */
int mcan_colsum_bf16(const void* x, int64_t rows, int64_t cols, int64_t ld, float* out,
                     void* stream);
/* out[c] += sum_r x[r,c]   (x fp32) */
int mcan_colsum_f32(const float* x, int64_t rows, int64_t cols, int64_t ld, float* out,
                    void* stream);

/* Launch plan mcan_gemm would use for this shape on `sms` SMs (0 = the current device): pure host logic,
 * callable without a GPU.  plan_out[7] = {block_n, CTAs per cluster, m_tiles, n_tiles, K splits,
 * full_tiles (tiles issued at full width; the remaining ones as two half-width units each), work units}. */
int mcan_gemm_plan(int64_t m, int64_t n, int64_t k, int32_t accumulate, int32_t split_k, int32_t block_n,
                   int32_t cta_group, int32_t sms, int32_t* plan_out);

/* -- fused multi-tensor AdamW (reference core/model/optim.py:58-64) -----------------
 * One launch updates every fp32 master parameter with torch.optim.AdamW semantics (decoupled weight
 * decay, bias correction with the step count t) and re-emits, in the same pass, the operand copy
 * the next forward's GEMMs read.  seg_table_dev: device array of `num_segments` records of eight
 * 64-bit words {p, g, m, v, shadow, n, first_chunk, shadow2}: fp32 pointers p (in/out), g (in), m, v
 * (in/out); shadow / shadow2 = optional copies of the updated p (bf16, or fp32 when bit 62 / bit 61 of
 * first_chunk is set; 0 = none) -- two, because a weight can sit in two stacked operand buffers;
 * n elements; first_chunk (bits 0..59) = running index of the segment's first 4096-element chunk; bit 60 set = g
 * points to bf16 values (the all-reduced bf16 staging buffer of a compressed gradient exchange).
 * lr_dev / step_dev: device scalars (fp32) holding the learning rate and t >= 1, so a captured CUDA
 * graph replays with new values.  flags bit 0: one chunk per CTA (short-lived CTAs) instead of a persistent
 * grid, for an update that shares the GPU with latency-bound kernels of a higher-priority stream. */
int mcan_adamw_multi(const void* seg_table_dev, int32_t num_segments, int64_t total_chunks,
                     const float* lr_dev, const float* step_dev, float beta1, float beta2, float eps,
                     float weight_decay, int32_t flags, void* stream);

/* experiment helper (tools/contention_bench.py): keep `ctas` SMs busy for `cycles` clocks */
int mcan_debug_hog(int32_t ctas, int64_t cycles, int32_t smem_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MCAN_B200_H_ */

// Shared declarations of the two attention implementations (attention.cu: mma.sync, one warp per 16-query tile;
// attention_tc.cu: tcgen05 / TMEM, one thread per query row).
#pragma once
#include "common.cuh"

namespace mcan {

constexpr int kAttnMaxSeq = 128;
constexpr int kAttnMaxNT = kAttnMaxSeq / 8;  // n-tiles of 8 keys
constexpr float kLog2e = 1.4426950408889634f;

struct AttnParams {
    const bf16* q;
    const bf16* k;
    const bf16* v;
    const bf16* q_lo;   // split precision ("bf16x3"): low-order halves, same addressing as q/k/v/out
    const bf16* k_lo;
    const bf16* v_lo;
    bf16* out_lo;
    long long ldq, ldk, ldv;
    const uint8_t* mask;
    bf16* out;
    long long ldo;
    int batch, heads, sq, sk;
    float scale;
    uint32_t drop_thr;
    float drop_scale;
    uint32_t drop_seed;
    const uint32_t* drop_seed_dev;
    // backward only
    const bf16* dout;
    long long lddo;
    bf16* dq;
    bf16* dk;
    bf16* dv;
    long long lddq, lddk, lddv;
    float* dbq;     // optional: += column sums of dq / dk / dv (the bias gradients of linear_q / _k / _v), fp32 [heads * D]
    float* dbk;
    float* dbv;
    int prefetch;   // backward: persistent grid, the operand tiles of the CTA's NEXT (batch, head) are staged into a
                    // second shared-memory set while the current one is being processed
};

// tcgen05 / TMEM path (attention_tc.cu): head_dim 64, query tiles of 49..128 rows (the image side of MCAN).
// *_eligible say whether a problem is taken by that path; the launchers return 0 / an error code.
bool attn_tc_fwd_eligible(const AttnParams& p, int head_dim);
bool attn_tc_bwd_eligible(const AttnParams& p, int head_dim);
int attn_tc_fwd_launch(const AttnParams& p, cudaStream_t st);
int attn_tc_bwd_launch(const AttnParams& p, cudaStream_t st);

}  // namespace mcan

// A1: fused masked-softmax attention for MCAN's short sequences (<= 128 tokens / regions).
//
// One CTA per (batch, head).  Q, K, V (and dO in backward) of that head are staged once in
// shared memory; QK^T, the reference's masked_fill(-1e9) + softmax (mca.py:68-75), dropout
// (mca.py:76) and P.V (mca.py:78) run on warp-level tensor-core MMAs with the score tile held
// in registers -- the [B,h,Sq,Sk] score tensor of the reference never exists in HBM, and the
// head split / merge transposes (mca.py:33-59) become pointer arithmetic on the [rows, H]
// activations.  These kernels use mma.sync m16n8k16 on tiles that fit one warp; they serve the small
// problems (question side, question-guided attention with <= 32 keys), head dim 128, split precision and
// the fused bias gradients.  The image self-attention (head dim 64, 49..128 queries, 33..128 keys: one
// M = 128 UMMA tile per (batch, head)) is dispatched to the tcgen05 / TMEM kernels of attention_tc.cu.
//
// Backward recomputes P from Q,K and the dropout mask from the seed, then
//   dV = Pd^T dO,  dPd = dO V^T,  dS = P o (m/(1-p) dPd - rowsum(Pd o dPd)), 0 where masked,
//   dQ = scale dS K,  dK = scale dS^T Q.
#include "../../include/mcan_b200.h"
#include "attention.cuh"
#include <stdlib.h>

namespace mcan {

int device_num_sms();

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                        uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                          uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2,
                                         uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
        "{%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc)
                 : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// Stage `rows` x D bf16 (global row stride ld) into smem rows of stride D+8; rows up to
// rows_pad are zero-filled.
template <int D>
__device__ __forceinline__ void stage_rows(bf16* s, const bf16* g, long long ld, int rows,
                                           int rows_pad) {
    constexpr int CPR = D / 8;  // 16-byte chunks per row
    constexpr int LDS = D + 8;
    for (int i = threadIdx.x; i < rows_pad * CPR; i += blockDim.x) {
        const int r = i / CPR, c = i % CPR;
        bf16* dst = s + r * LDS + c * 8;
        if (r < rows)
            cp_async16(dst, g + (long long)r * ld + c * 8);
        else
            *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
}

// S = Q_tile K^T for one 16-query tile.  acc[nt] covers keys nt*8..nt*8+7.
template <int D, int NT, bool ZERO = true>
__device__ __forceinline__ void qk_tile(const bf16* sQ, const bf16* sK, int mt, int nkt, int lane,
                                        float (&acc)[NT][4]) {
    constexpr int LDS = D + 8;
    uint32_t qf[D / 16][4];
    {
        const int mi = lane >> 3, r = lane & 7;
        const bf16* base = sQ + (mt * 16 + (mi & 1) * 8 + r) * LDS + (mi >> 1) * 8;
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk)
            ldsm_x4(smem_u32(base + kk * 16), qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3]);
    }
    if (ZERO) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[nt][j] = 0.f;
    }
    const int mi = lane >> 3, r = lane & 7;
#pragma unroll
    for (int np = 0; np < NT / 2; ++np) {
        if (np * 2 < nkt) {  // warp-uniform
            const bf16* kb = sK + (np * 16 + (mi >> 1) * 8 + r) * LDS + (mi & 1) * 8;
#pragma unroll
            for (int kk = 0; kk < D / 16; ++kk) {
                uint32_t b0, b1, b2, b3;
                ldsm_x4(smem_u32(kb + kk * 16), b0, b1, b2, b3);
                mma_bf16(acc[np * 2], qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3], b0, b1);
                mma_bf16(acc[np * 2 + 1], qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3], b2, b3);
            }
        }
    }
}

// Per-thread key bitmaps: bit (2*nt + j) describes key nt*8 + 2*(lane&3) + j.
__device__ __forceinline__ void key_bits(const uint8_t* sMask, int nkt, int sk, int lane,
                                         uint32_t& masked, uint32_t& valid) {
    const int t = lane & 3;
    masked = 0;
    valid = 0;
    for (int nt = 0; nt < nkt; ++nt)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int key = nt * 8 + 2 * t + j;
            if (key < sk) {
                valid |= 1U << (2 * nt + j);
                if (sMask[key]) masked |= 1U << (2 * nt + j);
            }
        }
}

// keep-scale factors of two adjacent elements idx, idx+1 (one hash when idx is even)
__device__ __forceinline__ void dropout_pair(uint32_t idx, uint32_t seed, uint32_t thr, float scale,
                                             float& k0, float& k1) {
    uint32_t u0, u1;
    if ((idx & 1U) == 0) {
        const uint32_t r = dropout_bits_pair(idx >> 1, seed);
        u0 = r & 0xFFFFU;
        u1 = r >> 16;
    } else {
        u0 = dropout_u16(idx, seed);
        u1 = dropout_u16(idx + 1, seed);
    }
    k0 = u0 >= thr ? scale : 0.f;
    k1 = u1 >= thr ? scale : 0.f;
}

// In-register masked softmax of the score tile (rows g and g+8 of this thread's quad).
// Matches: scores/sqrt(d) -> masked_fill(mask, -1e9) -> softmax (mca.py:68-75), evaluated in the
// log2 domain: p = exp2(s*log2e - max).
template <int NT>
__device__ __forceinline__ void softmax_tile(float (&acc)[NT][4], uint32_t masked,
                                             uint32_t valid, int nkt, float scale) {
    const float c = scale * kLog2e;
    const float kMasked = -1e9f * kLog2e;
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        if (nt < nkt) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const uint32_t bit = 1U << (2 * nt + j);
                float s0 = acc[nt][j] * c, s1 = acc[nt][2 + j] * c;
                if (!(valid & bit)) {
                    s0 = -INFINITY;
                    s1 = -INFINITY;
                } else if (masked & bit) {
                    s0 = kMasked;
                    s1 = kMasked;
                }
                acc[nt][j] = s0;
                acc[nt][2 + j] = s1;
                mx0 = fmaxf(mx0, s0);
                mx1 = fmaxf(mx1, s1);
            }
        }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffU, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffU, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffU, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffU, mx1, 2));
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        if (nt < nkt) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float e0 = exp2f(acc[nt][j] - mx0);
                const float e1 = exp2f(acc[nt][2 + j] - mx1);
                acc[nt][j] = e0;
                acc[nt][2 + j] = e1;
                sum0 += e0;
                sum1 += e1;
            }
        }
    }
    sum0 += __shfl_xor_sync(0xffffffffU, sum0, 1);
    sum0 += __shfl_xor_sync(0xffffffffU, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffU, sum1, 1);
    sum1 += __shfl_xor_sync(0xffffffffU, sum1, 2);
    const float inv0 = 1.f / sum0, inv1 = 1.f / sum1;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        if (nt < nkt) {
            acc[nt][0] *= inv0;
            acc[nt][1] *= inv0;
            acc[nt][2] *= inv1;
            acc[nt][3] *= inv1;
        }
    }
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
// NT = compile-time number of 8-key tiles; EXACT: the runtime tile count equals NT (no guards).
template <int D, int NT, bool EXACT>
__global__ void __launch_bounds__(128) attn_fwd_kernel(const AttnParams p) {
    pdl_launch_dependents();
    pdl_wait();
    constexpr int LDS = D + 8;
    extern __shared__ __align__(16) uint8_t smem_attn[];
    const int b = blockIdx.x / p.heads, h = blockIdx.x % p.heads;
    const int sqp = (p.sq + 15) & ~15, skp = (p.sk + 15) & ~15;
    bf16* sQ = reinterpret_cast<bf16*>(smem_attn);
    bf16* sK = sQ + sqp * LDS;
    bf16* sV = sK + skp * LDS;
    uint8_t* sMask = reinterpret_cast<uint8_t*>(sV + skp * LDS);

    stage_rows<D>(sQ, p.q + (long long)b * p.sq * p.ldq + h * D, p.ldq, p.sq, sqp);
    stage_rows<D>(sK, p.k + (long long)b * p.sk * p.ldk + h * D, p.ldk, p.sk, skp);
    stage_rows<D>(sV, p.v + (long long)b * p.sk * p.ldv + h * D, p.ldv, p.sk, skp);
    for (int i = threadIdx.x; i < skp; i += blockDim.x)
        sMask[i] = (p.mask != nullptr && i < p.sk) ? p.mask[(long long)b * p.sk + i] : 0;
    cp_async_wait_all();
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = blockDim.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int nkt = EXACT ? NT : skp / 8;
    const uint32_t drop_seed =
        p.drop_seed ^ ((p.drop_thr != 0 && p.drop_seed_dev != nullptr) ? __ldg(p.drop_seed_dev) : 0U);
    uint32_t kmasked, kvalid;
    key_bits(sMask, nkt, p.sk, lane, kmasked, kvalid);

    for (int mt = warp; mt < sqp / 16; mt += nwarps) {
        float acc[NT][4];
        qk_tile<D, NT>(sQ, sK, mt, nkt, lane, acc);
        softmax_tile(acc, kmasked, kvalid, nkt, p.scale);

        const int row0 = mt * 16 + g, row1 = row0 + 8;
        if (p.drop_thr != 0) {
            const uint32_t base0 = (uint32_t)(((long long)blockIdx.x * p.sq + row0) * p.sk);
            const uint32_t base1 = (uint32_t)(((long long)blockIdx.x * p.sq + row1) * p.sk);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                if (nt < nkt) {
                    const uint32_t key = nt * 8 + 2 * t;
                    float k0, k1;
                    dropout_pair(base0 + key, drop_seed, p.drop_thr, p.drop_scale, k0, k1);
                    acc[nt][0] *= k0;
                    acc[nt][1] *= k1;
                    dropout_pair(base1 + key, drop_seed, p.drop_thr, p.drop_scale, k0, k1);
                    acc[nt][2] *= k0;
                    acc[nt][3] *= k1;
                }
            }
        }

        // O = P V
        float o[D / 8][4];
#pragma unroll
        for (int dn = 0; dn < D / 8; ++dn)
#pragma unroll
            for (int j = 0; j < 4; ++j) o[dn][j] = 0.f;
        const int mi = lane >> 3, r = lane & 7;
#pragma unroll
        for (int ks = 0; ks < NT / 2; ++ks) {
            if (ks * 2 < nkt) {
                const uint32_t a0 = pack_bf16x2(acc[2 * ks][0], acc[2 * ks][1]);
                const uint32_t a1 = pack_bf16x2(acc[2 * ks][2], acc[2 * ks][3]);
                const uint32_t a2 = pack_bf16x2(acc[2 * ks + 1][0], acc[2 * ks + 1][1]);
                const uint32_t a3 = pack_bf16x2(acc[2 * ks + 1][2], acc[2 * ks + 1][3]);
                const bf16* vb = sV + (ks * 16 + (mi & 1) * 8 + r) * LDS + (mi >> 1) * 8;
#pragma unroll
                for (int dp = 0; dp < D / 16; ++dp) {
                    uint32_t b0, b1, b2, b3;
                    ldsm_x4_t(smem_u32(vb + dp * 16), b0, b1, b2, b3);
                    mma_bf16(o[dp * 2], a0, a1, a2, a3, b0, b1);
                    mma_bf16(o[dp * 2 + 1], a0, a1, a2, a3, b2, b3);
                }
            }
        }
        bf16* orow0 = p.out + ((long long)b * p.sq + row0) * p.ldo + h * D + 2 * t;
        bf16* orow1 = p.out + ((long long)b * p.sq + row1) * p.ldo + h * D + 2 * t;
#pragma unroll
        for (int dn = 0; dn < D / 8; ++dn) {
            if (row0 < p.sq) *reinterpret_cast<uint32_t*>(orow0 + dn * 8) = pack_bf16x2(o[dn][0], o[dn][1]);
            if (row1 < p.sq) *reinterpret_cast<uint32_t*>(orow1 + dn * 8) = pack_bf16x2(o[dn][2], o[dn][3]);
        }
    }
}

// ------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------
// forward, split precision: every bf16 operand comes as hi + lo and every product is evaluated as
// hi*hi + hi*lo + lo*hi (fp32 accumulate), which reproduces the fp32 reference to ~2e-5.
// Used by the "fp32" inference mode (north star: 1e-4 logits, identical top-1 answers).
// ------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(128) attn_fwd_split_kernel(const AttnParams p) {
    pdl_launch_dependents();
    pdl_wait();
    constexpr int LDS = D + 8;
    constexpr int NT = kAttnMaxNT;
    extern __shared__ __align__(16) uint8_t smem_attn[];
    const int b = blockIdx.x / p.heads, h = blockIdx.x % p.heads;
    const int sqp = (p.sq + 15) & ~15, skp = (p.sk + 15) & ~15;
    bf16* sQ = reinterpret_cast<bf16*>(smem_attn);
    bf16* sQl = sQ + sqp * LDS;
    bf16* sK = sQl + sqp * LDS;
    bf16* sKl = sK + skp * LDS;
    bf16* sV = sKl + skp * LDS;
    bf16* sVl = sV + skp * LDS;
    uint8_t* sMask = reinterpret_cast<uint8_t*>(sVl + skp * LDS);
    const long long qo = (long long)b * p.sq * p.ldq + h * D;
    const long long ko = (long long)b * p.sk * p.ldk + h * D;
    const long long vo = (long long)b * p.sk * p.ldv + h * D;
    stage_rows<D>(sQ, p.q + qo, p.ldq, p.sq, sqp);
    stage_rows<D>(sQl, p.q_lo + qo, p.ldq, p.sq, sqp);
    stage_rows<D>(sK, p.k + ko, p.ldk, p.sk, skp);
    stage_rows<D>(sKl, p.k_lo + ko, p.ldk, p.sk, skp);
    stage_rows<D>(sV, p.v + vo, p.ldv, p.sk, skp);
    stage_rows<D>(sVl, p.v_lo + vo, p.ldv, p.sk, skp);
    for (int i = threadIdx.x; i < skp; i += blockDim.x)
        sMask[i] = (p.mask != nullptr && i < p.sk) ? p.mask[(long long)b * p.sk + i] : 0;
    cp_async_wait_all();
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = blockDim.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int nkt = skp / 8;
    uint32_t kmasked, kvalid;
    key_bits(sMask, nkt, p.sk, lane, kmasked, kvalid);

    for (int mt = warp; mt < sqp / 16; mt += nwarps) {
        float acc[NT][4];
        qk_tile<D, NT, true>(sQ, sK, mt, nkt, lane, acc);
        qk_tile<D, NT, false>(sQ, sKl, mt, nkt, lane, acc);
        qk_tile<D, NT, false>(sQl, sK, mt, nkt, lane, acc);
        softmax_tile(acc, kmasked, kvalid, nkt, p.scale);

        float o[D / 8][4];
#pragma unroll
        for (int dn = 0; dn < D / 8; ++dn)
#pragma unroll
            for (int j = 0; j < 4; ++j) o[dn][j] = 0.f;
        const int mi = lane >> 3, r = lane & 7;
#pragma unroll
        for (int ks = 0; ks < NT / 2; ++ks) {
            if (ks * 2 < nkt) {
                uint32_t ah[4], al[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float x0 = acc[2 * ks + (q >> 1)][2 * (q & 1)], x1 = acc[2 * ks + (q >> 1)][2 * (q & 1) + 1];
                    ah[q] = pack_bf16x2(x0, x1);
                    al[q] = pack_bf16x2(x0 - bf16_lo_to_f(ah[q]), x1 - bf16_hi_to_f(ah[q]));
                }
                const int voff = (ks * 16 + (mi & 1) * 8 + r) * LDS + (mi >> 1) * 8;
#pragma unroll
                for (int dp = 0; dp < D / 16; ++dp) {
                    uint32_t b0, b1, b2, b3;
                    ldsm_x4_t(smem_u32(sV + voff + dp * 16), b0, b1, b2, b3);
                    mma_bf16(o[dp * 2], ah[0], ah[1], ah[2], ah[3], b0, b1);
                    mma_bf16(o[dp * 2 + 1], ah[0], ah[1], ah[2], ah[3], b2, b3);
                    mma_bf16(o[dp * 2], al[0], al[1], al[2], al[3], b0, b1);
                    mma_bf16(o[dp * 2 + 1], al[0], al[1], al[2], al[3], b2, b3);
                    ldsm_x4_t(smem_u32(sVl + voff + dp * 16), b0, b1, b2, b3);
                    mma_bf16(o[dp * 2], ah[0], ah[1], ah[2], ah[3], b0, b1);
                    mma_bf16(o[dp * 2 + 1], ah[0], ah[1], ah[2], ah[3], b2, b3);
                }
            }
        }
        const int row0 = mt * 16 + g, row1 = row0 + 8;
        const long long o0 = ((long long)b * p.sq + row0) * p.ldo + h * D + 2 * t;
        const long long o1 = ((long long)b * p.sq + row1) * p.ldo + h * D + 2 * t;
#pragma unroll
        for (int dn = 0; dn < D / 8; ++dn) {
            if (row0 < p.sq) {
                const uint32_t hi = pack_bf16x2(o[dn][0], o[dn][1]);
                *reinterpret_cast<uint32_t*>(p.out + o0 + dn * 8) = hi;
                *reinterpret_cast<uint32_t*>(p.out_lo + o0 + dn * 8) =
                    pack_bf16x2(o[dn][0] - bf16_lo_to_f(hi), o[dn][1] - bf16_hi_to_f(hi));
            }
            if (row1 < p.sq) {
                const uint32_t hi = pack_bf16x2(o[dn][2], o[dn][3]);
                *reinterpret_cast<uint32_t*>(p.out + o1 + dn * 8) = hi;
                *reinterpret_cast<uint32_t*>(p.out_lo + o1 + dn * 8) =
                    pack_bf16x2(o[dn][2] - bf16_lo_to_f(hi), o[dn][3] - bf16_hi_to_f(hi));
            }
        }
    }
}

// Column sums of one 16-row accumulator tile o[dn][0..3] (rows g / g + 8, columns dn*8 + 2t, +1), as they are stored
// (bf16-rounded), accumulated into THIS WARP's private shared-memory vector cs[D] (no atomics: shuffles over the 8
// row groups, then the 4 lanes of row group 0 add their 2 x D/8 columns).
template <int D>
__device__ __forceinline__ void tile_colsum(const float (&o)[D / 8][4], bool v0, bool v1, float* cs, int lane) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int dn = 0; dn < D / 8; ++dn) {
        const uint32_t lo = pack_bf16x2(o[dn][0], o[dn][1]), hi = pack_bf16x2(o[dn][2], o[dn][3]);
        float c0 = (v0 ? bf16_lo_to_f(lo) : 0.f) + (v1 ? bf16_lo_to_f(hi) : 0.f);
        float c1 = (v0 ? bf16_hi_to_f(lo) : 0.f) + (v1 ? bf16_hi_to_f(hi) : 0.f);
#pragma unroll
        for (int sh = 4; sh < 32; sh <<= 1) {
            c0 += __shfl_xor_sync(0xffffffffU, c0, sh);
            c1 += __shfl_xor_sync(0xffffffffU, c1, sh);
        }
        if (g == 0) {
            float2* dst = reinterpret_cast<float2*>(cs + dn * 8 + 2 * t);
            float2 v = *dst;
            v.x += c0;
            v.y += c1;
            *dst = v;
        }
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------
template <int D, int NT, bool EXACT>
__global__ void __launch_bounds__(256) attn_bwd_kernel(const AttnParams p) {
    pdl_launch_dependents();
    pdl_wait();
    constexpr int LDS = D + 8;
    extern __shared__ __align__(16) uint8_t smem_attn[];
    const int sqp = (p.sq + 15) & ~15, skp = (p.sk + 15) & ~15;
    const int ldp = skp + 8;
    // operand tiles of one (batch, head): Q, dO [sqp][LDS], K, V [skp][LDS]; two such sets when prefetching
    const int set_elems = (2 * sqp + 2 * skp) * LDS;
    const int nsets = p.prefetch ? 2 : 1;
    bf16* sets = reinterpret_cast<bf16*>(smem_attn);
    bf16* sP = sets + (size_t)nsets * set_elems;   // dropped probabilities Pd   [sqp][ldp]
    bf16* sdS = sP + sqp * ldp;                    // scale * dS                 [sqp][ldp]
    uint8_t* sMaskAll = reinterpret_cast<uint8_t*>(sdS + sqp * ldp);      // [nsets][round16(skp)]
    const int mask_stride = (skp + 15) & ~15;
    // [warp][dq | dk | dv][column], behind the mask bytes (16-byte aligned): column sums of this (batch, head)'s
    // dQ / dK / dV (optional: the bias gradients of the three projections)
    float (*s_cs)[3][D] = reinterpret_cast<float (*)[3][D]>(sMaskAll + nsets * mask_stride);
    const bool want_cs = p.dbq != nullptr || p.dbk != nullptr || p.dbv != nullptr;

    const int items = p.batch * p.heads;
    auto stage_item = [&](int item, int set) {
        const int b = item / p.heads, h = item % p.heads;
        bf16* sQ = sets + (size_t)set * set_elems;
        bf16* sdO = sQ + sqp * LDS;
        bf16* sK = sdO + sqp * LDS;
        bf16* sV = sK + skp * LDS;
        stage_rows<D>(sQ, p.q + (long long)b * p.sq * p.ldq + h * D, p.ldq, p.sq, sqp);
        stage_rows<D>(sdO, p.dout + (long long)b * p.sq * p.lddo + h * D, p.lddo, p.sq, sqp);
        stage_rows<D>(sK, p.k + (long long)b * p.sk * p.ldk + h * D, p.ldk, p.sk, skp);
        stage_rows<D>(sV, p.v + (long long)b * p.sk * p.ldv + h * D, p.ldv, p.sk, skp);
        uint8_t* sm = sMaskAll + set * mask_stride;
        for (int i = threadIdx.x; i < skp; i += blockDim.x)
            sm[i] = (p.mask != nullptr && i < p.sk) ? p.mask[(long long)b * p.sk + i] : 0;
    };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = blockDim.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int mi = lane >> 3, r = lane & 7;
    const int nkt = EXACT ? NT : skp / 8;
    const uint32_t drop_seed =
        p.drop_seed ^ ((p.drop_thr != 0 && p.drop_seed_dev != nullptr) ? __ldg(p.drop_seed_dev) : 0U);

    if ((int)blockIdx.x < items) stage_item(blockIdx.x, 0);
    int it = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
    const int set = p.prefetch ? (it & 1) : 0;
    const int b = item / p.heads, h = item % p.heads;
    bf16* sQ = sets + (size_t)set * set_elems;
    bf16* sdO = sQ + sqp * LDS;
    bf16* sK = sdO + sqp * LDS;
    bf16* sV = sK + skp * LDS;
    const uint8_t* sMask = sMaskAll + set * mask_stride;
    if (!p.prefetch && it > 0) {      // single tile set and more items than CTAs: stage synchronously
        __syncthreads();
        stage_item(item, 0);
    }
    cp_async_wait_all();
    __syncthreads();          // this item's tiles are in place; every warp is done with the previous item
    if (p.prefetch && item + (int)gridDim.x < items) stage_item(item + gridDim.x, set ^ 1);   // in flight during this item
    if (want_cs) {
        for (int i = threadIdx.x; i < 8 * 3 * D; i += blockDim.x) (&s_cs[0][0][0])[i] = 0.f;
        __syncthreads();
    }
    uint32_t kmasked, kvalid;
    key_bits(sMask, nkt, p.sk, lane, kmasked, kvalid);

    // ---- phase 1: per 16-query tile: P, dPd, dS, dQ ----
    for (int mt = warp; mt < sqp / 16; mt += nwarps) {
        float acc[NT][4];
        qk_tile<D, NT>(sQ, sK, mt, nkt, lane, acc);
        softmax_tile(acc, kmasked, kvalid, nkt, p.scale);

        // dPd = dO V^T (same operand pattern as Q K^T)
        float dp[NT][4];
        qk_tile<D, NT>(sdO, sV, mt, nkt, lane, dp);

        const int row0 = mt * 16 + g, row1 = row0 + 8;
        const uint32_t base0 = (uint32_t)(((long long)item * p.sq + row0) * p.sk);
        const uint32_t base1 = (uint32_t)(((long long)item * p.sq + row1) * p.sk);
        const bool v0 = row0 < p.sq, v1 = row1 < p.sq;
        float d0 = 0.f, d1 = 0.f;
        // acc = P (undropped); Pd = keep*P goes to smem; dp <- dP = keep * dPd;
        // D_i = sum_j Pd_ij dPd_ij = sum_j P_ij dP_ij
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            if (nt < nkt) {
                float pd[4], keep[4] = {1.f, 1.f, 1.f, 1.f};
                if (p.drop_thr != 0) {
                    const uint32_t key = nt * 8 + 2 * t;
                    dropout_pair(base0 + key, drop_seed, p.drop_thr, p.drop_scale, keep[0], keep[1]);
                    dropout_pair(base1 + key, drop_seed, p.drop_thr, p.drop_scale, keep[2], keep[3]);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    pd[j] = acc[nt][j] * keep[j];
                    dp[nt][j] *= keep[j];
                    if (j < 2) d0 += acc[nt][j] * dp[nt][j]; else d1 += acc[nt][j] * dp[nt][j];
                }
                const int col = nt * 8 + 2 * t;
                *reinterpret_cast<uint32_t*>(sP + row0 * ldp + col) = v0 ? pack_bf16x2(pd[0], pd[1]) : 0U;
                *reinterpret_cast<uint32_t*>(sP + row1 * ldp + col) = v1 ? pack_bf16x2(pd[2], pd[3]) : 0U;
            }
        }
        d0 += __shfl_xor_sync(0xffffffffU, d0, 1);
        d0 += __shfl_xor_sync(0xffffffffU, d0, 2);
        d1 += __shfl_xor_sync(0xffffffffU, d1, 1);
        d1 += __shfl_xor_sync(0xffffffffU, d1, 2);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            if (nt < nkt) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t bit = 1U << (2 * nt + (j & 1));
                    const bool dead = !(kvalid & bit) || (kmasked & bit) || !((j < 2) ? v0 : v1);
                    dp[nt][j] = dead ? 0.f : acc[nt][j] * (dp[nt][j] - ((j < 2) ? d0 : d1)) * p.scale;
                }
                const int col = nt * 8 + 2 * t;
                *reinterpret_cast<uint32_t*>(sdS + row0 * ldp + col) = pack_bf16x2(dp[nt][0], dp[nt][1]);
                *reinterpret_cast<uint32_t*>(sdS + row1 * ldp + col) = pack_bf16x2(dp[nt][2], dp[nt][3]);
            }
        }
        // dQ = (scale dS) K
        float o[D / 8][4];
#pragma unroll
        for (int dn = 0; dn < D / 8; ++dn)
#pragma unroll
            for (int j = 0; j < 4; ++j) o[dn][j] = 0.f;
#pragma unroll
        for (int ks = 0; ks < NT / 2; ++ks) {
            if (ks * 2 < nkt) {
                const uint32_t a0 = pack_bf16x2(dp[2 * ks][0], dp[2 * ks][1]);
                const uint32_t a1 = pack_bf16x2(dp[2 * ks][2], dp[2 * ks][3]);
                const uint32_t a2 = pack_bf16x2(dp[2 * ks + 1][0], dp[2 * ks + 1][1]);
                const uint32_t a3 = pack_bf16x2(dp[2 * ks + 1][2], dp[2 * ks + 1][3]);
                const bf16* kb = sK + (ks * 16 + (mi & 1) * 8 + r) * LDS + (mi >> 1) * 8;
#pragma unroll
                for (int dpp = 0; dpp < D / 16; ++dpp) {
                    uint32_t b0, b1, b2, b3;
                    ldsm_x4_t(smem_u32(kb + dpp * 16), b0, b1, b2, b3);
                    mma_bf16(o[dpp * 2], a0, a1, a2, a3, b0, b1);
                    mma_bf16(o[dpp * 2 + 1], a0, a1, a2, a3, b2, b3);
                }
            }
        }
        bf16* q0 = p.dq + ((long long)b * p.sq + row0) * p.lddq + h * D + 2 * t;
        bf16* q1 = p.dq + ((long long)b * p.sq + row1) * p.lddq + h * D + 2 * t;
#pragma unroll
        for (int dn = 0; dn < D / 8; ++dn) {
            if (v0) *reinterpret_cast<uint32_t*>(q0 + dn * 8) = pack_bf16x2(o[dn][0], o[dn][1]);
            if (v1) *reinterpret_cast<uint32_t*>(q1 + dn * 8) = pack_bf16x2(o[dn][2], o[dn][3]);
        }
        if (p.dbq != nullptr) tile_colsum<D>(o, v0, v1, s_cs[warp][0], lane);
    }
    __syncthreads();

    // ---- phase 2: per 16-key tile: dV = Pd^T dO, dK = (scale dS)^T Q ----
    const int nqs = sqp / 16;
    for (int kt = warp; kt < skp / 16; kt += nwarps) {
        float ov[D / 8][4], ok[D / 8][4];
#pragma unroll
        for (int dn = 0; dn < D / 8; ++dn)
#pragma unroll
            for (int j = 0; j < 4; ++j) { ov[dn][j] = 0.f; ok[dn][j] = 0.f; }
        for (int qs = 0; qs < nqs; ++qs) {
            // A[m=key, k=query] read transposed from the [query][key] tiles
            const int arow = qs * 16 + (mi >> 1) * 8 + r;
            const int acol = kt * 16 + (mi & 1) * 8;
            uint32_t pa0, pa1, pa2, pa3, sa0, sa1, sa2, sa3;
            ldsm_x4_t(smem_u32(sP + arow * ldp + acol), pa0, pa1, pa2, pa3);
            ldsm_x4_t(smem_u32(sdS + arow * ldp + acol), sa0, sa1, sa2, sa3);
            const int brow = qs * 16 + (mi & 1) * 8 + r;
            const bf16* dob = sdO + brow * LDS + (mi >> 1) * 8;
            const bf16* qb = sQ + brow * LDS + (mi >> 1) * 8;
#pragma unroll
            for (int dpp = 0; dpp < D / 16; ++dpp) {
                uint32_t b0, b1, b2, b3;
                ldsm_x4_t(smem_u32(dob + dpp * 16), b0, b1, b2, b3);
                mma_bf16(ov[dpp * 2], pa0, pa1, pa2, pa3, b0, b1);
                mma_bf16(ov[dpp * 2 + 1], pa0, pa1, pa2, pa3, b2, b3);
                ldsm_x4_t(smem_u32(qb + dpp * 16), b0, b1, b2, b3);
                mma_bf16(ok[dpp * 2], sa0, sa1, sa2, sa3, b0, b1);
                mma_bf16(ok[dpp * 2 + 1], sa0, sa1, sa2, sa3, b2, b3);
            }
        }
        const int key0 = kt * 16 + g, key1 = key0 + 8;
        bf16* dv0 = p.dv + ((long long)b * p.sk + key0) * p.lddv + h * D + 2 * t;
        bf16* dv1 = p.dv + ((long long)b * p.sk + key1) * p.lddv + h * D + 2 * t;
        bf16* dk0 = p.dk + ((long long)b * p.sk + key0) * p.lddk + h * D + 2 * t;
        bf16* dk1 = p.dk + ((long long)b * p.sk + key1) * p.lddk + h * D + 2 * t;
#pragma unroll
        for (int dn = 0; dn < D / 8; ++dn) {
            if (key0 < p.sk) {
                *reinterpret_cast<uint32_t*>(dv0 + dn * 8) = pack_bf16x2(ov[dn][0], ov[dn][1]);
                *reinterpret_cast<uint32_t*>(dk0 + dn * 8) = pack_bf16x2(ok[dn][0], ok[dn][1]);
            }
            if (key1 < p.sk) {
                *reinterpret_cast<uint32_t*>(dv1 + dn * 8) = pack_bf16x2(ov[dn][2], ov[dn][3]);
                *reinterpret_cast<uint32_t*>(dk1 + dn * 8) = pack_bf16x2(ok[dn][2], ok[dn][3]);
            }
        }
        if (p.dbk != nullptr) tile_colsum<D>(ok, key0 < p.sk, key1 < p.sk, s_cs[warp][1], lane);
        if (p.dbv != nullptr) tile_colsum<D>(ov, key0 < p.sk, key1 < p.sk, s_cs[warp][2], lane);
    }
    if (want_cs) {
        __syncthreads();
        for (int i = threadIdx.x; i < 3 * D; i += blockDim.x) {
            float* dst = (i < D) ? p.dbq : (i < 2 * D ? p.dbk : p.dbv);
            if (dst != nullptr) {
                float v = 0.f;
#pragma unroll
                for (int w = 0; w < 8; ++w) v += s_cs[w][i / D][i % D];
                atomicAdd(dst + h * D + (i % D), v);
            }
        }
    }
    }   // items of this CTA
}

static size_t attn_fwd_smem(int sq, int sk, int d) {
    const int sqp = (sq + 15) & ~15, skp = (sk + 15) & ~15;
    return (size_t)(sqp + 2 * skp) * (d + 8) * 2 + skp + 16;
}
static size_t attn_bwd_smem(int sq, int sk, int d, int nsets) {
    const int sqp = (sq + 15) & ~15, skp = (sk + 15) & ~15;
    return (size_t)nsets * (2 * sqp + 2 * skp) * (d + 8) * 2 + (size_t)2 * sqp * (skp + 8) * 2 + (size_t)nsets * skp + 16 +
           (size_t)8 * 3 * d * 4;
}

static int check_attn(const mcan_attn_args* a, const char* who) {
    MCAN_REQUIRE(a->q && a->k && a->v, "%s: null q/k/v", who);
    MCAN_REQUIRE(a->head_dim == 64 || a->head_dim == 128, "%s: head_dim=%d (64 or 128)", who, a->head_dim);
    MCAN_REQUIRE(a->sq >= 1 && a->sq <= kAttnMaxSeq && a->sk >= 1 && a->sk <= kAttnMaxSeq,
                 "%s: sq=%d sk=%d (1..128)", who, a->sq, a->sk);
    MCAN_REQUIRE(a->batch >= 1 && a->heads >= 1, "%s: batch=%d heads=%d", who, a->batch, a->heads);
    MCAN_REQUIRE(a->ldq % 8 == 0 && a->ldk % 8 == 0 && a->ldv % 8 == 0, "%s: ld not multiple of 8", who);
    MCAN_REQUIRE((((uintptr_t)a->q | (uintptr_t)a->k | (uintptr_t)a->v) & 15) == 0, "%s: q/k/v not 16-byte aligned", who);
    MCAN_REQUIRE(a->dropout_p >= 0.f && a->dropout_p < 1.f, "%s: dropout_p=%f", who, a->dropout_p);
    MCAN_REQUIRE((long long)a->batch * a->heads * a->sq * a->sk < (1LL << 32), "%s: too many scores", who);
    return 0;
}

static void fill_attn_params(AttnParams& p, const mcan_attn_args* a) {
    p.q = reinterpret_cast<const bf16*>(a->q);
    p.k = reinterpret_cast<const bf16*>(a->k);
    p.v = reinterpret_cast<const bf16*>(a->v);
    p.q_lo = reinterpret_cast<const bf16*>(a->q_lo);
    p.k_lo = reinterpret_cast<const bf16*>(a->k_lo);
    p.v_lo = reinterpret_cast<const bf16*>(a->v_lo);
    p.out_lo = reinterpret_cast<bf16*>(a->out_lo);
    p.ldq = a->ldq; p.ldk = a->ldk; p.ldv = a->ldv;
    p.mask = a->key_mask;
    p.out = reinterpret_cast<bf16*>(a->out);
    p.ldo = a->ldo;
    p.batch = a->batch; p.heads = a->heads; p.sq = a->sq; p.sk = a->sk;
    p.scale = a->scale;
    p.drop_thr = a->dropout_p > 0.f ? dropout_threshold(a->dropout_p) : 0;
    p.drop_scale = a->dropout_p > 0.f ? 1.f / (1.f - a->dropout_p) : 1.f;
    p.drop_seed = a->dropout_seed;
    p.drop_seed_dev = a->dropout_seed_dev;
}

template <typename K>
static int set_smem_once(K kernel, size_t bytes, size_t* configured) {
    if (bytes > *configured) {
        MCAN_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        *configured = bytes;
    }
    return 0;
}

}  // namespace mcan

using namespace mcan;

template <int D, int NT, bool EXACT>
static int launch_attn_fwd(const AttnParams& p, int grid, int threads, size_t smem, cudaStream_t st) {
    static size_t configured = 48 * 1024;
    if (int rc = set_smem_once(attn_fwd_kernel<D, NT, EXACT>, smem, &configured)) return rc;
    MCAN_CHECK_CUDA(launch_kernel(attn_fwd_kernel<D, NT, EXACT>, dim3(grid), dim3(threads), smem, st, p));
    return 0;
}
template <int D, int NT, bool EXACT>
static int launch_attn_bwd(const AttnParams& p, int grid, int threads, size_t smem, cudaStream_t st) {
    static size_t configured = 48 * 1024;
    if (int rc = set_smem_once(attn_bwd_kernel<D, NT, EXACT>, smem, &configured)) return rc;
    MCAN_CHECK_CUDA(launch_kernel(attn_bwd_kernel<D, NT, EXACT>, dim3(grid), dim3(threads), smem, st, p));
    return 0;
}
// tile-count specialisations: 2 (<= 16 keys: question), 14 (<= 112 keys: 100 image regions), 16 (any)
#define MCAN_ATTN_DISPATCH(FN, D)                                                       \
    do {                                                                                \
        const int nkt = ((sk + 15) & ~15) / 8;                                          \
        if (nkt == 2) return FN<D, 2, true>(p, grid, threads, smem, st);                \
        if (nkt == 14) return FN<D, 14, true>(p, grid, threads, smem, st);              \
        if (nkt == 16) return FN<D, 16, true>(p, grid, threads, smem, st);              \
        if (nkt < 8) return FN<D, 8, false>(p, grid, threads, smem, st);                \
        return FN<D, 16, false>(p, grid, threads, smem, st);                            \
    } while (0)

extern "C" int mcan_attn_fwd(const mcan_attn_args* a) {
    MCAN_REQUIRE(a != nullptr, "mcan_attn_fwd: null args");
    if (int rc = check_attn(a, "mcan_attn_fwd")) return rc;
    MCAN_REQUIRE(a->out && a->ldo % 2 == 0 && ((uintptr_t)a->out & 3) == 0, "mcan_attn_fwd: bad out");
    AttnParams p{};
    fill_attn_params(p, a);
    const size_t smem = attn_fwd_smem(a->sq, a->sk, a->head_dim);
    const int mtiles = (a->sq + 15) / 16;
    const int threads = 32 * (mtiles < 4 ? mtiles : 4);
    const int grid = a->batch * a->heads;
    const int sk = a->sk;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(a->stream);
    if (a->q_lo || a->k_lo || a->v_lo || a->out_lo) {
        MCAN_REQUIRE(a->q_lo && a->k_lo && a->v_lo && a->out_lo, "mcan_attn_fwd: split precision needs all four lo pointers");
        MCAN_REQUIRE(a->dropout_p == 0.f, "mcan_attn_fwd: split precision is an inference mode (no dropout)");
        MCAN_REQUIRE((((uintptr_t)a->q_lo | (uintptr_t)a->k_lo | (uintptr_t)a->v_lo) & 15) == 0 &&
                         ((uintptr_t)a->out_lo & 3) == 0, "mcan_attn_fwd: lo alignment");
        const size_t smem2 = 2 * (smem - 16) + 16;
        static size_t cfg64 = 48 * 1024, cfg128 = 48 * 1024;
        if (a->head_dim == 64) {
            if (int rc = set_smem_once(attn_fwd_split_kernel<64>, smem2, &cfg64)) return rc;
            MCAN_CHECK_CUDA(launch_kernel(attn_fwd_split_kernel<64>, dim3(grid), dim3(threads), smem2, st, p));
        } else {
            if (int rc = set_smem_once(attn_fwd_split_kernel<128>, smem2, &cfg128)) return rc;
            MCAN_CHECK_CUDA(launch_kernel(attn_fwd_split_kernel<128>, dim3(grid), dim3(threads), smem2, st, p));
        }
        MCAN_CHECK_CUDA(cudaGetLastError());
        return 0;
    }
    if (attn_tc_fwd_eligible(p, a->head_dim)) return attn_tc_fwd_launch(p, st);     // image-side queries: tcgen05 path
    if (a->head_dim == 64) MCAN_ATTN_DISPATCH(launch_attn_fwd, 64);
    MCAN_ATTN_DISPATCH(launch_attn_fwd, 128);
}

extern "C" int mcan_attn_bwd(const mcan_attn_bwd_args* a) {
    MCAN_REQUIRE(a != nullptr, "mcan_attn_bwd: null args");
    if (int rc = check_attn(&a->fwd, "mcan_attn_bwd")) return rc;
    MCAN_REQUIRE(a->dout && a->dq && a->dk && a->dv, "mcan_attn_bwd: null gradient pointer");
    MCAN_REQUIRE(a->lddo % 8 == 0 && ((uintptr_t)a->dout & 15) == 0, "mcan_attn_bwd: dout alignment");
    MCAN_REQUIRE(a->lddq % 2 == 0 && a->lddk % 2 == 0 && a->lddv % 2 == 0 &&
                     (((uintptr_t)a->dq | (uintptr_t)a->dk | (uintptr_t)a->dv) & 3) == 0,
                 "mcan_attn_bwd: dq/dk/dv alignment");
    AttnParams p{};
    fill_attn_params(p, &a->fwd);
    p.dout = reinterpret_cast<const bf16*>(a->dout);
    p.lddo = a->lddo;
    p.dq = reinterpret_cast<bf16*>(a->dq);
    p.dk = reinterpret_cast<bf16*>(a->dk);
    p.dv = reinterpret_cast<bf16*>(a->dv);
    p.lddq = a->lddq; p.lddk = a->lddk; p.lddv = a->lddv;
    p.dbq = a->dbq; p.dbk = a->dbk; p.dbv = a->dbv;
    if (attn_tc_bwd_eligible(p, a->fwd.head_dim)) return attn_tc_bwd_launch(p, reinterpret_cast<cudaStream_t>(a->fwd.stream));
    // Large tiles (one CTA per SM anyway: the image self-attention, 124 KB): persistent grid, the operand tiles of the
    // CTA's next (batch, head) are prefetched into a second shared-memory set while the current one is processed --
    // the load phase (cp.async of Q, K, V, dO from HBM) no longer sits exposed in front of every CTA's compute.
    // MCAN_ATTN_PREFETCH=0: one (batch, head) per CTA as before.
    static const bool allow_prefetch = [] { const char* e = getenv("MCAN_ATTN_PREFETCH"); return !(e && e[0] == '0'); }();
    const size_t single = attn_bwd_smem(a->fwd.sq, a->fwd.sk, a->fwd.head_dim, 1);
    const size_t dbl = attn_bwd_smem(a->fwd.sq, a->fwd.sk, a->fwd.head_dim, 2);
    const int items = a->fwd.batch * a->fwd.heads;
    const int sms = device_num_sms();
    p.prefetch = (allow_prefetch && single > 64 * 1024 && dbl <= 220 * 1024 && sms > 0 && items > sms) ? 1 : 0;
    const size_t smem = p.prefetch ? dbl : single;
    const int mtiles = (a->fwd.sq + 15) / 16, ktiles = (a->fwd.sk + 15) / 16;
    const int mx = mtiles > ktiles ? mtiles : ktiles;
    const int threads = 32 * (mx < 8 ? mx : 8);
    const int grid = p.prefetch ? sms : items;
    const int sk = a->fwd.sk;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(a->fwd.stream);
    if (a->fwd.head_dim == 64) MCAN_ATTN_DISPATCH(launch_attn_bwd, 64);
    MCAN_ATTN_DISPATCH(launch_attn_bwd, 128);
}

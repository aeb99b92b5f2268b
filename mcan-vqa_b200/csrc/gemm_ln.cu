// G2: GEMM with the residual add AND MCAN's LayerNorm fused into the epilogue.
//
//   s = resid + dropout(A W^T + bias)                      (mca.py:119-125,152-162: x + dropoutN(sublayer(x)))
//   y = a_2 * (s - mean(s)) / (std_unbiased(s) + eps) + b_2   (net_utils.py:56-60)
//
// replaces  GEMM[... + resid] -> s (fp32) to HBM -> ln_fwd reads s back  for the sub-layer outputs of SA / SGA
// (merge projection and FFN2, forward).  LayerNorm needs whole rows, so one CLUSTER owns 256 complete rows:
//
//   H = 512 :  cluster = 1 CTA pair ; the pair's accumulator is 256 x 512  (two cta_group::2 MMAs of N = 256)
//   H = 1024:  cluster = 2 CTA pairs; pair p owns columns [512p, 512p + 512); the A tile (the same 256 rows for
//              both pairs) is fetched once and TMA-MULTICAST across the pairs; the row statistics (count, mean,
//              M2) are exchanged through distributed shared memory and merged with Chan's formula.
//
// Every CTA uses all 512 TMEM columns for its 128 x 512 fp32 accumulator, so there is no second accumulator
// stage and the kernel is not persistent: grid = ceil(M / 256) clusters, one tile each.  Per k-block a CTA
// stages 16 KB of A (8 KB fetched + 8 KB received by multicast when H = 1024) and 32 KB of B for 8 MMAs of
// 128 cycles: 40-48 KB of L2 -> smem traffic per 1024 tensor cycles, against 32 KB per 512 cycles for the
// 256 x 256 tile of the persistent GEMM -- the fused tile is the cheaper one to feed.
//
// MEASURED (B200, tools/gemm_ln_bench.py, L2 flushed): 6400 x 1024 x 1024: fused 49 us vs GEMM 34 us + LayerNorm
// 12 us; 6400 x 1024 x 4096: 79 vs 67 us.  The main loop runs at the tensor-pipe rate (0.6 us per k-block), but
// with 25 clusters on 148 SMs there is ONE wave, so the two-pass epilogue (65 k elements per CTA on 8 warps:
// bias, dropout hash, residual, exact statistics, normalise -- about 20 us, instruction-latency bound; removing
// outputs or deepening the prefetch does not move it) is fully exposed, while the persistent GEMM hides the
// same per-element work behind the next tile's main loop and the stand-alone LayerNorm kernel runs at 64 warps per
// SM.  At batch 64 the fusion therefore does not pay and blocks.py keeps it opt-in (MCAN_FUSE_LN=1); it needs
// >= 3 row tiles per cluster slot (batch >= 512 per GPU) to hide its epilogue.
//
// Epilogue (8 warps, two per TMEM lane quarter, one per 256-column half):
//   pass 1: tcgen05.ld fragment -> + bias -> dropout -> + resid -> s ; s to global (fp32, saved for backward) and
//           back to TMEM (tcgen05.st); per-row (mean, M2) of the fragment exactly (two passes over registers),
//           merged over lanes (shuffles), chunks (registers), warps (shared memory), pairs (DSMEM);
//   pass 2: tcgen05.ld s -> y -> y_f32 / y_bf16 to global; (mean, sigma) per row for the backward kernel.
#include <string.h>

#include "../../include/mcan_b200.h"
#include "common.cuh"

namespace mcan {

int make_tmap_bf16(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_rows);
int device_num_sms_raw();

namespace gln {

constexpr int BLOCK_M = 128;      // rows per CTA
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int kCols = 512;        // accumulator columns per CTA (= all of TMEM)
constexpr int kStages = 4;
constexpr uint32_t kABytes = BLOCK_M * BLOCK_K * 2;          // 16 KB
constexpr uint32_t kBBytes = 2 * 128 * BLOCK_K * 2;          // two 128-row sub-tiles (one per N = 256 MMA): 32 KB
constexpr uint32_t kStageBytes = kABytes + kBBytes;          // 48 KB
constexpr int kEpilogueWarps = 8;
constexpr int kProducerWarp = 8;
constexpr int kMmaWarp = 9;
constexpr int kThreads = 32 * 10;
// after the stages: barriers, tmem slot, row statistics
constexpr uint32_t kBarBytes = 128;
constexpr uint32_t kStatBytes = 4 * BLOCK_M * 8;             // local halves [2][128] float2, merged [128] float2, remote [128] float2
constexpr uint32_t kSmemBytes = kStages * kStageBytes + kBarBytes + kStatBytes;

struct Params {
    CUtensorMap tma_a, tma_b;
    int m, n, k, kblocks;
    const float* bias;
    uint32_t drop_thr;
    float drop_scale;
    uint32_t drop_seed;
    const uint32_t* drop_seed_dev;
    const float* resid;
    long long ldr;
    const float* a2;
    const float* b2;
    float eps;
    float* s_f32;
    float* y_f32;
    bf16* y_bf16;
    float* mean;
    float* sigma;
};

__device__ __forceinline__ void tmem_st_16x256b_x8(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.16x256b.x8.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Chan et al.: statistics (n, mean, M2) of the union of two disjoint sets
__device__ __forceinline__ void chan_merge(float& n, float& mean, float& m2, float nb, float meanb, float m2b) {
    const float nt = n + nb;
    const float d = meanb - mean;
    const float f = nb / nt;
    mean += d * f;
    m2 += m2b + d * d * n * f;
    n = nt;
}

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void st_shared_remote_f32x2(float2* local_addr, uint32_t cta, float2 v) {
    asm volatile(
        "{\n"
        ".reg .b32 ra;\n"
        "mapa.shared::cluster.u32 ra, %0, %1;\n"
        "st.shared::cluster.v2.f32 [ra], {%2, %3};\n"
        "}\n" ::"r"(smem_u32(local_addr)),
        "r"(cta), "f"(v.x), "f"(v.y)
        : "memory");
}

// NC = CTA pairs per cluster (1: H = 512, 2: H = 1024)
template <int NC>
__global__ void __launch_bounds__(kThreads, 1)
gemm_ln_kernel(const __grid_constant__ Params p) {
    constexpr int CL = 2 * NC;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    pdl_launch_dependents();
    if ((smem_u32(smem) & 1023U) != 0) {
        if (threadIdx.x == 0) printf("mcan gemm_ln: dynamic smem base not 1024-byte aligned\n");
        __trap();
    }
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full_bar = empty_bar + kStages;
    uint64_t* stat_bar = tmem_full_bar + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stat_bar + 1);
    float2* stat_half = reinterpret_cast<float2*>(smem + kStages * kStageBytes + kBarBytes);   // [2][128] (mean, M2), n = 256
    float2* stat_own = stat_half + 2 * BLOCK_M;                                                 // [128] this CTA's 512 columns
    float2* stat_peer = stat_own + BLOCK_M;                                                     // [128] written by the other pair

    const int warp = __shfl_sync(0xffffffffU, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();
    const uint32_t rank = crank & 1U;          // M half inside the pair
    const uint32_t pair = crank >> 1;          // N half of the row (NC == 2)
    const bool leader = rank == 0;
    const int m_tile = (int)blockIdx.x / CL;
    const int m0 = m_tile * (2 * BLOCK_M) + (int)rank * BLOCK_M;      // first row of this CTA
    const int n0 = (int)pair * kCols;                                 // first column of this CTA

    if (warp == kProducerWarp && lane == 0) {
        prefetch_tmap(&p.tma_a);
        prefetch_tmap(&p.tma_b);
    }
    if (warp == kMmaWarp) {
        if (lane == 0) {
            for (int s = 0; s < kStages; ++s) {
                mbar_init(&full_bar[s], 1);
                mbar_init(&empty_bar[s], NC);      // one commit per pair whose MMAs read (multicast) data of this slot
            }
            mbar_init(tmem_full_bar, 1);
            mbar_init(stat_bar, 1);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc_cg2(tmem_slot, kCols);
        tmem_relinquish_cg2();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp == kProducerWarp) {
        // ===================== TMA producer =====================
        int stage = 0;
        uint32_t phase = 0;
        const uint16_t mc_mask = (uint16_t)((1U << rank) | (1U << (2 + rank)));
        for (int kb = 0; kb < p.kblocks; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (elect_one()) {
                uint8_t* sa = smem + stage * kStageBytes;
                uint8_t* sb = sa + kABytes;
                if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * kStageBytes);
                if (NC == 2) {
                    // this CTA fetches rows [64 * pair, +64) of the 128-row A tile and multicasts them to its
                    // counterpart in the other pair (same offset in both)
                    tma_load_2d_cg2_mc(sa + pair * (64 * 128), &p.tma_a, &full_bar[stage], kb * BLOCK_K,
                                       m0 + (int)pair * 64, mc_mask);
                } else {
                    tma_load_2d_cg2(sa, &p.tma_a, &full_bar[stage], kb * BLOCK_K, m0);
                }
                // B: for the j-th N = 256 MMA this CTA holds rows [n0 + 256 j + 128 rank, +128) of W
#pragma unroll
                for (int j = 0; j < 2; ++j)
                    tma_load_2d_cg2(sb + j * (128 * 128), &p.tma_b, &full_bar[stage], kb * BLOCK_K,
                                    n0 + 256 * j + 128 * (int)rank);
            }
            __syncwarp();
            if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
    } else if (warp == kMmaWarp) {
        // ===================== MMA issuer (leader CTA of each pair) =====================
        if (leader) {
            constexpr uint32_t idesc = make_idesc_bf16(2 * BLOCK_M, 256, 0, 0);
            constexpr uint32_t desc_hi = (uint32_t)(make_smem_desc_sw128_const(0, 0, 1024) >> 32);
            const uint32_t smem0 = smem_u32(smem);
            const uint32_t a_lo0 = (smem0 & 0x3FFFFU) >> 4;
            const uint32_t b_lo0 = ((smem0 + kABytes) & 0x3FFFFU) >> 4;
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < p.kblocks; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a_lo = a_lo0 + (uint32_t)stage * (kStageBytes >> 4);
                    const uint32_t b_lo = b_lo0 + (uint32_t)stage * (kStageBytes >> 4);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                            const uint64_t ad = ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo + ((k * UMMA_K * 2) >> 4));
                            const uint64_t bd = ((uint64_t)desc_hi << 32) |
                                                (uint64_t)(b_lo + ((j * 128 * 128 + k * UMMA_K * 2) >> 4));
                            umma_bf16_cg2(tmem_base + (uint32_t)(j * 256), ad, bd, idesc, (kb > 0 || k > 0) ? 1U : 0U);
                        }
                    }
                    // the slot is free (in every CTA that holds multicast data of it) when these MMAs retire
                    umma_commit_cg2(&empty_bar[stage], (uint16_t)((1U << CL) - 1U));
                    if (kb == p.kblocks - 1) umma_commit_cg2(tmem_full_bar, (uint16_t)(3U << (2 * pair)));
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue =====================
        const int quad = warp & 3;       // TMEM lanes / tile rows [32 quad, +32)
        const int half = warp >> 2;      // accumulator columns [256 half, +256)
        const int g = lane >> 2, t = lane & 3;
        const uint32_t drop_seed =
            p.drop_seed ^ ((p.drop_thr != 0 && p.drop_seed_dev != nullptr) ? __ldg(p.drop_seed_dev) : 0U);
        // running statistics of the four rows this lane touches: [hb][h]
        float rn[2][2], rmean[2][2], rm2[2][2];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) { rn[a][b] = 0.f; rmean[a][b] = 0.f; rm2[a][b] = 0.f; }

        // Fragments are visited in the order f = 2c + hb (c = 64-column chunk, hb = 16-row block).  All global
        // accesses use the T8 layout (8 consecutive columns per thread, 16-byte vectors); the residual of
        // fragment f + 1 is fetched while fragment f is processed.
        auto frag_rows = [&](int hb, long long (&rows)[2]) {
            rows[0] = (long long)m0 + quad * 32 + hb * 16 + g;
            rows[1] = rows[0] + 8;
        };
        auto load_resid = [&](int f, float4 (&res)[8]) {
            const int c = f >> 1, hb = f & 1;
            long long rows[2];
            frag_rows(hb, rows);
            const int col = n0 + 256 * half + 64 * c + 8 * t;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    if (rows[h] < p.m) {
                        const float4* rp = reinterpret_cast<const float4*>(p.resid + rows[h] * p.ldr + col + 32 * q);
                        res[2 * (2 * h + q)] = rp[0];
                        res[2 * (2 * h + q) + 1] = rp[1];
                    } else {
                        res[2 * (2 * h + q)] = make_float4(0.f, 0.f, 0.f, 0.f);
                        res[2 * (2 * h + q) + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
            }
        };
        float4 res_next[8];
        load_resid(0, res_next);            // before waiting for the MMAs
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        // ---- pass 1: s = resid + dropout(acc + bias) -> global + TMEM (T8 order), exact row statistics ----
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
            const int colc = 256 * half + 64 * c;            // column inside the CTA's accumulator
            const int col = n0 + colc + 8 * t;               // global column of this lane's group q = 0 (+ 32q + e)
            float4 bia[4];                                   // bias of this lane's 16 columns
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                bia[2 * q] = p.bias != nullptr ? __ldg(reinterpret_cast<const float4*>(p.bias + col + 32 * q)) : make_float4(0.f, 0.f, 0.f, 0.f);
                bia[2 * q + 1] = p.bias != nullptr ? __ldg(reinterpret_cast<const float4*>(p.bias + col + 32 * q) + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int hb = 0; hb < 2; ++hb) {
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32 + hb * 16) << 16) + (uint32_t)colc;
                uint32_t r[32];
                tmem_ld_16x256b_x8(taddr, r);
                float4 res[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) res[i] = res_next[i];
                if (2 * c + hb + 1 < 8) load_resid(2 * c + hb + 1, res_next);
                long long rows[2];
                frag_rows(hb, rows);
                tmem_ld_wait();
                float v[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
                quad_transpose(v, t);
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const float4 b0 = bia[2 * q], b1 = bia[2 * q + 1];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        T8(v, q, h, 0) += b0.x; T8(v, q, h, 1) += b0.y; T8(v, q, h, 2) += b0.z; T8(v, q, h, 3) += b0.w;
                        T8(v, q, h, 4) += b1.x; T8(v, q, h, 5) += b1.y; T8(v, q, h, 6) += b1.z; T8(v, q, h, 7) += b1.w;
                    }
                }
                if (p.drop_thr != 0) {
                    const uint32_t thr = p.drop_thr;
                    const float sc = p.drop_scale;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const uint32_t base = (uint32_t)(rows[h] * (long long)p.n + col + 32 * q) >> 1;
#pragma unroll
                            for (int e = 0; e < 8; e += 2) {
                                const uint32_t rnd = dropout_bits_pair(base + (e >> 1), drop_seed);
                                T8(v, q, h, e) = ((rnd & 0xFFFFU) >= thr) ? T8(v, q, h, e) * sc : 0.f;
                                T8(v, q, h, e + 1) = ((rnd >> 16) >= thr) ? T8(v, q, h, e + 1) * sc : 0.f;
                            }
                        }
                    }
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const float4 r0 = res[2 * (2 * h + q)], r1 = res[2 * (2 * h + q) + 1];
                        T8(v, q, h, 0) += r0.x; T8(v, q, h, 1) += r0.y; T8(v, q, h, 2) += r0.z; T8(v, q, h, 3) += r0.w;
                        T8(v, q, h, 4) += r1.x; T8(v, q, h, 5) += r1.y; T8(v, q, h, 6) += r1.z; T8(v, q, h, 7) += r1.w;
                    }
                }
                // s back to TMEM for pass 2 (the registers are stored in T8 order, so pass 2 reads T8 directly),
                // and to global memory for the backward pass
#pragma unroll
                for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(v[i]);
                tmem_st_16x256b_x8(taddr, r);
                if (p.s_f32 != nullptr) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        if (rows[h] < p.m) {
#pragma unroll
                            for (int q = 0; q < 2; ++q) {
                                float4* o = reinterpret_cast<float4*>(p.s_f32 + rows[h] * (long long)p.n + col + 32 * q);
                                o[0] = make_float4(T8(v, q, h, 0), T8(v, q, h, 1), T8(v, q, h, 2), T8(v, q, h, 3));
                                o[1] = make_float4(T8(v, q, h, 4), T8(v, q, h, 5), T8(v, q, h, 6), T8(v, q, h, 7));
                            }
                        }
                    }
                }
                // exact (mean, M2) of the 64 columns of this fragment, per row
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float sum = 0.f;
#pragma unroll
                    for (int q = 0; q < 2; ++q)
#pragma unroll
                        for (int e = 0; e < 8; ++e) sum += T8(v, q, h, e);
                    sum += __shfl_xor_sync(0xffffffffU, sum, 1);
                    sum += __shfl_xor_sync(0xffffffffU, sum, 2);
                    const float mean = sum * (1.f / 64.f);
                    float sq = 0.f;
#pragma unroll
                    for (int q = 0; q < 2; ++q)
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float d = T8(v, q, h, e) - mean;
                            sq += d * d;
                        }
                    sq += __shfl_xor_sync(0xffffffffU, sq, 1);
                    sq += __shfl_xor_sync(0xffffffffU, sq, 2);
                    chan_merge(rn[hb][h], rmean[hb][h], rm2[hb][h], 64.f, mean, sq);
                }
            }
        }
        tmem_st_wait();
        // ---- row statistics: warps of the other column half (smem), then the other pair (DSMEM) ----
        if (t == 0) {
#pragma unroll
            for (int hb = 0; hb < 2; ++hb)
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    stat_half[half * BLOCK_M + quad * 32 + hb * 16 + g + 8 * h] = make_float2(rmean[hb][h], rm2[hb][h]);
        }
        named_bar_sync(1, kEpilogueWarps * 32);
        const int er = threadIdx.x;      // epilogue thread index 0..255; threads 0..127 own one row each
        if (er < BLOCK_M) {
            const float2 a = stat_half[er], b = stat_half[BLOCK_M + er];
            float n = 256.f, mean = a.x, m2 = a.y;
            chan_merge(n, mean, m2, 256.f, b.x, b.y);
            stat_own[er] = make_float2(mean, m2);
            if (NC == 2) {
                st_shared_remote_f32x2(&stat_peer[er], crank ^ 2U, make_float2(mean, m2));
                asm volatile("fence.acq_rel.cluster;" ::: "memory");
            }
        }
        named_bar_sync(1, kEpilogueWarps * 32);          // stat_own complete; all remote stores issued ...
        if (NC == 2) {
            if (threadIdx.x == 0) mbar_arrive_release_cluster(stat_bar, crank ^ 2U);   // ... and released to the peer
            mbar_wait_acq_cluster(stat_bar, 0);
        }
        // final mean / 1/(sigma + eps) of the four rows of this lane
        float fmean[2][2], frstd[2][2];
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int row = quad * 32 + hb * 16 + g + 8 * h;
                const float2 a = stat_own[row];
                float n = (float)kCols, mean = a.x, m2 = a.y;
                if (NC == 2) {
                    const float2 b = stat_peer[row];
                    // merge in a fixed order (pair 0 first) so that both pairs compute bit-identical statistics
                    if (pair == 0) chan_merge(n, mean, m2, (float)kCols, b.x, b.y);
                    else { n = (float)kCols; mean = b.x; m2 = b.y; chan_merge(n, mean, m2, (float)kCols, a.x, a.y); }
                }
                const float sigma = sqrtf(m2 / (n - 1.f));
                fmean[hb][h] = mean;
                frstd[hb][h] = 1.f / (sigma + p.eps);
                const long long grow = (long long)m0 + row;
                if (half == 0 && t == 0 && pair == 0 && grow < p.m) {
                    if (p.mean != nullptr) p.mean[grow] = mean;
                    if (p.sigma != nullptr) p.sigma[grow] = sigma;
                }
            }
        }
        // ---- pass 2: y = a_2 (s - mean) / (sigma + eps) + b_2 ----
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
            const int colc = 256 * half + 64 * c;
            const int col = n0 + colc + 8 * t;
            float4 ga[4], be[4];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                ga[2 * q] = __ldg(reinterpret_cast<const float4*>(p.a2 + col + 32 * q));
                ga[2 * q + 1] = __ldg(reinterpret_cast<const float4*>(p.a2 + col + 32 * q) + 1);
                be[2 * q] = __ldg(reinterpret_cast<const float4*>(p.b2 + col + 32 * q));
                be[2 * q + 1] = __ldg(reinterpret_cast<const float4*>(p.b2 + col + 32 * q) + 1);
            }
#pragma unroll
            for (int hb = 0; hb < 2; ++hb) {
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32 + hb * 16) << 16) + (uint32_t)colc;
                uint32_t r[32];
                tmem_ld_16x256b_x8(taddr, r);
                long long rows[2];
                frag_rows(hb, rows);
                tmem_ld_wait();
                float v[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (rows[h] < p.m) {
                        const float mu = fmean[hb][h], rs = frstd[hb][h];
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const float ya[8] = {ga[2 * q].x, ga[2 * q].y, ga[2 * q].z, ga[2 * q].w,
                                                 ga[2 * q + 1].x, ga[2 * q + 1].y, ga[2 * q + 1].z, ga[2 * q + 1].w};
                            const float yb[8] = {be[2 * q].x, be[2 * q].y, be[2 * q].z, be[2 * q].w,
                                                 be[2 * q + 1].x, be[2 * q + 1].y, be[2 * q + 1].z, be[2 * q + 1].w};
                            float y[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) y[e] = ya[e] * (T8(v, q, h, e) - mu) * rs + yb[e];
                            const long long off = rows[h] * (long long)p.n + col + 32 * q;
                            if (p.y_f32 != nullptr) {
                                float4* o = reinterpret_cast<float4*>(p.y_f32 + off);
                                o[0] = make_float4(y[0], y[1], y[2], y[3]);
                                o[1] = make_float4(y[4], y[5], y[6], y[7]);
                            }
                            if (p.y_bf16 != nullptr)
                                *reinterpret_cast<uint4*>(p.y_bf16 + off) = make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]),
                                                                                        pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
                        }
                    }
                }
            }
        }
    }

    tc_fence_before();
    cluster_sync_all();       // peers' shared memory / barriers stay alive until everyone is done
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem_dealloc_cg2(tmem_base, kCols);
    }
}

template <int NC>
static int launch(const Params& p, cudaStream_t stream) {
    static bool configured[64] = {false};
    int dev = 0;
    MCAN_CHECK_CUDA(cudaGetDevice(&dev));
    MCAN_REQUIRE(dev >= 0 && dev < 64, "device index %d", dev);
    auto kernel = gemm_ln_kernel<NC>;
    if (!configured[dev]) {
        MCAN_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
        configured[dev] = true;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    const int m_tiles = (p.m + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
    cfg.gridDim = dim3((unsigned)(m_tiles * 2 * NC), 1, 1);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2 * NC;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    MCAN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, p));
    return 0;
}

}  // namespace gln
}  // namespace mcan

using namespace mcan;

extern "C" int mcan_gemm_ln(const mcan_gemm_ln_args* a) {
    MCAN_REQUIRE(a != nullptr, "mcan_gemm_ln: null args");
    MCAN_REQUIRE(a->m > 0 && a->k > 0 && (a->n == 512 || a->n == 1024),
                 "mcan_gemm_ln: m=%lld k=%lld n=%lld (n must be 512 or 1024: one cluster owns whole rows)",
                 (long long)a->m, (long long)a->k, (long long)a->n);
    MCAN_REQUIRE(a->m < (1LL << 31) && a->k < (1LL << 31) && a->m * a->n < (1LL << 32), "mcan_gemm_ln: shape too large");
    MCAN_REQUIRE(a->a && a->b && a->resid && a->ln_a2 && a->ln_b2, "mcan_gemm_ln: null operand");
    MCAN_REQUIRE(a->y_f32 || a->y_bf16, "mcan_gemm_ln: no output");
    MCAN_REQUIRE(a->dropout_p >= 0.f && a->dropout_p < 1.f, "mcan_gemm_ln: dropout_p=%f", a->dropout_p);
    MCAN_REQUIRE(a->ldr % 4 == 0 && ((uintptr_t)a->resid & 15) == 0, "mcan_gemm_ln: resid alignment");
    MCAN_REQUIRE((((uintptr_t)a->s_f32 | (uintptr_t)a->y_f32 | (uintptr_t)a->bias | (uintptr_t)a->ln_a2 | (uintptr_t)a->ln_b2 |
                   (uintptr_t)a->y_bf16) & 15) == 0,
                 "mcan_gemm_ln: every vector / matrix must be 16-byte aligned");
    gln::Params p;
    memset(&p, 0, sizeof(p));
    p.m = (int)a->m;
    p.n = (int)a->n;
    p.k = (int)a->k;
    p.kblocks = (int)((a->k + gln::BLOCK_K - 1) / gln::BLOCK_K);
    const int nc = a->n == 1024 ? 2 : 1;
    if (int rc = make_tmap_bf16(&p.tma_a, a->a, (uint64_t)a->k, (uint64_t)a->m, (uint64_t)a->lda, nc == 2 ? 64 : 128)) return rc;
    if (int rc = make_tmap_bf16(&p.tma_b, a->b, (uint64_t)a->k, (uint64_t)a->n, (uint64_t)a->ldb, 128)) return rc;
    p.bias = a->bias;
    p.drop_thr = a->dropout_p > 0.f ? dropout_threshold(a->dropout_p) : 0;
    p.drop_scale = a->dropout_p > 0.f ? 1.0f / (1.0f - a->dropout_p) : 1.0f;
    p.drop_seed = a->dropout_seed;
    p.drop_seed_dev = a->dropout_seed_dev;
    p.resid = a->resid;
    p.ldr = a->ldr;
    p.a2 = a->ln_a2;
    p.b2 = a->ln_b2;
    p.eps = a->eps;
    p.s_f32 = a->s_f32;
    p.y_f32 = a->y_f32;
    p.y_bf16 = reinterpret_cast<bf16*>(a->y_bf16);
    p.mean = a->mean;
    p.sigma = a->sigma;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(a->stream);
    return nc == 2 ? gln::launch<2>(p, st) : gln::launch<1>(p, st);
}

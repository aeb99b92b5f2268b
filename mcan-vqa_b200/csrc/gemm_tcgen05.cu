// G1/G2/G3: persistent, warp-specialised tcgen05 GEMM with a fused epilogue.
//
//   D[M,N] = epilogue( sum_s A_s[M,K] * B_s[N,K]^T ),  bf16 operands, fp32 accumulate in TMEM.
//
// It is the one GEMM behind every nn.Linear of the MCAN hot path -- forward
// (mca.py:33,40,47,61; net_utils.py:26,45; net.py:39,53), dgrad (B read MN-major straight
// from the (out,in) weight) and wgrad (both operands read MN-major straight from the
// activations, split-K with fp32 atomics).  No transposed copies are ever made.
//
// CTA = 352 threads:  warp 0 = TMA producer, warp 1 = TMEM owner + single-thread MMA issuer,
// warps 2..9 = epilogue (two warps per TMEM lane quarter, alternating 64-column chunks), warp 10 =
// tile scheduler.  Pipelines: smem ring full/empty (TMA <-> MMA) and a double-buffered TMEM
// accumulator full/empty (MMA <-> epilogue), so the epilogue of tile i overlaps the main loop of
// tile i+1.  Grid = min(#work units, #SMs); work unit = (output tile, K split).
// Two schedules (mcan_set_gemm_schedule):
//   static  -- cluster c takes units c, c + #clusters, ...: nothing on the critical path; used
//              when the GEMM owns the GPU;
//   dynamic -- the scheduler warp of the leader CTA claims units from a global counter and
//              broadcasts them to every role (of both CTAs of a pair) through a 2-slot smem ring,
//              one unit ahead of the producer: a CTA whose SM is held by a co-running kernel
//              (the NCCL all-reduce overlapping the backward pass) simply claims fewer units.
//              Under the static schedule every GEMM ran 1.5x slower as soon as another kernel
//              pinned 4 SMs (tools/contention_bench.py).
//
// Tile 128 x BLOCK_N (128 | 256) x 64.  Operand tiles are TMA boxes with the 128-byte swizzle:
//   K-major  tile [rows x 64 k]  : one box {64, rows}; UMMA desc SBO = 1024 B, k-step = +32 B
//   MN-major tile [64 k x rows]  : rows/64 boxes {64 mn, 64 k} of 8 KiB; UMMA desc
//                                  LBO = 8 KiB (next 64 mn), SBO = 1024 B (next 8 k), k-step = +2 KiB
#include <atomic>
#include <mutex>
#include <stdlib.h>
#include <string.h>
#include <type_traits>
#include <unordered_map>

#include "../../include/mcan_b200.h"
#include "common.cuh"

namespace mcan {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int kSchedStages = 2;      // depth of the work-unit broadcast ring (dynamic tile scheduler)
constexpr int kEpilogueWarps = 8;   // warps 0..7: two warps per TMEM lane quarter, alternating 64-column chunks
// The single-thread roles get the HIGHEST warp ids: the per-SMSP issue arbiter prefers the highest
// warp id, and the MMA issue loop is the critical path of the whole kernel (round-1 ncu source view:
// the MMA thread was never blocked on a barrier -- its ~105-instruction issue loop took ~800 cycles
// per k-block against a 512-cycle tensor-pipe floor, sharing its scheduler with two epilogue warps).
constexpr int kSchedWarp = kEpilogueWarps;        // tile scheduler (idle under the static schedule)
constexpr int kProducerWarp = kEpilogueWarps + 1;
constexpr int kMmaWarp = kEpilogueWarps + 2;
constexpr int kGemmThreads = 32 * (kMmaWarp + 1);

// Grouped launch (mcan_gemm_grouped): up to kMaxGroups independent problems D_g[M_g,N_g] += A_g B_g^T that share
// K and the operand layouts run as ONE persistent grid -- the weight-gradient GEMMs of one layer, which feed
// nothing in the backward chain.  Every launch costs ~8-10 us of fixed overhead (launch, prologue, pipeline
// fill, last epilogue) and each small problem alone fills only part of the machine; together their tiles
// form full waves.  Work unit -> (group, tile of the group, K split); operand maps and output per group.
constexpr int kMaxGroups = 8;
constexpr int kMaxTmaps = kMaxGroups > MCAN_MAX_GEMM_SEGMENTS ? kMaxGroups : MCAN_MAX_GEMM_SEGMENTS;
struct GroupDesc {
    int m, n, n_tiles, tile_start;     // tile_start: index of the group's first tile in the launch
    void* out;                         // fp32, or bf16 when GemmParams::group_bf16 is set
    long long ldo;
};

struct alignas(64) GemmParams {
    CUtensorMap tma_a[kMaxTmaps];      // per segment, or per group in a grouped launch
    CUtensorMap tma_b[kMaxTmaps];
    CUtensorMap tma_b_half[MCAN_MAX_GEMM_SEGMENTS];   // K-major B, box of half as many rows (tail splitting)
    int num_groups;                    // 0: one problem
    int group_bf16;                    // grouped launch writes bf16 outputs (EPI = 1 instantiation)
    GroupDesc grp[kMaxGroups];
    int num_seg;
    int m, n, k;
    int m_tiles, n_tiles, splits, kblocks;
    // Tail splitting (splits == 1, BLOCK_N == 256): output tiles [0, full_tiles) are work units of
    // the full BLOCK_N width; each remaining tile is TWO units of width BLOCK_N / 2.  The host picks
    // full_tiles = whole waves, so the last, partial wave runs as half-width tiles on twice as many
    // CTA pairs (100 tiles on 74 pairs: 2 tile-times -> ~1.6).  full_tiles == tiles: no splitting.
    int full_tiles;
    int units;
    // epilogue
    const float* bias;
    int relu;
    uint32_t drop_thr;
    float drop_scale;
    uint32_t drop_seed;
    const uint32_t* drop_seed_dev;
    const bf16* gate;
    long long ldg;
    float gate_scale;
    const float* resid;
    long long ldr;
    float* out_f32;
    long long ldo_f32;
    bf16* out_bf16;
    bf16* out_lo;
    long long ldo_bf16;
    int accumulate;
    float* colsum;   // += column sums of the epilogue output (fp32 [N]); see mcan_gemm_args::colsum
    int debug;   // profiling experiments only (MCAN_GEMM_DEBUG): bit0 skip all global stores, bit1 skip the TMA
                 // loads (MMAs run on stale shared memory), bit2 skip the MMAs (loads and commits only)
    int* tile_counter;   // dynamic tile scheduler: next unclaimed work unit (0 at launch, reset by the last claim)
};

// CG = CTAs per MMA (cta_group): 1 = one SM per 128 x BLOCK_N tile, 2 = a CTA pair computes a
// 256 x BLOCK_N tile with one tcgen05.mma.cta_group::2 -- each CTA stages its own 128 rows of A
// and only HALF of the B tile, which halves the L2->smem operand traffic and the per-SM shared
// memory read bandwidth of the UMMA (the two limits of the 1-CTA kernel).
template <int BLOCK_N, int CG>
struct GemmCfg {
    static constexpr uint32_t kABytes = BLOCK_M * BLOCK_K * 2;
    static constexpr uint32_t kBBytes = (BLOCK_N / CG) * BLOCK_K * 2;
    static constexpr uint32_t kStageBytes = kABytes + kBBytes;
    static constexpr int kStages = (192 * 1024) / kStageBytes;
    static constexpr uint32_t kTmemCols = 2 * BLOCK_N;
    static constexpr uint32_t kBarrierBytes = (2 * kStages + 4 + 2 * kSchedStages) * 8 + 16 + 4 * kSchedStages;
    static constexpr uint32_t kSmemBytes = kStages * kStageBytes + kBarrierBytes;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b),
                 "f"(c), "f"(d)
                 : "memory");
}

__device__ __forceinline__ void red_add_v2(float* addr, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

// Fused epilogue for a [16 rows x 64 columns] accumulator fragment.  tcgen05.ld 16x256b hands thread
// (g = lane/4, t = lane%4), for k < 8 and h < 2, the two adjacent columns col0 + 8k + 2t, +1 of row
// row0 + g + 8h (register r[4k + 2h + c]).  Storing straight from that layout means 4- / 8-byte
// accesses whose warp-wide footprint is 8 rows x 16 / 32 bytes: every instruction costs 8 L1TEX
// wavefronts for 128 / 256 bytes.  ncu on round 1's kernel: the epilogue's LSU wavefronts took 40 % of
// the L1TEX data pipe -- the same pipe the tensor core reads its shared-memory operands through (32 %) --
// and switching the stores off made the FFN1 GEMM 9 % faster.  So the fragment is first transposed
// inside each quad of lanes (32 SHFL): afterwards thread (g, t) owns, for q < 2 and h < 2, the EIGHT
// consecutive columns col0 + 32q + 8t .. +7 of row row0 + g + 8h (T8 layout), and every global access
// of the epilogue is a 16-byte vector: bf16 stores / gate loads touch 4x fewer wavefronts per byte,
// fp32 stores / residual loads / split-K reductions 2x fewer.
//
// The code is organised as one straight-line PASS per epilogue stage with the runtime flag tested
// once per pass: the first version tested every flag for every element pair, unrolled to 4096
// SASS instructions (64 KiB), and the epilogue warps starved on instruction-cache misses
// (ncu: stall_no_inst on every line).
constexpr int kChunkN = 64;

// the problem a work unit belongs to (a grouped launch has one per group)
struct EpiDims {
    int m, n;
    float* out_f32;
    long long ldo_f32;
    bf16* out_bf16;
    long long ldo_bf16;
};

// generic (slow) path for chunks that cross the N boundary or odd N: per element, rolled loops, on the
// UNtransposed fragment
__device__ __noinline__ void epilogue_frag_ragged(const GemmParams& p, const EpiDims& d, const float* v, int lane,
                                                  long long row0, int col0, uint32_t drop_seed, bool addends) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll 1
    for (int i = 0; i < 32; ++i) {
        const int k = i >> 2, h = (i >> 1) & 1, c = i & 1;
        const long long row = row0 + g + 8 * h;
        const int col = col0 + 8 * k + 2 * t + c;
        if (row >= d.m || col >= d.n) continue;
        float x = v[i];
        if (p.bias != nullptr && addends) x += __ldg(p.bias + col);
        if (p.relu) x = fmaxf(x, 0.f);
        if (p.drop_thr != 0)
            x = dropout_u16((uint32_t)(row * (long long)d.n + col), drop_seed) >= p.drop_thr ? x * p.drop_scale : 0.f;
        if (p.gate != nullptr) x = __bfloat162float(p.gate[row * p.ldg + col]) > 0.f ? x * p.gate_scale : 0.f;
        if (p.resid != nullptr && addends) x += p.resid[row * p.ldr + col];
        if (p.colsum != nullptr) atomicAdd(p.colsum + col, x);
        if (d.out_f32 != nullptr) {
            if (p.accumulate) atomicAdd(d.out_f32 + row * d.ldo_f32 + col, x);
            else d.out_f32[row * d.ldo_f32 + col] = x;
        }
        if (d.out_bf16 != nullptr) {
            const bf16 hi = __float2bfloat16_rn(x);
            d.out_bf16[row * d.ldo_bf16 + col] = hi;
            if (p.out_lo != nullptr) p.out_lo[row * d.ldo_bf16 + col] = __float2bfloat16_rn(x - __bfloat162float(hi));
        }
    }
}

// Residual / gate operands of one fragment (T8 layout), fetched from global memory one fragment AHEAD
// of their use so that the ~1 us load latency overlaps the previous fragment's work (and, for the
// first fragment of a tile, the wait for the accumulator).
struct EpiPrefetch {
    uint4 gt[4];      // [2h + q]: 8 bf16 of the gate operand
};

__device__ __forceinline__ void epilogue_prefetch(const GemmParams& p, const EpiDims& d, EpiPrefetch& pf, int lane,
                                                  long long row0, int col0) {
    if (col0 + kChunkN > d.n || ((d.n & 1) && p.drop_thr != 0)) return;   // ragged chunks use the slow path
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const long long row = row0 + g + 8 * h;
        if (row < d.m) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int col = col0 + 32 * q + 8 * t;
                if (p.gate != nullptr)
                    pf.gt[2 * h + q] = __ldg(reinterpret_cast<const uint4*>(p.gate + row * p.ldg + col));
            }
        }
    }
}

// addends: this work unit adds the bias and the residual (false for K splits > 0 of a split-K GEMM
// with a fused LINEAR epilogue: out = resid + keep*scale*(sum_s acc_s + bias), every split scales
// its partial sum, only split 0 contributes the addends; all through red.global.add).
template <bool PF>
__device__ __forceinline__ void epilogue_frag(const GemmParams& p, const EpiDims& d, const uint32_t (&r)[32], int lane,
                                              long long row0, int col0, uint32_t drop_seed,
                                              const EpiPrefetch& pf, bool addends, float (&cs)[16], bool cs_first,
                                              bool cs_flush) {
    const int g = lane >> 2, t = lane & 3;
    // (odd N: only the dropout pair index needs N even; every other access is addressed through the
    // 16-byte aligned leading dimensions)
    if (col0 + kChunkN > d.n || ((d.n & 1) && p.drop_thr != 0)) {
        // only this copy has its address taken; v[] below must stay in registers (an escaping v[]
        // made the compiler mirror it to local memory after every pass: +15 us per epilogue stage)
        float tmp[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) tmp[i] = __uint_as_float(r[i]);
        epilogue_frag_ragged(p, d, tmp, lane, row0, col0, drop_seed, addends);
        return;
    }
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
    quad_transpose(v, t);
    const int col = col0 + 8 * t;   // + 32q + e
    const long long rows[2] = {row0 + g, row0 + g + 8};
    const bool ok[2] = {rows[0] < d.m, rows[1] < d.m};

    if (p.bias != nullptr && addends) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col + 32 * q));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col + 32 * q) + 1);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                T8(v, q, h, 0) += b0.x; T8(v, q, h, 1) += b0.y; T8(v, q, h, 2) += b0.z; T8(v, q, h, 3) += b0.w;
                T8(v, q, h, 4) += b1.x; T8(v, q, h, 5) += b1.y; T8(v, q, h, 6) += b1.z; T8(v, q, h, 7) += b1.w;
            }
        }
    }
    if (p.relu) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    if (p.drop_thr != 0) {
        const uint32_t thr = p.drop_thr;
        const float sc = p.drop_scale;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const uint32_t base = (uint32_t)(rows[h] * (long long)d.n + col + 32 * q) >> 1;   // pair index (n even)
#pragma unroll
                for (int e = 0; e < 8; e += 2) {
                    const uint32_t rnd = dropout_bits_pair(base + (e >> 1), drop_seed);
                    T8(v, q, h, e) = ((rnd & 0xFFFFU) >= thr) ? T8(v, q, h, e) * sc : 0.f;
                    T8(v, q, h, e + 1) = ((rnd >> 16) >= thr) ? T8(v, q, h, e + 1) * sc : 0.f;
                }
            }
        }
    }
    if (PF && p.gate != nullptr) {
        const float gs = p.gate_scale;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (ok[h]) {
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const uint4 gq = pf.gt[2 * h + q];
                    const uint32_t gw[4] = {gq.x, gq.y, gq.z, gq.w};
#pragma unroll
                    for (int e = 0; e < 8; e += 2) {
                        T8(v, q, h, e) = (bf16_lo_to_f(gw[e >> 1]) > 0.f) ? T8(v, q, h, e) * gs : 0.f;
                        T8(v, q, h, e + 1) = (bf16_hi_to_f(gw[e >> 1]) > 0.f) ? T8(v, q, h, e + 1) * gs : 0.f;
                    }
                }
            }
        }
    }
    if (p.debug & 1) {
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += v[i];
        if (acc == 123.456f && d.out_f32 != nullptr) d.out_f32[0] = acc;
        return;
    }
    if (p.colsum != nullptr) {
        // column sums of the epilogue output (the bias gradient of the layer that produced this GEMM's
        // input gradient): the two 16-row fragments of a chunk are summed in registers first, then the 8
        // row groups of the warp through shuffles, then one red.global.add per column and warp
#pragma unroll
        for (int q = 0; q < 2; ++q) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float x = (ok[0] ? T8(v, q, 0, e) : 0.f) + (ok[1] ? T8(v, q, 1, e) : 0.f);
                cs[8 * q + e] = cs_first ? x : cs[8 * q + e] + x;
            }
        }
        if (cs_flush) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float x = cs[i];
                x += __shfl_xor_sync(0xffffffffU, x, 4);
                x += __shfl_xor_sync(0xffffffffU, x, 8);
                x += __shfl_xor_sync(0xffffffffU, x, 16);
                cs[i] = x;
            }
            if (g == 0) {
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    float* o = p.colsum + col + 32 * q;
                    red_add_v4(o, cs[8 * q], cs[8 * q + 1], cs[8 * q + 2], cs[8 * q + 3]);
                    red_add_v4(o + 4, cs[8 * q + 4], cs[8 * q + 5], cs[8 * q + 6], cs[8 * q + 7]);
                }
            }
        }
    }
    if (d.out_bf16 != nullptr) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (ok[h]) {
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    uint32_t pk[4];
#pragma unroll
                    for (int e = 0; e < 8; e += 2) pk[e >> 1] = pack_bf16x2(T8(v, q, h, e), T8(v, q, h, e + 1));
                    *reinterpret_cast<uint4*>(d.out_bf16 + rows[h] * d.ldo_bf16 + col + 32 * q) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    if (p.out_lo != nullptr) {
                        uint32_t lo[4];
#pragma unroll
                        for (int e = 0; e < 8; e += 2)
                            lo[e >> 1] = pack_bf16x2(T8(v, q, h, e) - bf16_lo_to_f(pk[e >> 1]),
                                                     T8(v, q, h, e + 1) - bf16_hi_to_f(pk[e >> 1]));
                        *reinterpret_cast<uint4*>(p.out_lo + rows[h] * d.ldo_bf16 + col + 32 * q) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    }
                }
            }
        }
    }
}

// ---- direct-layout epilogue (fp32 outputs, residual add, split-K reductions) -------------------------
// Measured A/B inside the training step: for fp32 outputs the quad transpose does NOT pay -- these
// epilogues are latency bound (the last tile of a GEMM is not overlapped) and the 32 extra shuffles per
// fragment cost more than the halved wavefront count saves (+2 .. +5 % per GEMM); bf16-only outputs
// gain (QKV forward 953 -> 1037 TFLOP/s).  So: T8 layout for bf16-only outputs, this one otherwise.
// Residual / gate operands of one fragment, fetched from global memory one fragment AHEAD of
// their use so that the ~1 us load latency overlaps the previous fragment's work (and, for the
// first fragment of a tile, the wait for the accumulator).
struct EpiPrefetchDirect {
    float2 res[16];    // [8*h + k]
    uint32_t gt[16];
};

__device__ __forceinline__ void epilogue_prefetch_direct(const GemmParams& p, const EpiDims& d, EpiPrefetchDirect& pf, int lane,
                                                  long long row0, int col0) {
    if (col0 + kChunkN > d.n || ((d.n & 1) && p.drop_thr != 0)) return;   // ragged chunks use the slow path
    const int g = lane >> 2, t = lane & 3;
    const int col = col0 + 2 * t;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const long long row = row0 + g + 8 * h;
        if (row < d.m) {
            if (p.resid != nullptr) {
                const float2* rp = reinterpret_cast<const float2*>(p.resid + row * p.ldr + col);
#pragma unroll
                for (int k = 0; k < 8; ++k) pf.res[8 * h + k] = rp[4 * k];
            }
            if (p.gate != nullptr) {
                const uint32_t* gp = reinterpret_cast<const uint32_t*>(p.gate + row * p.ldg + col);
#pragma unroll
                for (int k = 0; k < 8; ++k) pf.gt[8 * h + k] = __ldg(gp + 4 * k);
            }
        }
    }
}

// addends: this work unit adds the bias and the residual (false for K splits > 0 of a split-K GEMM
// with a fused LINEAR epilogue: out = resid + keep*scale*(sum_s acc_s + bias), every split scales
// its partial sum, only split 0 contributes the addends; all through red.global.add).
template <bool PF>
__device__ __forceinline__ void epilogue_frag_direct(const GemmParams& p, const EpiDims& d, const uint32_t (&r)[32], int lane,
                                              long long row0, int col0, uint32_t drop_seed,
                                              const EpiPrefetchDirect& pf, bool addends) {
    const int g = lane >> 2, t = lane & 3;
    // (odd N: only the dropout pair index needs N even; every other access is addressed through the
    // 16-byte aligned leading dimensions)
    if (col0 + kChunkN > d.n || ((d.n & 1) && p.drop_thr != 0)) {
        // only this copy has its address taken; v[] below must stay in registers (an escaping v[]
        // made the compiler mirror it to local memory after every pass: +15 us per epilogue stage)
        float tmp[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) tmp[i] = __uint_as_float(r[i]);
        epilogue_frag_ragged(p, d, tmp, lane, row0, col0, drop_seed, addends);
        return;
    }
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
    const int col = col0 + 2 * t;   // + 8k
    const long long rows[2] = {row0 + g, row0 + g + 8};
    const bool ok[2] = {rows[0] < d.m, rows[1] < d.m};

    if (p.bias != nullptr && addends) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float2 b = __ldg(reinterpret_cast<const float2*>(p.bias + col + 8 * k));
            v[4 * k] += b.x; v[4 * k + 1] += b.y; v[4 * k + 2] += b.x; v[4 * k + 3] += b.y;
        }
    }
    if (p.relu) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    if (p.drop_thr != 0) {
        const uint32_t thr = p.drop_thr;
        const float sc = p.drop_scale;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t base = (uint32_t)(rows[h] * (long long)d.n + col) >> 1;   // pair index (n even)
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t rnd = dropout_bits_pair(base + 4 * k, drop_seed);
                v[4 * k + 2 * h] = ((rnd & 0xFFFFU) >= thr) ? v[4 * k + 2 * h] * sc : 0.f;
                v[4 * k + 2 * h + 1] = ((rnd >> 16) >= thr) ? v[4 * k + 2 * h + 1] * sc : 0.f;
            }
        }
    }
    if (PF && p.gate != nullptr) {
        const float gs = p.gate_scale;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (ok[h]) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint32_t gt = pf.gt[8 * h + k];
                    v[4 * k + 2 * h] = (bf16_lo_to_f(gt) > 0.f) ? v[4 * k + 2 * h] * gs : 0.f;
                    v[4 * k + 2 * h + 1] = (bf16_hi_to_f(gt) > 0.f) ? v[4 * k + 2 * h + 1] * gs : 0.f;
                }
            }
        }
    }
    if (PF && p.resid != nullptr && addends) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (ok[h]) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    v[4 * k + 2 * h] += pf.res[8 * h + k].x;
                    v[4 * k + 2 * h + 1] += pf.res[8 * h + k].y;
                }
            }
        }
    }
    if (p.debug & 1) {
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += v[i];
        if (acc == 123.456f && d.out_f32 != nullptr) d.out_f32[0] = acc;
        return;
    }
    if (d.out_f32 != nullptr) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (ok[h]) {
                float* o = d.out_f32 + rows[h] * d.ldo_f32 + col;
                if (p.accumulate) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) red_add_v2(o + 8 * k, v[4 * k + 2 * h], v[4 * k + 2 * h + 1]);
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        *reinterpret_cast<float2*>(o + 8 * k) = make_float2(v[4 * k + 2 * h], v[4 * k + 2 * h + 1]);
                }
            }
        }
    }
    if (d.out_bf16 != nullptr) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (ok[h]) {
                uint32_t* o = reinterpret_cast<uint32_t*>(d.out_bf16 + rows[h] * d.ldo_bf16 + col);
                uint32_t pk[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    pk[k] = pack_bf16x2(v[4 * k + 2 * h], v[4 * k + 2 * h + 1]);
                    o[4 * k] = pk[k];
                }
                if (p.out_lo != nullptr) {
                    uint32_t* ol = reinterpret_cast<uint32_t*>(p.out_lo + rows[h] * d.ldo_bf16 + col);
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        ol[4 * k] = pack_bf16x2(v[4 * k + 2 * h] - bf16_lo_to_f(pk[k]),
                                                v[4 * k + 2 * h + 1] - bf16_hi_to_f(pk[k]));
                }
            }
        }
    }
}

// Epilogue of one 128-row output tile for one of the 8 epilogue warps.  PF = the residual / gate
// operands are prefetched one fragment ahead (separate instantiation so that plain epilogues do
// not carry the prefetch registers).
template <int CG, bool PF, bool T8L>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, const EpiDims& d, int lane, int chunk_par,
                                              long long row0, int n0, int width, uint32_t taddr,
                                              uint64_t* full_bar, uint64_t* empty_bar,
                                              uint32_t acc_phase, uint32_t drop_seed, uint32_t lead_rank,
                                              bool addends) {
    using Pref = typename std::conditional<T8L, EpiPrefetch, EpiPrefetchDirect>::type;
    const int nchunks = min(width / kChunkN, (d.n - n0 + kChunkN - 1) / kChunkN);
    const int last_c = ((nchunks - 1 - chunk_par) & ~1) + chunk_par;   // this warp's last chunk (< 0: none)
    Pref pf_next;
    if (PF && last_c >= 0) {   // operands of the first fragment, before waiting for the MMAs
        if constexpr (T8L) epilogue_prefetch(p, d, pf_next, lane, row0, n0 + chunk_par * kChunkN);
        else epilogue_prefetch_direct(p, d, pf_next, lane, row0, n0 + chunk_par * kChunkN);
    }
    mbar_wait(full_bar, acc_phase);
    tc_fence_after();
    if (last_c < 0) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
            if (CG == 2) mbar_arrive_remote(empty_bar, lead_rank);
            else mbar_arrive(empty_bar);
        }
    }
    float cs[16];   // column sums of the current chunk (T8 epilogue with colsum)
#pragma unroll 1
    for (int c = chunk_par; c < nchunks; c += 2) {
#pragma unroll 1
        for (int hb = 0; hb < 2; ++hb) {
            uint32_t r[32];
            tmem_ld_16x256b_x8(taddr + ((uint32_t)(hb * 16) << 16) + (uint32_t)(c * kChunkN), r);
            Pref pf;
            if (PF) {
                pf = pf_next;
                const int nc = hb ? c + 2 : c, nhb = hb ^ 1;    // next fragment of this tile
                if (nc < nchunks && row0 + nhb * 16 < d.m) {
                    if constexpr (T8L) epilogue_prefetch(p, d, pf_next, lane, row0 + nhb * 16, n0 + nc * kChunkN);
                    else epilogue_prefetch_direct(p, d, pf_next, lane, row0 + nhb * 16, n0 + nc * kChunkN);
                }
            }
            tmem_ld_wait();
            if (c == last_c && hb == 1) {
                // accumulator stage fully drained: hand it back to the MMA warp now
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (CG == 2) mbar_arrive_remote(empty_bar, lead_rank);
                    else mbar_arrive(empty_bar);
                }
            }
            if (row0 + hb * 16 < d.m) {
                if constexpr (T8L)
                    epilogue_frag<PF>(p, d, r, lane, row0 + hb * 16, n0 + c * kChunkN, drop_seed, pf, addends, cs, hb == 0,
                                      hb == 1 || row0 + 16 >= d.m);
                else epilogue_frag_direct<PF>(p, d, r, lane, row0 + hb * 16, n0 + c * kChunkN, drop_seed, pf, addends);
            }
        }
    }
}

// MC = CTA pairs per cluster (CG == 2 only).  MC == 2: a cluster of 4 CTAs computes a 512 x 256
// super-tile -- two pairs on vertically adjacent 256-row tiles that need the SAME B tile.  Each
// CTA then fetches only a quarter of B (64 rows) and TMA-multicasts it to its counterpart in the
// other pair: L2 -> shared-memory traffic per CTA drops from 32 KB to 24 KB per k-block, and that
// traffic (not the tensor pipe) is what bounds this kernel.
// work unit -> (output tile, K split, first column offset inside the tile, width in columns)
struct UnitInfo { int tile, split, ncol, width, group, mt, nt; };    // (mt, nt): tile coordinates inside its problem
template <int BLOCK_N>
__device__ __forceinline__ UnitInfo decode_unit(const GemmParams& p, int unit, int tiles) {
    UnitInfo u;
    u.group = 0;
    if (p.full_tiles >= tiles) {
        u.tile = unit % tiles;
        u.split = unit / tiles;
        u.ncol = 0;
        u.width = BLOCK_N;
    } else if (unit < p.full_tiles) {
        u.tile = unit;
        u.split = 0;
        u.ncol = 0;
        u.width = BLOCK_N;
    } else {
        const int r = unit - p.full_tiles;
        u.tile = p.full_tiles + (r >> 1);
        u.split = 0;
        u.ncol = (r & 1) * (BLOCK_N / 2);
        u.width = BLOCK_N / 2;
    }
    if (p.num_groups > 0) {
        int g = 0;
        while (g + 1 < p.num_groups && u.tile >= p.grp[g + 1].tile_start) ++g;
        const int local = u.tile - p.grp[g].tile_start;
        u.group = g;
        u.mt = local / p.grp[g].n_tiles;
        u.nt = local - u.mt * p.grp[g].n_tiles;
    } else {
        u.mt = u.tile / p.n_tiles;
        u.nt = u.tile - u.mt * p.n_tiles;
    }
    return u;
}

// EPI = 1: bf16-only outputs, quad-transposed (T8) epilogue with 16-byte accesses; EPI = 0: fp32 outputs /
// residual / split-K reductions in the direct layout (see the A/B note above epilogue_frag_direct).  A template
// parameter, not a runtime branch: with both epilogues inlined into one kernel ptxas spilled 100+ bytes.
template <int BLOCK_N, int A_MN, int B_MN, int CG, int MC, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ GemmParams p) {
    using Cfg = GemmCfg<BLOCK_N, CG>;
    static_assert(MC == 1 || (MC == 2 && CG == 2 && BLOCK_N == 256), "pair multicast needs 256-wide CTA-pair tiles");
    constexpr int kStages = Cfg::kStages;
    constexpr int kBRows = BLOCK_N / CG;   // rows of the B tile staged in this CTA's shared memory
    constexpr int CL = CG * MC;            // CTAs per cluster

    // SWIZZLE_128B tiles need 1024-byte alignment; the kernel has no static shared memory, so the
    // dynamic window starts at the (1024-aligned) base of the CTA's shared memory.
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    pdl_launch_dependents();   // the next kernel's launch + prologue may overlap this grid (see common.cuh)
    if ((smem_u32(smem) & 1023U) != 0) {
        if (threadIdx.x == 0) printf("mcan gemm: dynamic smem base not 1024-byte aligned\n");
        __trap();
    }
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full_bar = empty_bar + kStages;
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;
    uint64_t* sched_full = tmem_empty_bar + 2;
    uint64_t* sched_empty = sched_full + kSchedStages;
    uint32_t* sched_unit = reinterpret_cast<uint32_t*>(sched_empty + kSchedStages);
    uint32_t* tmem_slot = sched_unit + kSchedStages;

    // Warp-uniform role index: the shuffle makes it provably uniform for ptxas, so the single-thread
    // roles below run their loops on the uniform datapath (UTCHMMA / UTMALDG / UTCBAR take their
    // operands straight from uniform registers).  Round 1 ran these roles as `if (lane == 0)` code:
    // every MMA then sat in an ELECT + 5 x R2UR + BRA.U.ANY waterfall loop, ~200 cycles per MMA.
    const int warp = __shfl_sync(0xffffffffU, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const uint32_t crank = (CL > 1) ? cluster_ctarank() : 0U;   // position in the cluster
    const uint32_t rank = crank & (uint32_t)(CG - 1);             // position in the CTA pair
    const uint32_t pair = crank / CG;                             // which pair of the cluster (MC == 2)
    const uint32_t lead_rank = pair * CG;                         // cluster rank of this pair's leader
    const bool leader = (rank == 0);                              // issues the pair's MMAs, owns its barriers
    const bool cl_leader = (crank == 0);                          // runs the dynamic tile scheduler

    if (warp == kProducerWarp && lane == 0) {
        const int nmaps = p.num_groups > 0 ? p.num_groups : p.num_seg;
        for (int s = 0; s < nmaps; ++s) {
            prefetch_tmap(&p.tma_a[s]);
            prefetch_tmap(&p.tma_b[s]);
        }
    }
    if (warp == kMmaWarp) {
        if (lane == 0) {
            for (int s = 0; s < kStages; ++s) {
                mbar_init(&full_bar[s], 1);    // CG==2: only the leader's is used (bytes of both CTAs)
                mbar_init(&empty_bar[s], MC);  // one commit per pair whose MMAs read (multicast) data of this slot
            }
            for (int s = 0; s < 2; ++s) {
                mbar_init(&tmem_full_bar[s], 1);
                mbar_init(&tmem_empty_bar[s], kEpilogueWarps * CG);   // CG==2: leader's, both CTAs arrive
            }
            for (int s = 0; s < kSchedStages; ++s) {
                mbar_init(&sched_full[s], 1);
                // readers: producer + MMA warp + epilogue warps of the leader, producer + epilogue warps of the peer
                mbar_init(&sched_empty[s], ((2 + kEpilogueWarps) + (CG == 2 ? 1 + kEpilogueWarps : 0)) * MC);
            }
            fence_mbar_init();
        }
        __syncwarp();
        if (CG == 2) {
            tmem_alloc_cg2(tmem_slot, Cfg::kTmemCols);
            tmem_relinquish_cg2();
        } else {
            tmem_alloc(tmem_slot, Cfg::kTmemCols);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // everything above (barriers, TMEM, descriptor prefetch) overlapped the previous kernel's tail;
    // from here on global memory is touched: wait for the prerequisite grids to complete
    pdl_wait();

    const int tiles = p.m_tiles * p.n_tiles;      // m_tiles counts (128*CG*MC)-row (super-)tiles
    const int units = p.units;                    // tiles * splits, or full_tiles + 2 * (tiles - full_tiles)
    const int nclusters = gridDim.x / CL;
    // p.tile_counter == nullptr: static schedule (cluster c takes units c, c + nclusters, ...), no
    // atomics and no broadcast on the critical path -- the default when nothing else shares the GPU.
    const bool dyn = p.tile_counter != nullptr;
    int sunit = (int)blockIdx.x / CL;
    int sslot = 0;            // position in the work-unit ring (every role walks it in lock step)
    uint32_t sphase = 0;
    // consumer side of the ring, called by ALL lanes of a role's warp: returns the next work unit
    // (>= units: no more work); one arrival per warp
    auto next_unit = [&]() -> int {
        if (!dyn) {
            const int u = sunit;
            sunit += nclusters;
            return u;
        }
        // the leader's own consumers see a local write (CTA scope is enough and much cheaper);
        // the peer's consumers need cluster-scope acquire / release
        if (CL > 1 && !cl_leader) mbar_wait_acq_cluster(&sched_full[sslot], sphase);
        else mbar_wait(&sched_full[sslot], sphase);
        int u = (int)*reinterpret_cast<volatile uint32_t*>(&sched_unit[sslot]);
        u = __shfl_sync(0xffffffffU, u, 0);
        if (lane == 0) {     // (the shuffle above ordered every lane's read before this arrival)
            if (CL > 1 && !cl_leader) mbar_arrive_release_cluster(&sched_empty[sslot], 0);
            else mbar_arrive(&sched_empty[sslot]);
        }
        if (++sslot == kSchedStages) { sslot = 0; sphase ^= 1; }
        return u;
    };

    if (warp == kProducerWarp) {
        // ===================== TMA producer (one elected lane issues) =====================
        int stage = 0;
        uint32_t phase = 0;
        while (true) {
            const int unit = next_unit();
            if (unit >= units) break;
            const UnitInfo ui = decode_unit<BLOCK_N>(p, unit, tiles);
            const int split = ui.split;
            const bool half = ui.width != BLOCK_N;          // half-width unit: this CTA stages kBRows / 2 rows of B
            const int brows = half ? kBRows / 2 : kBRows;
            const uint32_t stage_tx = Cfg::kABytes + (uint32_t)brows * BLOCK_K * 2;
            const int m0 = ui.mt * (BLOCK_M * CL) + (int)crank * BLOCK_M;
            // MC == 2: this CTA fetches rows [pair*64, pair*64+64) of its half of B for both pairs
            const int n0 = ui.nt * BLOCK_N + ui.ncol + (int)rank * brows + (MC == 2 ? (int)pair * (kBRows / 2) : 0);
            const uint16_t mc_mask = (uint16_t)((1U << rank) | (1U << (CG + rank)));
            const int kb0 = (int)((long long)p.kblocks * split / p.splits);
            const int kb1 = (int)((long long)p.kblocks * (split + 1) / p.splits);
            for (int seg = 0; seg < p.num_seg; ++seg) {
                const CUtensorMap* ta = &p.tma_a[seg + ui.group];      // (a grouped launch has one segment)
                const CUtensorMap* tb = (half && !B_MN) ? &p.tma_b_half[seg] : &p.tma_b[seg + ui.group];
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (p.debug & 2) {
                        if (lane == 0 && (CG == 1 || leader)) mbar_arrive(&full_bar[stage]);
                    } else if (elect_one()) {
                        uint8_t* sa = smem + stage * Cfg::kStageBytes;
                        uint8_t* sb = sa + Cfg::kABytes;
                        if (CG == 1) mbar_arrive_expect_tx(&full_bar[stage], stage_tx);
                        else if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * stage_tx);
                        auto load = [&](void* dst, const CUtensorMap* m, int c0, int c1) {
                            if (CG == 2) tma_load_2d_cg2(dst, m, &full_bar[stage], c0, c1);
                            else tma_load_2d(dst, m, &full_bar[stage], c0, c1);
                        };
                        if (A_MN) {
#pragma unroll
                            for (int c = 0; c < BLOCK_M / 64; ++c)
                                load(sa + c * (BLOCK_K * 128), ta, m0 + c * 64, kb * BLOCK_K);
                        } else {
                            load(sa, ta, kb * BLOCK_K, m0);
                        }
                        if (MC == 2) {
                            // one 64-row (K-major) / 64-column (MN-major) box = 8 KiB, at the same offset in both pairs
                            uint8_t* dst = sb + pair * (BLOCK_K * 128);
                            if (B_MN) tma_load_2d_cg2_mc(dst, tb, &full_bar[stage], n0, kb * BLOCK_K, mc_mask);
                            else tma_load_2d_cg2_mc(dst, tb, &full_bar[stage], kb * BLOCK_K, n0, mc_mask);
                        } else if (B_MN) {
#pragma unroll
                            for (int c = 0; c < kBRows / 64; ++c)
                                if (c * 64 < brows) load(sb + c * (BLOCK_K * 128), tb, n0 + c * 64, kb * BLOCK_K);
                        } else {
                            load(sb, tb, kb * BLOCK_K, n0);
                        }
                    }
                    __syncwarp();
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
            } else if (warp == kMmaWarp) {
        // ===================== MMA issuer (one elected lane of the leader CTA) =====================
        if (leader) {
            constexpr uint32_t idesc_full = make_idesc_bf16(BLOCK_M * CG, BLOCK_N, A_MN, B_MN);
            constexpr uint32_t idesc_half = make_idesc_bf16(BLOCK_M * CG, BLOCK_N / 2, A_MN, B_MN);
            constexpr uint32_t a_lbo = A_MN ? BLOCK_K * 128 : 0;
            constexpr uint32_t b_lbo = B_MN ? BLOCK_K * 128 : 0;
            constexpr uint32_t a_kstep = A_MN ? (UMMA_K * 128) : (UMMA_K * 2);
            constexpr uint32_t b_kstep = B_MN ? (UMMA_K * 128) : (UMMA_K * 2);
            // shared-memory matrix descriptors: the high word is constant (SBO = 1024 B, version 1,
            // SWIZZLE_128B); the low word = start address >> 4 | LBO >> 4 << 16 advances per stage / k-step
            constexpr uint32_t desc_hi = (uint32_t)(make_smem_desc_sw128_const(0, 0, 1024) >> 32);
            const uint32_t smem0 = smem_u32(smem);
            const uint32_t a_lo0 = ((smem0 & 0x3FFFFU) >> 4) | ((a_lbo >> 4) << 16);
            const uint32_t b_lo0 = (((smem0 + Cfg::kABytes) & 0x3FFFFU) >> 4) | ((b_lbo >> 4) << 16);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            int unit = next_unit();
            while (unit < units) {
                // units are published one tile ahead: fetch the next one now, off the critical path
                const int unit_after = next_unit();
                const UnitInfo ui = decode_unit<BLOCK_N>(p, unit, tiles);
                const int split = ui.split;
                const uint32_t idesc = (ui.width == BLOCK_N) ? idesc_full : idesc_half;
                const int kb0 = (int)((long long)p.kblocks * split / p.splits);
                const int kb1 = (int)((long long)p.kblocks * (split + 1) / p.splits);
                const int iters = (kb1 - kb0) * p.num_seg;
                mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
                for (int it = 0; it < iters; ++it) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t a_lo = a_lo0 + (uint32_t)stage * (Cfg::kStageBytes >> 4);
                        const uint32_t b_lo = b_lo0 + (uint32_t)stage * (Cfg::kStageBytes >> 4);
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                            if (p.debug & 4) break;
                            const uint64_t ad = ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo + ((k * a_kstep) >> 4));
                            const uint64_t bd = ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + ((k * b_kstep) >> 4));
                            const uint32_t accum = (it > 0 || k > 0) ? 1U : 0U;
                            if (CG == 2) umma_bf16_cg2(tmem_d, ad, bd, idesc, accum);
                            else umma_bf16(tmem_d, ad, bd, idesc, accum);
                        }
                        // frees the smem slot (in both CTAs) when the MMAs retire
                        // (MC == 2: in all four CTAs -- the other pair multicasts into our slot and vice versa)
                        if (CG == 2) umma_commit_cg2(&empty_bar[stage], (uint16_t)((1U << CL) - 1U)); else umma_commit(&empty_bar[stage]);
                        // accumulator ready for the epilogue warps (of both CTAs)
                        if (it == iters - 1) {
                            if (CG == 2) umma_commit_cg2(&tmem_full_bar[acc], (uint16_t)(3U << lead_rank)); else umma_commit(&tmem_full_bar[acc]);
                        }
                    }
                    __syncwarp();
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                unit = unit_after;
            }
        }
    } else if (warp == kSchedWarp) {
        // ===================== tile scheduler (dynamic schedule, leader CTA only) =====================
        // Claims work units from the global counter and publishes them to every role of the CTA
        // (pair) through the ring; runs kSchedStages units ahead of the slowest reader, so the
        // atomic's round trip to L2 and the cross-CTA broadcast never stall a tile.  Exactly one
        // failing claim per cluster; the very last claim of the grid resets the counter.
        if (dyn && cl_leader && lane == 0) {
            int slot = 0;
            uint32_t ph = 0;
            while (true) {
                const int unit = atomicAdd(p.tile_counter, 1);
                if (unit == units + nclusters - 1) atomicExch(p.tile_counter, 0);
                mbar_wait(&sched_empty[slot], ph ^ 1);
                sched_unit[slot] = (uint32_t)unit;
                if (CL > 1) {
#pragma unroll
                    for (int c = 1; c < CL; ++c) {
                        st_shared_remote_u32(&sched_unit[slot], (uint32_t)c, (uint32_t)unit);
                        mbar_arrive_release_cluster(&sched_full[slot], (uint32_t)c);
                    }
                    mbar_arrive_release_cluster(&sched_full[slot], 0);
                } else {
                    mbar_arrive(&sched_full[slot]);
                }
                if (++slot == kSchedStages) { slot = 0; ph ^= 1; }
                if (unit >= units) break;
            }
        }
    } else {
        // ===================== epilogue: TMEM -> registers -> global =====================
        const int quad = warp & 3;            // TMEM lanes [32*quad, 32*quad+32) belong to this warp
        const int chunk_par = warp >> 2;      // which of the two warps of this quarter: even / odd chunks
        const uint32_t drop_seed =
            p.drop_seed ^ ((p.drop_thr != 0 && p.drop_seed_dev != nullptr) ? __ldg(p.drop_seed_dev) : 0U);
        int acc = 0;
        uint32_t acc_phase = 0;
        int unit = next_unit();
        while (unit < units) {
            const int unit_after = next_unit();     // published one tile ahead
            const UnitInfo ui = decode_unit<BLOCK_N>(p, unit, tiles);
            const int m0 = ui.mt * (BLOCK_M * CL) + (int)crank * BLOCK_M;
            const int n0 = ui.nt * BLOCK_N + ui.ncol;
            const long long row0 = m0 + quad * 32;
            EpiDims d;
            if (p.num_groups > 0) {
                d.m = p.grp[ui.group].m; d.n = p.grp[ui.group].n;
                d.out_f32 = p.group_bf16 ? nullptr : reinterpret_cast<float*>(p.grp[ui.group].out);
                d.out_bf16 = p.group_bf16 ? reinterpret_cast<bf16*>(p.grp[ui.group].out) : nullptr;
                d.ldo_f32 = d.ldo_bf16 = p.grp[ui.group].ldo;
            } else {
                d.m = p.m; d.n = p.n; d.out_f32 = p.out_f32; d.ldo_f32 = p.ldo_f32;
                d.out_bf16 = p.out_bf16; d.ldo_bf16 = p.ldo_bf16;
            }
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BLOCK_N);
#define MCAN_EPI_TILE(PF, T8L) epilogue_tile<CG, PF, T8L>(p, d, lane, chunk_par, row0, n0, ui.width, taddr, &tmem_full_bar[acc], \
                                                          &tmem_empty_bar[acc], acc_phase, drop_seed, lead_rank, ui.split == 0)
            if (EPI == 1) {
                if (p.gate != nullptr) MCAN_EPI_TILE(true, true); else MCAN_EPI_TILE(false, true);
            } else {
                if (p.resid != nullptr || p.gate != nullptr) MCAN_EPI_TILE(true, false); else MCAN_EPI_TILE(false, false);
            }
#undef MCAN_EPI_TILE
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            unit = unit_after;
        }
    }

    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();   // the peer's smem / barriers stay alive until here
    if (warp == kMmaWarp) {
        tc_fence_after();
        if (CG == 2) tmem_dealloc_cg2(tmem_base, Cfg::kTmemCols);
        else tmem_dealloc(tmem_base, Cfg::kTmemCols);
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, []() {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) ==
                cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(ptr);
    });
    return fn;
}

struct TmapKey {
    const void* ptr;
    uint64_t inner, outer, ld;
    uint32_t box_rows;
    bool operator==(const TmapKey& o) const {
        return ptr == o.ptr && inner == o.inner && outer == o.outer && ld == o.ld &&
               box_rows == o.box_rows;
    }
};
struct TmapKeyHash {
    size_t operator()(const TmapKey& k) const {
        uint64_t h = reinterpret_cast<uint64_t>(k.ptr) * 0x9E3779B97F4A7C15ULL;
        h ^= k.inner + 0x9E3779B97F4A7C15ULL + (h << 6) + (h >> 2);
        h ^= k.outer + 0x9E3779B97F4A7C15ULL + (h << 6) + (h >> 2);
        h ^= k.ld + 0x9E3779B97F4A7C15ULL + (h << 6) + (h >> 2);
        h ^= k.box_rows + 0x9E3779B97F4A7C15ULL + (h << 6) + (h >> 2);
        return (size_t)h;
    }
};

// bf16 2-D tensor map over a row-major [outer, inner] array with leading dimension ld
// (elements); box = {64 inner elements (=128 B, the swizzle span), box_rows}.
int make_tmap_bf16(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld,
                   uint32_t box_rows) {
    static std::mutex mu;
    static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
    TmapKey key{ptr, inner, outer, ld, box_rows};
    {
        std::lock_guard<std::mutex> g(mu);
        auto it = cache.find(key);
        if (it != cache.end()) {
            *out = it->second;
            return 0;
        }
    }
    PFN_encodeTiled enc = get_encode_fn();
    MCAN_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available (no CUDA driver / no GPU)");
    MCAN_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "GEMM operand %p not 16-byte aligned", ptr);
    MCAN_REQUIRE((ld * 2) % 16 == 0, "GEMM operand leading dimension %llu not a multiple of 8",
                 (unsigned long long)ld);
    cuuint64_t gdim[2] = {inner, outer};
    cuuint64_t gstride[1] = {ld * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride,
                     box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MCAN_REQUIRE(r == CUDA_SUCCESS,
                 "cuTensorMapEncodeTiled failed (%d) ptr=%p inner=%llu outer=%llu ld=%llu box=%u",
                 (int)r, ptr, (unsigned long long)inner, (unsigned long long)outer,
                 (unsigned long long)ld, box_rows);
    {
        std::lock_guard<std::mutex> g(mu);
        if (cache.size() > 8192) cache.clear();
        cache.emplace(key, *out);
    }
    return 0;
}

static int g_sm_limit = 0;   // 0 = use every SM (set through mcan_set_sm_limit)

int device_num_sms_raw();

int device_num_sms() {
    const int raw = device_num_sms_raw();
    return (g_sm_limit > 0 && g_sm_limit < raw) ? g_sm_limit : raw;
}

int device_num_sms_raw() {
    static int sms[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
    if (sms[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
        sms[dev] = v;
    }
    return sms[dev];
}

// 0: static round-robin tile schedule, 1: dynamic (work-unit counter).  See mcan_set_gemm_schedule.
static std::atomic<int> g_dynamic_schedule{0};
// MCAN_GEMM_TAIL_SPLIT=0 disables tail splitting (A/B timing)
static std::atomic<int> g_tail_split{[] { const char* e = getenv("MCAN_GEMM_TAIL_SPLIT"); return (e && e[0] == '0') ? 0 : 1; }()};

// Pool of zero-initialised work-unit counters (one per in-flight launch; each kernel resets its own
// counter with its last claim).  Besides the debug sink this pool is the only device memory the library owns.
static int next_tile_counter(int** out) {
    constexpr int kPool = 256;
    static int* pool[64] = {nullptr};
    static unsigned next[64] = {0};
    int dev = 0;
    MCAN_CHECK_CUDA(cudaGetDevice(&dev));
    MCAN_REQUIRE(dev >= 0 && dev < 64, "device index %d", dev);
    if (pool[dev] == nullptr) {
        MCAN_CHECK_CUDA(cudaMalloc(&pool[dev], kPool * sizeof(int)));
        MCAN_CHECK_CUDA(cudaMemset(pool[dev], 0, kPool * sizeof(int)));
        MCAN_CHECK_CUDA(cudaDeviceSynchronize());
    }
    *out = pool[dev] + (next[dev]++ % kPool);
    return 0;
}

template <int BLOCK_N, int A_MN, int B_MN, int CG, int MC, int EPI>
static int launch_gemm(const GemmParams& p, int64_t units, int sms, cudaStream_t stream) {
    using Cfg = GemmCfg<BLOCK_N, CG>;
    constexpr int CL = CG * MC;
    static bool configured[64] = {false};
    static int max_clusters[64] = {0};
    int dev = 0;
    MCAN_CHECK_CUDA(cudaGetDevice(&dev));
    MCAN_REQUIRE(dev >= 0 && dev < 64, "device index %d", dev);
    auto kernel = gemm_tcgen05_kernel<BLOCK_N, A_MN, B_MN, CG, MC, EPI>;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.blockDim = dim3(kGemmThreads, 1, 1);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    if (!configured[dev]) {
        MCAN_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)Cfg::kSmemBytes));
        // how many clusters of this shape fit on the device at once (GPC boundaries cost a few SMs
        // for clusters of 4): a persistent grid must be fully co-resident
        int n = 0;
        cfg.gridDim = dim3((unsigned)(device_num_sms_raw() / CL * CL), 1, 1);
        cfg.numAttrs = 1;
        if (CL > 2) {
            MCAN_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&n, kernel, &cfg));
            MCAN_REQUIRE(n > 0, "mcan_gemm: no cluster of %d CTAs fits", CL);
        } else {
            n = device_num_sms_raw() / CL;
        }
        max_clusters[dev] = n;
        configured[dev] = true;
    }
    int64_t slots = sms / CL;
    if (slots > max_clusters[dev]) slots = max_clusters[dev];
    const int grid = (int)(units < slots ? units : slots) * CL;
    cfg.gridDim = dim3((unsigned)grid, 1, 1);
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    MCAN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, p));
    return 0;
}

// Tile configuration = (cta_group, BLOCK_N).  Cost model: number of waves over the SMs x the time
// of one tile; relative tile throughputs measured on B200 (tools/gemm_bench.py): the CTA-pair
// 256x256 tile has the highest arithmetic intensity, the single-CTA 128x128 tile the lowest.
struct TileCfg { int cg, bn; };

static TileCfg pick_tile(int64_t m, int64_t n, int64_t k, bool accumulate, int sms) {
    // {1, 64}: short-M GEMMs (the 896-row question side) are latency bound -- twice as many,
    // half as long tiles
    const TileCfg cand[5] = {{2, 256}, {2, 128}, {1, 256}, {1, 128}, {1, 64}};
    const double rate[5] = {1.00, 0.62, 0.85, 0.62, 0.45};
    TileCfg best = cand[3];
    double best_cost = 1e30;
    // deep split-K GEMMs (wgrad over 6400 rows): K splits fill the machine whatever the tile, so take
    // the tile with the highest arithmetic intensity (1024x1024x6400: 23.5 us vs 27.7 us with 256x128)
    // ... but only with many rows: for the 896-row split-K dgrads of the question side (K 3072 / 4096 / 12288) the
    // single-CTA 128 x 128 tile is 17-24 % faster than the pair tile (profiles/r02_tile_sweep.txt: 25.6 -> 21.5 us,
    // 25.5 -> 19.3 us): 7 x 8 tiles x the K splits fill the machine without padding 896 rows to 1024
    static const bool short_m_rule = [] { const char* e = getenv("MCAN_GEMM_SHORT_M_RULE"); return !(e && e[0] == '0'); }();
    if (short_m_rule && accumulate && k >= 2048 && m < 1024 && m > 128 && m % 256 != 0 && n >= 256) return cand[3];
    if (accumulate && k >= 2048 && m >= 256 && n >= 256) return cand[0];
    for (int i = 0; i < 5; ++i) {
        const int cg = cand[i].cg, bn = cand[i].bn;
        if (bn == 256 && n <= 128) continue;
        if (cg == 2 && m <= 128) continue;
        // 128 x 64 tiles re-read the operands from L2 most often: only for short contractions
        if (bn == 64 && (k > 4096 || (accumulate && k > 1024))) continue;
        const int64_t tiles = ((m + 128 * cg - 1) / (128 * cg)) * ((n + bn - 1) / bn);
        const int64_t slots = sms / cg;
        const int64_t waves = (tiles + slots - 1) / slots;
        const double cost = (double)waves * (128.0 * bn) / rate[i];
        if (cost < best_cost) { best_cost = cost; best = cand[i]; }
    }
    return best;
}

static int pick_splits(int64_t tiles, int kblocks, int slots) {
    int best = 1;
    double best_cost = 1e30;
    const int max_s = kblocks < 32 ? kblocks : 32;
    for (int s = 1; s <= max_s; ++s) {
        const int64_t units = tiles * s;
        const int64_t waves = (units + slots - 1) / slots;
        const double per_unit = (double)((kblocks + s - 1) / s) + 6.0;  // +epilogue/fill overhead
        const double cost = (double)waves * per_unit;
        if (cost < best_cost * 0.98) {
            best_cost = cost;
            best = s;
        }
    }
    return best;
}

// Launch plan of one GEMM: tile shape, cluster shape, K splits, tail splitting.  Pure host logic
// (exposed as mcan_gemm_plan so that it can be tested without a GPU).
struct GemmPlan {
    int block_n, cl, cg, mc;        // cl = CTAs per cluster (1, 2, 4), cg = CTAs per MMA, mc = pairs per cluster
    int m_tiles, n_tiles, splits, full_tiles, units;
};

static int plan_gemm(int64_t m, int64_t n, int64_t k, bool accumulate, int split_k, int block_n_req, int cg_req,
                     int sms, GemmPlan* out) {
    const int kblocks = (int)((k + BLOCK_K - 1) / BLOCK_K);
    TileCfg tc = pick_tile(m, n, k, accumulate, sms);
    if (block_n_req) tc.bn = block_n_req;
    if (cg_req) tc.cg = cg_req;
    if (tc.bn == 64 && tc.cg != 1 && !block_n_req) tc.bn = 128;     // forced cta_group: 64-wide tiles are single-CTA
    // cta_group 4 = CTA pairs (cta_group::2 MMAs) in clusters of two pairs that share the B tile by multicast
    const int block_n = tc.bn, cl = tc.cg, cg = cl == 4 ? 2 : cl, mc = cl == 4 ? 2 : 1;
    MCAN_REQUIRE(block_n == 64 || block_n == 128 || block_n == 256, "mcan_gemm: block_n=%d", block_n);
    MCAN_REQUIRE(block_n != 64 || cl == 1, "mcan_gemm: block_n 64 is a single-CTA tile");
    MCAN_REQUIRE(cl == 1 || cl == 2 || cl == 4, "mcan_gemm: cta_group=%d", cl);
    MCAN_REQUIRE(cl != 4 || block_n == 256, "mcan_gemm: cta_group 4 needs block_n 256");
    out->block_n = block_n;
    out->cl = cl;
    out->cg = cg;
    out->mc = mc;
    out->m_tiles = (int)((m + BLOCK_M * cl - 1) / (BLOCK_M * cl));
    out->n_tiles = (int)((n + block_n - 1) / block_n);
    const int slots = sms / cl;
    MCAN_REQUIRE(slots > 0, "mcan_gemm: %d SMs cannot hold a cluster of %d", sms, cl);
    int splits = 1;
    if (accumulate) {
        splits = split_k > 0 ? split_k : pick_splits((int64_t)out->m_tiles * out->n_tiles, kblocks, slots);
        if (splits > kblocks) splits = kblocks;
    }
    out->splits = splits;
    // tail splitting: the last, partial wave as half-width tiles (see GemmParams::full_tiles)
    const int tiles = out->m_tiles * out->n_tiles;
    out->full_tiles = tiles;
    const int rem = tiles % slots;
    if (g_tail_split.load(std::memory_order_relaxed) && splits == 1 && mc == 1 && block_n == 256 &&
        n % 256 == 0 && tiles > slots && rem > 0 && 2 * rem <= slots)
        out->full_tiles = tiles - rem;
    out->units = splits > 1 ? tiles * splits : out->full_tiles + 2 * (tiles - out->full_tiles);
    return 0;
}

}  // namespace mcan

using namespace mcan;

extern "C" int mcan_gemm_plan(int64_t m, int64_t n, int64_t k, int32_t accumulate, int32_t split_k,
                              int32_t block_n, int32_t cta_group, int32_t sms, int32_t* plan_out) {
    MCAN_REQUIRE(plan_out != nullptr && m > 0 && n > 0 && k > 0, "mcan_gemm_plan: bad args");
    if (sms <= 0) sms = device_num_sms();
    MCAN_REQUIRE(sms > 0, "mcan_gemm_plan: no CUDA device and no SM count given");
    GemmPlan pl;
    if (int rc = plan_gemm(m, n, k, accumulate != 0, split_k, block_n, cta_group, sms, &pl)) return rc;
    plan_out[0] = pl.block_n;
    plan_out[1] = pl.cl;
    plan_out[2] = pl.m_tiles;
    plan_out[3] = pl.n_tiles;
    plan_out[4] = pl.splits;
    plan_out[5] = pl.full_tiles;
    plan_out[6] = pl.units;
    return 0;
}

extern "C" int mcan_set_sm_limit(int sms) {
    MCAN_REQUIRE(sms >= 0, "mcan_set_sm_limit: %d", sms);
    mcan::g_sm_limit = sms & ~1;   // keep it even: CTA pairs
    return 0;
}

extern "C" int mcan_set_gemm_schedule(int dynamic) {
    mcan::g_dynamic_schedule.store(dynamic ? 1 : 0, std::memory_order_relaxed);
    return 0;
}

extern "C" int mcan_gemm(const mcan_gemm_args* a) {
    MCAN_REQUIRE(a != nullptr, "mcan_gemm: null args");
    MCAN_REQUIRE(a->num_seg >= 1 && a->num_seg <= MCAN_MAX_GEMM_SEGMENTS, "mcan_gemm: num_seg=%d",
                 a->num_seg);
    MCAN_REQUIRE(a->m > 0 && a->n > 0 && a->k > 0, "mcan_gemm: bad shape m=%lld n=%lld k=%lld",
                 (long long)a->m, (long long)a->n, (long long)a->k);
    MCAN_REQUIRE(a->m < (1LL << 31) && a->n < (1LL << 31) && a->k < (1LL << 31) &&
                     a->m * a->n < (1LL << 32),
                 "mcan_gemm: shape too large");
    MCAN_REQUIRE(a->out_f32 || a->out_bf16, "mcan_gemm: no output");
    MCAN_REQUIRE(!(a->out_bf16_lo && !a->out_bf16), "mcan_gemm: out_bf16_lo needs out_bf16");
    if (a->accumulate) {
        // split-K / accumulate: every epilogue stage must be linear in the accumulator (bias and
        // residual are added by K split 0 only; dropout and the gate scale every partial sum)
        MCAN_REQUIRE(a->out_f32 && !a->out_bf16 && !a->relu,
                     "mcan_gemm: accumulate mode needs out_f32 only and no ReLU");
    } else {
        MCAN_REQUIRE(a->split_k <= 1, "mcan_gemm: split_k needs accumulate");
    }
    MCAN_REQUIRE(a->dropout_p >= 0.f && a->dropout_p < 1.f, "mcan_gemm: dropout_p=%f", a->dropout_p);
    // vector epilogue alignment
    if (a->out_f32) MCAN_REQUIRE(a->ldo_f32 % 4 == 0 && ((uintptr_t)a->out_f32 & 15) == 0, "mcan_gemm: out_f32 alignment");
    if (a->out_bf16) MCAN_REQUIRE(a->ldo_bf16 % 8 == 0 && ((uintptr_t)a->out_bf16 & 15) == 0, "mcan_gemm: out_bf16 alignment");
    if (a->out_bf16_lo) MCAN_REQUIRE(((uintptr_t)a->out_bf16_lo & 15) == 0, "mcan_gemm: out_bf16_lo alignment");
    if (a->resid) MCAN_REQUIRE(a->ldr % 4 == 0 && ((uintptr_t)a->resid & 15) == 0, "mcan_gemm: resid alignment");
    if (a->gate) MCAN_REQUIRE(a->ldg % 8 == 0 && ((uintptr_t)a->gate & 15) == 0, "mcan_gemm: gate alignment");
    if (a->bias) MCAN_REQUIRE(((uintptr_t)a->bias & 15) == 0, "mcan_gemm: bias alignment");
    if (a->colsum)
        MCAN_REQUIRE(((uintptr_t)a->colsum & 15) == 0 && !a->out_f32 && !a->resid && !a->accumulate,
                     "mcan_gemm: colsum needs a 16-byte aligned buffer and a bf16-only output without residual");

    const int sms = device_num_sms();
    MCAN_REQUIRE(sms > 0, "mcan_gemm: no CUDA device");

    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.num_seg = a->num_seg;
    p.m = (int)a->m;
    p.n = (int)a->n;
    p.k = (int)a->k;
    p.kblocks = (int)((a->k + BLOCK_K - 1) / BLOCK_K);
    // Tile override for the fp32 + residual epilogue on short contractions (merge / q-projection dgrad of MCAN-large:
    // 6400 x 1024 x 1024).  That epilogue moves 256 KB per 128 x 256 CTA tile through 8-byte accesses and is as long as
    // the K = 1024 main loop; 100 CTA-pair tiles on 74 pairs leave most of it exposed, 200 single-CTA tiles on 148 SMs
    // overlap it better: 33.5 -> 29.4 us (fp32 + resid), 37.6 -> 32.8 us (+ bias + dropout); for K >= 3072 the pair
    // tile stays ahead (profiles/r02_epilogue_tile_choice.txt).  MCAN_GEMM_RESID_1CTA=0 switches the rule off.
    static const bool resid_1cta = [] { const char* e = getenv("MCAN_GEMM_RESID_1CTA"); return !(e && e[0] == '0'); }();
    int block_n_req = a->block_n, cg_req = a->cta_group;
    if (resid_1cta && block_n_req == 0 && cg_req == 0 && a->out_f32 && a->resid && !a->accumulate && a->num_seg == 1 &&
        a->k <= 1024 && a->n == 1024 && a->m >= 4096) {
        cg_req = 1;
        block_n_req = 256;
    }
    GemmPlan plan;
    if (int rc = plan_gemm(a->m, a->n, a->k, a->accumulate != 0, a->split_k, block_n_req, cg_req, sms, &plan))
        return rc;
    const int block_n = plan.block_n, cl = plan.cl, cg = plan.cg, mc = plan.mc;
    p.m_tiles = plan.m_tiles;
    p.n_tiles = plan.n_tiles;
    p.splits = plan.splits;
    p.full_tiles = plan.full_tiles;
    p.units = plan.units;

    for (int s = 0; s < a->num_seg; ++s) {
        MCAN_REQUIRE(a->a[s] && a->b[s], "mcan_gemm: null operand in segment %d", s);
        int rc;
        if (a->a_layout == 0)
            rc = make_tmap_bf16(&p.tma_a[s], a->a[s], (uint64_t)a->k, (uint64_t)a->m, (uint64_t)a->lda, BLOCK_M);
        else
            rc = make_tmap_bf16(&p.tma_a[s], a->a[s], (uint64_t)a->m, (uint64_t)a->k, (uint64_t)a->lda, 64);
        if (rc) return rc;
        if (a->b_layout == 0) {
            rc = make_tmap_bf16(&p.tma_b[s], a->b[s], (uint64_t)a->k, (uint64_t)a->n, (uint64_t)a->ldb, (uint32_t)(block_n / cl));
            if (rc) return rc;
            if (p.full_tiles < p.m_tiles * p.n_tiles)
                rc = make_tmap_bf16(&p.tma_b_half[s], a->b[s], (uint64_t)a->k, (uint64_t)a->n, (uint64_t)a->ldb, (uint32_t)(block_n / cl / 2));
        } else
            rc = make_tmap_bf16(&p.tma_b[s], a->b[s], (uint64_t)a->n, (uint64_t)a->k, (uint64_t)a->ldb, 64);
        if (rc) return rc;
    }

    p.bias = a->bias;
    p.relu = a->relu;
    p.drop_thr = a->dropout_p > 0.f ? dropout_threshold(a->dropout_p) : 0;
    p.drop_scale = a->dropout_p > 0.f ? 1.0f / (1.0f - a->dropout_p) : 1.0f;
    p.drop_seed = a->dropout_seed;
    p.drop_seed_dev = a->dropout_seed_dev;
    p.gate = reinterpret_cast<const bf16*>(a->gate);
    p.ldg = a->ldg;
    p.gate_scale = a->gate_scale;
    p.resid = a->resid;
    p.ldr = a->ldr;
    p.out_f32 = a->out_f32;
    p.ldo_f32 = a->ldo_f32;
    p.out_bf16 = reinterpret_cast<bf16*>(a->out_bf16);
    p.out_lo = reinterpret_cast<bf16*>(a->out_bf16_lo);
    p.ldo_bf16 = a->ldo_bf16;
    p.accumulate = a->accumulate;
    p.colsum = a->colsum;
    p.tile_counter = nullptr;
    if (g_dynamic_schedule.load(std::memory_order_relaxed)) {
        if (int rc = next_tile_counter(&p.tile_counter)) return rc;
    }
    { const char* d = getenv("MCAN_GEMM_DEBUG"); p.debug = d ? atoi(d) : 0; }

    const int64_t units = p.units;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(a->stream);
    const int am = a->a_layout ? 1 : 0, bm = a->b_layout ? 1 : 0;

    const int epi = (a->out_f32 == nullptr && a->resid == nullptr) ? 1 : 0;
#define MCAN_GEMM_CASE(BN, AM, BM, CG, MC) \
    if (block_n == BN && am == AM && bm == BM && cg == CG && mc == MC) \
        return epi ? launch_gemm<BN, AM, BM, CG, MC, 1>(p, units, sms, st) : launch_gemm<BN, AM, BM, CG, MC, 0>(p, units, sms, st);
#define MCAN_GEMM_CASES(BN, CG, MC) \
    MCAN_GEMM_CASE(BN, 0, 0, CG, MC) MCAN_GEMM_CASE(BN, 0, 1, CG, MC) MCAN_GEMM_CASE(BN, 1, 0, CG, MC) MCAN_GEMM_CASE(BN, 1, 1, CG, MC)
    MCAN_GEMM_CASES(64, 1, 1)
    MCAN_GEMM_CASES(128, 1, 1)
    MCAN_GEMM_CASES(256, 1, 1)
    MCAN_GEMM_CASES(128, 2, 1)
    MCAN_GEMM_CASES(256, 2, 1)
    MCAN_GEMM_CASES(256, 2, 2)
#undef MCAN_GEMM_CASES
#undef MCAN_GEMM_CASE
    set_last_error("mcan_gemm: no kernel for block_n=%d", block_n);
    return -1;
}

// Grouped weight-gradient launch: D_g[M_g,N_g] += A_g^T B_g for up to MCAN_MAX_GEMM_GROUPS problems that share K,
// A_g = bf16 [K, M_g] and B_g = bf16 [K, N_g] read MN-major straight from the activation buffers (dW = dY^T X).
extern "C" int mcan_gemm_grouped(const mcan_gemm_grouped_args* a) {
    MCAN_REQUIRE(a != nullptr, "mcan_gemm_grouped: null args");
    MCAN_REQUIRE(a->num_groups >= 1 && a->num_groups <= MCAN_MAX_GEMM_GROUPS && a->num_groups <= kMaxGroups,
                 "mcan_gemm_grouped: num_groups=%d", a->num_groups);
    MCAN_REQUIRE(a->k > 0 && a->k < (1LL << 31), "mcan_gemm_grouped: k=%lld", (long long)a->k);
    const int sms = device_num_sms();
    MCAN_REQUIRE(sms >= 2, "mcan_gemm_grouped: no CUDA device");
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.num_seg = 1;
    p.num_groups = a->num_groups;
    p.k = (int)a->k;
    p.kblocks = (int)((a->k + BLOCK_K - 1) / BLOCK_K);
    int tiles = 0;
    for (int g = 0; g < a->num_groups; ++g) {
        const mcan_gemm_group& q = a->g[g];
        MCAN_REQUIRE(q.a && q.b && q.out && q.m > 0 && q.n > 0 && q.m < (1LL << 31) && q.n < (1LL << 31) &&
                         q.m * q.n < (1LL << 32),
                     "mcan_gemm_grouped: bad group %d", g);
        MCAN_REQUIRE(q.ldo % (a->out_bf16 ? 8 : 4) == 0 && ((uintptr_t)q.out & 15) == 0, "mcan_gemm_grouped: out alignment (group %d)", g);
        if (int rc = make_tmap_bf16(&p.tma_a[g], q.a, (uint64_t)q.m, (uint64_t)a->k, (uint64_t)q.lda, 64)) return rc;
        if (int rc = make_tmap_bf16(&p.tma_b[g], q.b, (uint64_t)q.n, (uint64_t)a->k, (uint64_t)q.ldb, 64)) return rc;
        p.grp[g].m = (int)q.m;
        p.grp[g].n = (int)q.n;
        p.grp[g].n_tiles = (int)((q.n + 255) / 256);
        p.grp[g].tile_start = tiles;
        p.grp[g].out = q.out;
        p.grp[g].ldo = q.ldo;
        tiles += (int)((q.m + 2 * BLOCK_M - 1) / (2 * BLOCK_M)) * p.grp[g].n_tiles;
    }
    MCAN_REQUIRE(!a->out_bf16 || a->accumulate == 0, "mcan_gemm_grouped: bf16 outputs are written, not accumulated");
    p.group_bf16 = a->out_bf16 ? 1 : 0;
    p.m = p.grp[0].m;
    p.n = p.grp[0].n;
    if (a->out_bf16) {
        p.out_bf16 = reinterpret_cast<bf16*>(p.grp[0].out);
        p.ldo_bf16 = p.grp[0].ldo;
    } else {
        p.out_f32 = reinterpret_cast<float*>(p.grp[0].out);
        p.ldo_f32 = p.grp[0].ldo;
    }
    p.m_tiles = tiles;
    p.n_tiles = 1;
    const int slots = sms / 2;
    // overwrite mode (accumulate == 0): every output element is written by exactly one work unit, plain stores into
    // memory that need not be initialised -- no zero-fill pass and no read-modify-write of the gradient in L2
    MCAN_REQUIRE(a->accumulate != 0 || a->split_k <= 1, "mcan_gemm_grouped: overwrite mode cannot split K");
    int splits = a->accumulate == 0 ? 1 : (a->split_k > 0 ? a->split_k : pick_splits(tiles, p.kblocks, slots));
    if (splits > p.kblocks) splits = p.kblocks;
    p.splits = splits;
    p.full_tiles = tiles;
    p.units = tiles * splits;
    p.drop_scale = 1.0f;
    p.gate_scale = 1.0f;
    p.accumulate = a->accumulate != 0 ? 1 : 0;
    p.tile_counter = nullptr;
    if (g_dynamic_schedule.load(std::memory_order_relaxed)) {
        if (int rc = next_tile_counter(&p.tile_counter)) return rc;
    }
    { const char* d = getenv("MCAN_GEMM_DEBUG"); p.debug = d ? atoi(d) : 0; }
    if (a->out_bf16) return launch_gemm<256, 1, 1, 2, 1, 1>(p, p.units, sms, reinterpret_cast<cudaStream_t>(a->stream));
    return launch_gemm<256, 1, 1, 2, 1, 0>(p, p.units, sms, reinterpret_cast<cudaStream_t>(a->stream));
}


// G1/G2/G3: persistent, warp-specialised tcgen05 GEMM with a fused epilogue.
//
//   D[M,N] = epilogue( sum_s A_s[M,K] * B_s[N,K]^T ),  bf16 operands, fp32 accumulate in TMEM.
//
// It is the one GEMM behind every nn.Linear of the MCAN hot path -- forward
// (mca.py:33,40,47,61; net_utils.py:26,45; net.py:39,53), dgrad (B read MN-major straight
// from the (out,in) weight) and wgrad (both operands read MN-major straight from the
// activations, split-K with fp32 atomics).  No transposed copies are ever made.
//
// CTA = 192 threads:  warp 0 = TMA producer, warp 1 = TMEM owner + single-thread MMA issuer,
// warps 2..5 = epilogue (one TMEM lane quarter each).  Pipelines: smem ring full/empty
// (TMA <-> MMA) and a double-buffered TMEM accumulator full/empty (MMA <-> epilogue), so the
// epilogue of tile i overlaps the main loop of tile i+1.  Grid = min(#work units, #SMs);
// work unit = (output tile, K split), statically strided over the CTAs.
//
// Tile 128 x BLOCK_N (128 | 256) x 64.  Operand tiles are TMA boxes with the 128-byte swizzle:
//   K-major  tile [rows x 64 k]  : one box {64, rows}; UMMA desc SBO = 1024 B, k-step = +32 B
//   MN-major tile [64 k x rows]  : rows/64 boxes {64 mn, 64 k} of 8 KiB; UMMA desc
//                                  LBO = 8 KiB (next 64 mn), SBO = 1024 B (next 8 k), k-step = +2 KiB
#include <mutex>
#include <string.h>
#include <unordered_map>

#include "../../include/mcan_b200.h"
#include "common.cuh"

namespace mcan {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int kGemmThreads = 192;
constexpr int kEpilogueWarps = 4;

struct alignas(64) GemmParams {
    CUtensorMap tma_a[MCAN_MAX_GEMM_SEGMENTS];
    CUtensorMap tma_b[MCAN_MAX_GEMM_SEGMENTS];
    int num_seg;
    int m, n, k;
    int m_tiles, n_tiles, splits, kblocks;
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
};

template <int BLOCK_N>
struct GemmCfg {
    static constexpr int kStages = (BLOCK_N == 256) ? 4 : 6;
    static constexpr uint32_t kABytes = BLOCK_M * BLOCK_K * 2;
    static constexpr uint32_t kBBytes = BLOCK_N * BLOCK_K * 2;
    static constexpr uint32_t kStageBytes = kABytes + kBBytes;
    static constexpr uint32_t kTmemCols = 2 * BLOCK_N;
    static constexpr uint32_t kBarrierBytes = (2 * kStages + 4) * 8 + 16;
    static constexpr uint32_t kEpiBytes = kEpilogueWarps * 32 * 68 * 4;   // per-warp transpose tiles
    static constexpr uint32_t kSmemBytes = kStages * kStageBytes + kEpiBytes + kBarrierBytes;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b),
                 "f"(c), "f"(d)
                 : "memory");
}

__device__ __forceinline__ void red_add_v2(float* addr, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

// Epilogue of one 32-row x 64-column chunk, executed by one warp AFTER the chunk has been
// transposed through shared memory: lane l owns columns col, col+1 (col = nc + 2*l) of every
// row, so each global load/store instruction of the warp covers one contiguous 128/256-byte
// row segment (fully coalesced), the bias is loaded once per chunk and one dropout hash serves
// an element pair.  `stage` = this warp's [32][kStageLd] fp32 staging tile.
constexpr int kStageLd = 68;   // floats; 16-byte aligned rows, conflict-free v4 writes / v2 reads
constexpr int kChunkN = 64;

__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, const float* stage, int lane,
                                               long long row0, int rows_valid, int nc,
                                               uint32_t drop_seed) {
    const int col = nc + 2 * lane;
    const bool c0 = col < p.n, c1 = col + 1 < p.n;
    if (!c0) return;
    const bool pair = c1;   // both columns valid -> vector accesses (col is even)
    float b0 = 0.f, b1 = 0.f;
    if (p.bias != nullptr) {
        b0 = __ldg(p.bias + col);
        if (c1) b1 = __ldg(p.bias + col + 1);
    }
    const bool drop_pair = ((p.n & 1) == 0);
#pragma unroll 1
    for (int r0 = 0; r0 < rows_valid; r0 += 8) {
        float2 res[8];
        uint32_t gt[8];
        if (p.resid != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                res[j] = make_float2(0.f, 0.f);
                if (r0 + j < rows_valid) {
                    const float* rp = p.resid + (row0 + r0 + j) * p.ldr + col;
                    if (pair) res[j] = *reinterpret_cast<const float2*>(rp);
                    else res[j].x = *rp;
                }
            }
        }
        if (p.gate != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                gt[j] = 0;
                if (r0 + j < rows_valid) {
                    const bf16* gp = p.gate + (row0 + r0 + j) * p.ldg + col;
                    if (pair) gt[j] = __ldg(reinterpret_cast<const uint32_t*>(gp));
                    else gt[j] = (uint32_t)(*reinterpret_cast<const unsigned short*>(gp));
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (r0 + j >= rows_valid) break;
            const long long row = row0 + r0 + j;
            const float2 a = *reinterpret_cast<const float2*>(stage + (r0 + j) * kStageLd + 2 * lane);
            float v0 = a.x + b0, v1 = a.y + b1;
            if (p.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
            if (p.drop_thr != 0) {
                const uint32_t idx = (uint32_t)(row * (long long)p.n + col);
                uint32_t u0, u1;
                if (drop_pair) {
                    const uint32_t rnd = dropout_bits_pair(idx >> 1, drop_seed);
                    u0 = rnd & 0xFFFFU; u1 = rnd >> 16;
                } else {
                    u0 = dropout_u16(idx, drop_seed); u1 = dropout_u16(idx + 1, drop_seed);
                }
                v0 = (u0 >= p.drop_thr) ? v0 * p.drop_scale : 0.f;
                v1 = (u1 >= p.drop_thr) ? v1 * p.drop_scale : 0.f;
            }
            if (p.gate != nullptr) {
                v0 = (bf16_lo_to_f(gt[j]) > 0.f) ? v0 * p.gate_scale : 0.f;
                v1 = (bf16_hi_to_f(gt[j]) > 0.f) ? v1 * p.gate_scale : 0.f;
            }
            if (p.resid != nullptr) { v0 += res[j].x; v1 += res[j].y; }
            if (p.out_f32 != nullptr) {
                float* o = p.out_f32 + row * p.ldo_f32 + col;
                if (p.accumulate) {
                    if (pair) red_add_v2(o, v0, v1);
                    else atomicAdd(o, v0);
                } else if (pair) {
                    *reinterpret_cast<float2*>(o) = make_float2(v0, v1);
                } else {
                    *o = v0;
                }
            }
            if (p.out_bf16 != nullptr) {
                bf16* o = p.out_bf16 + row * p.ldo_bf16 + col;
                const uint32_t packed = pack_bf16x2(v0, v1);
                if (pair) *reinterpret_cast<uint32_t*>(o) = packed;
                else *o = __float2bfloat16_rn(v0);
                if (p.out_lo != nullptr) {
                    bf16* ol = p.out_lo + row * p.ldo_bf16 + col;
                    const float l0 = v0 - bf16_lo_to_f(packed), l1 = v1 - bf16_hi_to_f(packed);
                    if (pair) *reinterpret_cast<uint32_t*>(ol) = pack_bf16x2(l0, l1);
                    else *ol = __float2bfloat16_rn(l0);
                }
            }
        }
    }
}

template <int BLOCK_N, int A_MN, int B_MN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ GemmParams p) {
    using Cfg = GemmCfg<BLOCK_N>;
    constexpr int kStages = Cfg::kStages;

    // SWIZZLE_128B tiles need 1024-byte alignment; the kernel has no static shared memory, so the
    // dynamic window starts at the (1024-aligned) base of the CTA's shared memory.
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if ((smem_u32(smem) & 1023U) != 0) {
        if (threadIdx.x == 0) printf("mcan gemm: dynamic smem base not 1024-byte aligned\n");
        __trap();
    }
    float* epi_stage = reinterpret_cast<float*>(smem + kStages * Cfg::kStageBytes);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes + Cfg::kEpiBytes);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full_bar = empty_bar + kStages;
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < p.num_seg; ++s) {
            prefetch_tmap(&p.tma_a[s]);
            prefetch_tmap(&p.tma_b[s]);
        }
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < kStages; ++s) {
                mbar_init(&full_bar[s], 1);
                mbar_init(&empty_bar[s], 1);
            }
            for (int s = 0; s < 2; ++s) {
                mbar_init(&tmem_full_bar[s], 1);
                mbar_init(&tmem_empty_bar[s], kEpilogueWarps);
            }
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, Cfg::kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int tiles = p.m_tiles * p.n_tiles;
    const int units = tiles * p.splits;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
                const int tile = unit % tiles, split = unit / tiles;
                const int m0 = (tile / p.n_tiles) * BLOCK_M;
                const int n0 = (tile % p.n_tiles) * BLOCK_N;
                const int kb0 = (int)((long long)p.kblocks * split / p.splits);
                const int kb1 = (int)((long long)p.kblocks * (split + 1) / p.splits);
                for (int seg = 0; seg < p.num_seg; ++seg) {
                    for (int kb = kb0; kb < kb1; ++kb) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        uint8_t* sa = smem + stage * Cfg::kStageBytes;
                        uint8_t* sb = sa + Cfg::kABytes;
                        mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
                        if (A_MN) {
#pragma unroll
                            for (int c = 0; c < BLOCK_M / 64; ++c)
                                tma_load_2d(sa + c * (BLOCK_K * 128), &p.tma_a[seg],
                                            &full_bar[stage], m0 + c * 64, kb * BLOCK_K);
                        } else {
                            tma_load_2d(sa, &p.tma_a[seg], &full_bar[stage], kb * BLOCK_K, m0);
                        }
                        if (B_MN) {
#pragma unroll
                            for (int c = 0; c < BLOCK_N / 64; ++c)
                                tma_load_2d(sb + c * (BLOCK_K * 128), &p.tma_b[seg],
                                            &full_bar[stage], n0 + c * 64, kb * BLOCK_K);
                        } else {
                            tma_load_2d(sb, &p.tma_b[seg], &full_bar[stage], kb * BLOCK_K, n0);
                        }
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BLOCK_N, A_MN, B_MN);
            constexpr uint32_t a_lbo = A_MN ? BLOCK_K * 128 : 0;
            constexpr uint32_t b_lbo = B_MN ? BLOCK_K * 128 : 0;
            constexpr uint32_t a_kstep = A_MN ? (UMMA_K * 128) : (UMMA_K * 2);
            constexpr uint32_t b_kstep = B_MN ? (UMMA_K * 128) : (UMMA_K * 2);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
                const int split = unit / tiles;
                const int kb0 = (int)((long long)p.kblocks * split / p.splits);
                const int kb1 = (int)((long long)p.kblocks * (split + 1) / p.splits);
                const int iters = (kb1 - kb0) * p.num_seg;
                mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
                for (int it = 0; it < iters; ++it) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
                    const uint32_t sb = sa + Cfg::kABytes;
                    const uint64_t adesc = make_smem_desc_sw128(sa, a_lbo, 1024);
                    const uint64_t bdesc = make_smem_desc_sw128(sb, b_lbo, 1024);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        umma_bf16(tmem_d, adesc + (uint64_t)((k * a_kstep) >> 4),
                                  bdesc + (uint64_t)((k * b_kstep) >> 4), idesc,
                                  (it > 0 || k > 0) ? 1U : 0U);
                    }
                    umma_commit(&empty_bar[stage]);  // frees the smem slot when the MMAs retire
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tmem_full_bar[acc]);  // accumulator ready for the epilogue
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ========== epilogue: TMEM -> registers -> smem transpose -> coalesced global ==========
        const int quad = warp & 3;  // TMEM lanes [32*quad, 32*quad+32) belong to this warp
        const uint32_t drop_seed =
            p.drop_seed ^ ((p.drop_thr != 0 && p.drop_seed_dev != nullptr) ? __ldg(p.drop_seed_dev) : 0U);
        float* stage = epi_stage + (warp - 2) * (32 * kStageLd);
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
            const int tile = unit % tiles;
            const int m0 = (tile / p.n_tiles) * BLOCK_M;
            const int n0 = (tile % p.n_tiles) * BLOCK_N;
            const long long row0 = m0 + quad * 32;
            const int rows_valid = (int)max(0LL, min(32LL, (long long)p.m - row0));
            mbar_wait(&tmem_full_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BLOCK_N);
            const int nchunks = min(BLOCK_N / kChunkN, (p.n - n0 + kChunkN - 1) / kChunkN);
#pragma unroll 1
            for (int c = 0; c < nchunks; ++c) {
                uint32_t r[32];
                float4* srow = reinterpret_cast<float4*>(stage + lane * kStageLd);
#pragma unroll
                for (int hlf = 0; hlf < 2; ++hlf) {
                    tmem_ld_32x32(taddr + (uint32_t)(c * kChunkN + hlf * 32), r);
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        srow[hlf * 8 + q] = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                                                        __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
                }
                if (c == nchunks - 1) {
                    // the accumulator stage is fully drained: hand it back to the MMA warp now
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
                } else {
                    __syncwarp();
                }
                if (rows_valid > 0)
                    epilogue_chunk(p, stage, lane, row0, rows_valid, n0 + c * kChunkN, drop_seed);
                __syncwarp();   // staging tile is reused by the next chunk
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::kTmemCols);
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
static int make_tmap(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld,
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

int device_num_sms() {
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

template <int BLOCK_N, int A_MN, int B_MN>
static int launch_gemm(const GemmParams& p, int grid, cudaStream_t stream) {
    using Cfg = GemmCfg<BLOCK_N>;
    static bool configured[64] = {false};
    int dev = 0;
    MCAN_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        MCAN_CHECK_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<BLOCK_N, A_MN, B_MN>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)Cfg::kSmemBytes));
        configured[dev] = true;
    }
    gemm_tcgen05_kernel<BLOCK_N, A_MN, B_MN><<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(p);
    MCAN_CHECK_CUDA(cudaGetLastError());
    return 0;
}

static int pick_block_n(int64_t m_tiles, int64_t n, int sms) {
    if (n <= 128) return 128;
    // cost model: waves x tile time; a 128-wide tile is smem-bandwidth bound (~15 % slower per flop)
    const int64_t t256 = m_tiles * ((n + 255) / 256);
    const int64_t t128 = m_tiles * ((n + 127) / 128);
    const double c256 = (double)((t256 + sms - 1) / sms) * 2.0;
    const double c128 = (double)((t128 + sms - 1) / sms) * 1.15;
    return (c128 < c256) ? 128 : 256;
}

static int pick_splits(int64_t tiles, int kblocks, int sms) {
    int best = 1;
    double best_cost = 1e30;
    const int max_s = kblocks < 32 ? kblocks : 32;
    for (int s = 1; s <= max_s; ++s) {
        const int64_t units = tiles * s;
        const int64_t waves = (units + sms - 1) / sms;
        const double per_unit = (double)((kblocks + s - 1) / s) + 6.0;  // +epilogue/fill overhead
        const double cost = (double)waves * per_unit;
        if (cost < best_cost * 0.98) {
            best_cost = cost;
            best = s;
        }
    }
    return best;
}

}  // namespace mcan

using namespace mcan;

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
        MCAN_REQUIRE(a->out_f32 && !a->out_bf16 && !a->bias && !a->relu && a->dropout_p == 0.f &&
                         !a->gate && !a->resid,
                     "mcan_gemm: accumulate mode supports only out_f32");
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

    const int sms = device_num_sms();
    MCAN_REQUIRE(sms > 0, "mcan_gemm: no CUDA device");

    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.num_seg = a->num_seg;
    p.m = (int)a->m;
    p.n = (int)a->n;
    p.k = (int)a->k;
    p.m_tiles = (int)((a->m + BLOCK_M - 1) / BLOCK_M);
    p.kblocks = (int)((a->k + BLOCK_K - 1) / BLOCK_K);
    int block_n = a->block_n ? a->block_n : pick_block_n(p.m_tiles, a->n, sms);
    MCAN_REQUIRE(block_n == 128 || block_n == 256, "mcan_gemm: block_n=%d", block_n);
    p.n_tiles = (int)((a->n + block_n - 1) / block_n);
    int splits = 1;
    if (a->accumulate) {
        splits = a->split_k > 0 ? a->split_k : pick_splits((int64_t)p.m_tiles * p.n_tiles, p.kblocks, sms);
        if (splits > p.kblocks) splits = p.kblocks;
    }
    p.splits = splits;

    for (int s = 0; s < a->num_seg; ++s) {
        MCAN_REQUIRE(a->a[s] && a->b[s], "mcan_gemm: null operand in segment %d", s);
        int rc;
        if (a->a_layout == 0)
            rc = make_tmap(&p.tma_a[s], a->a[s], (uint64_t)a->k, (uint64_t)a->m, (uint64_t)a->lda, BLOCK_M);
        else
            rc = make_tmap(&p.tma_a[s], a->a[s], (uint64_t)a->m, (uint64_t)a->k, (uint64_t)a->lda, 64);
        if (rc) return rc;
        if (a->b_layout == 0)
            rc = make_tmap(&p.tma_b[s], a->b[s], (uint64_t)a->k, (uint64_t)a->n, (uint64_t)a->ldb, (uint32_t)block_n);
        else
            rc = make_tmap(&p.tma_b[s], a->b[s], (uint64_t)a->n, (uint64_t)a->k, (uint64_t)a->ldb, 64);
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

    const int64_t units = (int64_t)p.m_tiles * p.n_tiles * p.splits;
    const int grid = (int)(units < sms ? units : sms);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(a->stream);
    const int am = a->a_layout ? 1 : 0, bm = a->b_layout ? 1 : 0;

#define MCAN_GEMM_CASE(BN, AM, BM) \
    if (block_n == BN && am == AM && bm == BM) return launch_gemm<BN, AM, BM>(p, grid, st);
    MCAN_GEMM_CASE(128, 0, 0)
    MCAN_GEMM_CASE(128, 0, 1)
    MCAN_GEMM_CASE(128, 1, 0)
    MCAN_GEMM_CASE(128, 1, 1)
    MCAN_GEMM_CASE(256, 0, 0)
    MCAN_GEMM_CASE(256, 0, 1)
    MCAN_GEMM_CASE(256, 1, 0)
    MCAN_GEMM_CASE(256, 1, 1)
#undef MCAN_GEMM_CASE
    set_last_error("mcan_gemm: no kernel for block_n=%d", block_n);
    return -1;
}

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
    static constexpr uint32_t kSmemBytes = kStages * kStageBytes + kBarrierBytes + 1024;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b),
                 "f"(c), "f"(d)
                 : "memory");
}

// Applies the fused epilogue to 8 consecutive columns of one output row and stores them.
// `full` = all 8 columns are inside N (vector path), otherwise per-element guards.
__device__ __forceinline__ void epilogue_store8(const GemmParams& p, float (&v)[8], long long row,
                                                int n0, bool full, uint32_t drop_seed) {
    const int nvalid = full ? 8 : max(0, min(8, p.n - n0));
    if (nvalid == 0) return;
    if (p.bias != nullptr) {
        if (full) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n0));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + 4));
            v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
            v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < nvalid) v[j] += __ldg(p.bias + n0 + j);
        }
    }
    if (p.relu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    if (p.drop_thr != 0) {
        const uint32_t base = (uint32_t)(row * (long long)p.n + n0);
        if ((base & 1U) == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t r = dropout_bits_pair((base >> 1) + j, drop_seed);
                v[2 * j] = ((r & 0xFFFFU) >= p.drop_thr) ? v[2 * j] * p.drop_scale : 0.f;
                v[2 * j + 1] = ((r >> 16) >= p.drop_thr) ? v[2 * j + 1] * p.drop_scale : 0.f;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t u = dropout_u16(base + j, drop_seed);
                v[j] = (u >= p.drop_thr) ? v[j] * p.drop_scale : 0.f;
            }
        }
    }
    if (p.gate != nullptr) {
        const bf16* g = p.gate + row * p.ldg + n0;
        if (full) {
            const uint4 gv = __ldg(reinterpret_cast<const uint4*>(g));
            const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                v[2 * j] = (bf16_lo_to_f(gw[j]) > 0.f) ? v[2 * j] * p.gate_scale : 0.f;
                v[2 * j + 1] = (bf16_hi_to_f(gw[j]) > 0.f) ? v[2 * j + 1] * p.gate_scale : 0.f;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < nvalid)
                    v[j] = (__bfloat162float(g[j]) > 0.f) ? v[j] * p.gate_scale : 0.f;
        }
    }
    if (p.resid != nullptr) {
        const float* r = p.resid + row * p.ldr + n0;
        if (full) {
            const float4 r0 = *reinterpret_cast<const float4*>(r);
            const float4 r1 = *reinterpret_cast<const float4*>(r + 4);
            v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
            v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < nvalid) v[j] += r[j];
        }
    }
    if (p.out_f32 != nullptr) {
        float* o = p.out_f32 + row * p.ldo_f32 + n0;
        if (p.accumulate) {
            if (full) {
                red_add_v4(o, v[0], v[1], v[2], v[3]);
                red_add_v4(o + 4, v[4], v[5], v[6], v[7]);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (j < nvalid) atomicAdd(o + j, v[j]);
            }
        } else if (full) {
            *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < nvalid) o[j] = v[j];
        }
    }
    if (p.out_bf16 != nullptr) {
        bf16* o = p.out_bf16 + row * p.ldo_bf16 + n0;
        if (full) {
            uint4 w;
            w.x = pack_bf16x2(v[0], v[1]);
            w.y = pack_bf16x2(v[2], v[3]);
            w.z = pack_bf16x2(v[4], v[5]);
            w.w = pack_bf16x2(v[6], v[7]);
            *reinterpret_cast<uint4*>(o) = w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < nvalid) o[j] = __float2bfloat16_rn(v[j]);
        }
    }
    if (p.out_lo != nullptr) {
        bf16* o = p.out_lo + row * p.ldo_bf16 + n0;
        float l[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) l[j] = v[j] - __bfloat162float(__float2bfloat16_rn(v[j]));
        if (full) {
            uint4 w;
            w.x = pack_bf16x2(l[0], l[1]);
            w.y = pack_bf16x2(l[2], l[3]);
            w.z = pack_bf16x2(l[4], l[5]);
            w.w = pack_bf16x2(l[6], l[7]);
            *reinterpret_cast<uint4*>(o) = w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < nvalid) o[j] = __float2bfloat16_rn(l[j]);
        }
    }
}

template <int BLOCK_N, int A_MN, int B_MN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ GemmParams p) {
    using Cfg = GemmCfg<BLOCK_N>;
    constexpr int kStages = Cfg::kStages;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
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
        // ===================== epilogue: TMEM -> registers -> global =====================
        const int quad = warp & 3;  // TMEM lanes [32*quad, 32*quad+32) belong to this warp
        const uint32_t drop_seed =
            p.drop_seed ^ ((p.drop_thr != 0 && p.drop_seed_dev != nullptr) ? __ldg(p.drop_seed_dev) : 0U);
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
            const int tile = unit % tiles;
            const int m0 = (tile / p.n_tiles) * BLOCK_M;
            const int n0 = (tile % p.n_tiles) * BLOCK_N;
            const long long row = m0 + quad * 32 + lane;
            const bool row_ok = row < p.m;
            mbar_wait(&tmem_full_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BLOCK_N);
#pragma unroll 1
            for (int c = 0; c < BLOCK_N / 32; ++c) {
                const int nc = n0 + c * 32;
                if (nc >= p.n) break;  // warp-uniform
                uint32_t r[32];
                tmem_ld_32x32(taddr + (uint32_t)(c * 32), r);
                tmem_ld_wait();
                if (row_ok) {
                    const bool full = (nc + 32 <= p.n);
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        float v[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[g * 8 + j]);
                        epilogue_store8(p, v, row, nc + g * 8, full, drop_seed);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
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

// Shared device helpers for the MCAN co-attention kernels (sm_100a only).
//
// Everything here is a thin wrapper around one PTX instruction (mbarrier, TMA,
// tcgen05/TMEM, ldmatrix/mma.sync) plus the counter-based dropout hash that the
// forward and backward kernels must agree on.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace mcan {

typedef __nv_bfloat16 bf16;

constexpr int kWarp = 32;

// ---------------------------------------------------------------------------
// error plumbing shared by all translation units (defined in c_api.cu)
// ---------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);

#define MCAN_CHECK_CUDA(expr)                                                          \
    do {                                                                               \
        cudaError_t _e = (expr);                                                       \
        if (_e != cudaSuccess) {                                                       \
            ::mcan::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,       \
                                   cudaGetErrorString(_e));                            \
            return -2;                                                                 \
        }                                                                              \
    } while (0)

#define MCAN_REQUIRE(cond, ...)                                                        \
    do {                                                                               \
        if (!(cond)) {                                                                 \
            ::mcan::set_last_error(__VA_ARGS__);                                       \
            return -1;                                                                 \
        }                                                                              \
    } while (0)

// ---------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  Every kernel of the library
//   * calls pdl_launch_dependents() first: once all CTAs of a grid have done so, the NEXT kernel
//     of the stream may start being scheduled onto SMs as they drain -- its launch latency and
//     prologue (barrier init, TMEM allocation, descriptor prefetch) hide behind this grid's tail;
//   * calls pdl_wait() before its first access to global memory: it returns when all
//     prerequisite grids have COMPLETED and their writes are visible, so no data hazard is
//     introduced (completion is transitive because every kernel waits before it finishes).
// Both are no-ops for a kernel launched without the attribute (mcan_set_pdl(0)).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void pdl_launch_dependents() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool pdl_enabled();   // c_api.cu

// <<<>>> replacement: launches `kernel` with the PDL attribute when enabled.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                 cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------------------
// dropout: counter-based hash, 16 random bits per element.
//   keep(idx) <=> u16(idx) >= thr,   thr = round(p * 65536)
// Forward epilogues and backward kernels regenerate the same mask from
// (seed, linear element index); no mask tensor is ever stored.
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7feb352dU;
    x ^= x >> 15;
    x *= 0x846ca68bU;
    x ^= x >> 16;
    return x;
}

__host__ __device__ __forceinline__ uint32_t dropout_bits_pair(uint32_t pair_idx, uint32_t seed) {
    return mix32(pair_idx * 0x9E3779B9U + seed);
}

// random 16 bits of element `idx`
__host__ __device__ __forceinline__ uint32_t dropout_u16(uint32_t idx, uint32_t seed) {
    uint32_t r = dropout_bits_pair(idx >> 1, seed);
    return (idx & 1U) ? (r >> 16) : (r & 0xFFFFU);
}

__host__ __device__ __forceinline__ uint32_t dropout_threshold(float p) {
    float t = p * 65536.0f + 0.5f;
    if (t < 0.f) t = 0.f;
    if (t > 65535.f) t = 65535.f;
    return (uint32_t)t;
}

// ---------------------------------------------------------------------------
// small PTX wrappers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n"
        ".reg .b32 %%rx;\n"
        ".reg .pred %%px;\n"
        "     elect.sync %%rx|%%px, %1;\n"
        "@%%px mov.s32 %0, 1;\n"
        "}\n"
        : "+r"(pred)
        : "r"(0xFFFFFFFFU));
    return pred;
}

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must trap, never hang the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    long long t0 = clock64();
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 0x3FFU) == 0 && (clock64() - t0) > 4000000000LL) {
            printf("mcan: mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x,
                   (int)threadIdx.x);
            __trap();
        }
    }
}

// ---- TMA ----
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tiled load: c0 = coordinate in the contiguous dimension, c1 = row.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// ---- tcgen05 / TMEM ----
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(slot_in_smem)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> fp32
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all previously issued MMAs of this thread arrive on `bar` when they retire
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::
                     "r"(smem_u32(bar))
                 : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
          "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
// 16 lanes x (8 repetitions of 256 bit): thread (g = lane/4, t = lane%4) receives
// r[4k + 2h + c] = D[lane_base + g + 8h][col_base + 8k + 2t + c]   (k < 8, h < 2, c < 2)
__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
          "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- cta_group::2 (CTA pair) variants ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.aligned;\nbarrier.cluster.wait.aligned;" ::: "memory");
}
// In a CTA pair the shared::cluster address of the even (leader) CTA is the local address with
// bit 24 cleared (same convention as CUTLASS' Sm100MmaPeerBitMask).
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFU;
// TMA load whose completion bytes are signalled on the LEADER CTA's mbarrier.
__device__ __forceinline__ void tma_load_2d_cg2(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                                int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
        : "memory");
}
// Same, multicast: the box lands at the same CTA-relative offset in every CTA of `cta_mask`, and
// each destination's bytes are signalled on the mbarrier of the leader of ITS pair.
__device__ __forceinline__ void tma_load_2d_cg2_mc(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                                   int32_t c0, int32_t c1, uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        ".multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;"
        ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "h"(cta_mask), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t* slot_in_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(slot_in_smem)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish_cg2() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
}
// D[tmem of both CTAs] (+)= A * B with M = 256 split over the CTA pair; issued by the leader only
__device__ __forceinline__ void umma_bf16_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                              uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// commit: arrive on the mbarrier at the same offset in every CTA of `cta_mask`
__device__ __forceinline__ void umma_commit_cg2(uint64_t* bar, uint16_t cta_mask) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
        "[%0], %1;" ::"r"(smem_u32(bar)),
        "h"(cta_mask)
        : "memory");
}
// arrive on the mbarrier at the same offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n"
        ".reg .b32 ra;\n"
        "mapa.shared::cluster.u32 ra, %0, %1;\n"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(cta)
        : "memory");
}

// cluster-scope release / acquire variants: used where DATA written to a peer CTA's shared memory
// must be visible to the thread that wakes up on the barrier (dynamic tile scheduler)
__device__ __forceinline__ void mbar_arrive_release_cluster(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n"
        ".reg .b32 ra;\n"
        "mapa.shared::cluster.u32 ra, %0, %1;\n"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(cta)
        : "memory");
}
__device__ __forceinline__ void st_shared_remote_u32(uint32_t* local_addr, uint32_t cta, uint32_t val) {
    asm volatile(
        "{\n"
        ".reg .b32 ra;\n"
        "mapa.shared::cluster.u32 ra, %0, %1;\n"
        "st.shared::cluster.u32 [ra], %2;\n"
        "}\n" ::"r"(smem_u32(local_addr)),
        "r"(cta), "r"(val)
        : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_acq_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_acq_cluster(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait_acq_cluster(bar, parity)) return;
    long long t0 = clock64();
    uint32_t spins = 0;
    while (!mbar_try_wait_acq_cluster(bar, parity)) {
        if ((++spins & 0x3FFU) == 0 && (clock64() - t0) > 4000000000LL) {
            printf("mcan: scheduler mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x,
                   (int)threadIdx.x);
            __trap();
        }
    }
}

// ---- UMMA descriptors (sm_100 "version 1" shared-memory matrix descriptor) ----
// layout_type 2 = SWIZZLE_128B. Offsets are in bytes here, encoded >> 4.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes,
                                                         uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFU) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFU) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFU) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;  // SWIZZLE_128B
    return d;
}
__host__ __device__ constexpr uint64_t make_smem_desc_sw128_const(uint32_t smem_addr, uint32_t lbo_bytes,
                                                                  uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr & 0x3FFFFU) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFU) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFU) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// instruction descriptor, kind::f16: bf16 A/B, fp32 accumulate
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major,
                                                       int b_mn_major) {
    return (1U << 4) | (1U << 7) | (1U << 10) | ((uint32_t)a_mn_major << 15) |
           ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- misc ----
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffU, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffU, v, o));
    return v;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo_to_f(uint32_t packed) {
    return __uint_as_float(packed << 16);
}
__device__ __forceinline__ float bf16_hi_to_f(uint32_t packed) {
    return __uint_as_float(packed & 0xFFFF0000U);
}

// ---- accumulator fragment layouts (tcgen05.ld 16x256b.x8: [16 rows x 64 columns] per warp) ----
// direct: thread (g = lane/4, t = lane%4) holds r[4k + 2h + c] = D[g + 8h][8k + 2t + c]  (k < 8, h < 2, c < 2)
// T8    : after quad_transpose, v[16q + 4p + 2h + c] = D[g + 8h][32q + 8t + 2p + c]: eight consecutive columns
//         per (q, h), so that every global access of an epilogue is a 16-byte vector
// 4 x 4 transpose inside every quad of lanes, on each of the 8 groups {q, h, c} of four registers
// v[16q + 4kk + 2h + c] (kk < 4): afterwards register kk holds what lane (t & ~3) + kk held in register t,
// i.e. v[16q + 4p + 2h + c] = D[row0 + g + 8h][col0 + 32q + 8t + 2p + c].
__device__ __forceinline__ void quad_transpose(float (&v)[32], int t) {
    const bool odd = (t & 1) != 0, hi = (t & 2) != 0;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
#pragma unroll
        for (int hc = 0; hc < 4; ++hc) {
            const int b = 16 * q + hc;
#pragma unroll
            for (int i = 0; i < 4; i += 2) {          // exchange with lane ^ 1
                const float x = v[b + 4 * i], y = v[b + 4 * (i + 1)];
                const float recv = __shfl_xor_sync(0xffffffffU, odd ? x : y, 1);
                v[b + 4 * i] = odd ? recv : x;
                v[b + 4 * (i + 1)] = odd ? y : recv;
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {             // exchange with lane ^ 2
                const float x = v[b + 4 * i], y = v[b + 4 * (i + 2)];
                const float recv = __shfl_xor_sync(0xffffffffU, hi ? x : y, 2);
                v[b + 4 * i] = hi ? recv : x;
                v[b + 4 * (i + 2)] = hi ? y : recv;
            }
        }
    }
}
// element e (< 8) of column group q, row half h in the T8 layout
#define T8(v, q, h, e) v[16 * (q) + 4 * ((e) >> 1) + 2 * (h) + ((e) & 1)]


}  // namespace mcan

// Small memory-bound helpers around the GEMMs: fp32 -> bf16 (hi [+ lo]) casts for GEMM
// operands, and column sums (bias gradients = colsum of the GEMM output gradient).
#include "../../include/mcan_b200.h"
#include "common.cuh"

namespace mcan {

int device_num_sms();

__global__ void __launch_bounds__(256)
cast_bf16_kernel(const float* __restrict__ x, long long n, bf16* __restrict__ hi,
                 bf16* __restrict__ lo) {
    pdl_launch_dependents();
    pdl_wait();
    const long long nv = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
        const float4 v = reinterpret_cast<const float4*>(x)[i];
        uint2 w;
        w.x = pack_bf16x2(v.x, v.y);
        w.y = pack_bf16x2(v.z, v.w);
        reinterpret_cast<uint2*>(hi)[i] = w;
        if (lo != nullptr) {
            uint2 l;
            l.x = pack_bf16x2(v.x - bf16_lo_to_f(w.x), v.y - bf16_hi_to_f(w.x));
            l.y = pack_bf16x2(v.z - bf16_lo_to_f(w.y), v.w - bf16_hi_to_f(w.y));
            reinterpret_cast<uint2*>(lo)[i] = l;
        }
    }
    // tail (n % 4 elements)
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const long long i = (nv << 2) + threadIdx.x;
        const bf16 h = __float2bfloat16_rn(x[i]);
        hi[i] = h;
        if (lo != nullptr) lo[i] = __float2bfloat16_rn(x[i] - __bfloat162float(h));
    }
}

// Multi-tensor cast/copy: one launch refreshes the bf16 GEMM-operand copies of MANY fp32 master
// weights (and concatenates the fp32 biases of stacked layers).  The segment table lives in
// device memory so the launch is CUDA-graph friendly.
struct CastSeg {
    const float* src;
    void* dst;
    long long n;            // elements
    long long first_chunk;  // index of this segment's first 4096-element chunk; bit 62 set = fp32 destination
    void* dst_lo;           // optional bf16(x - bf16(x)) destination (split precision)
};
constexpr long long kCastChunk = 4096;
constexpr long long kCastF32Flag = 1LL << 62;

__global__ void __launch_bounds__(256)
cast_multi_kernel(const CastSeg* __restrict__ segs, int nseg, long long total_chunks) {
    pdl_launch_dependents();
    pdl_wait();
    for (long long chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
        int lo = 0, hi = nseg - 1;
        while (lo < hi) {   // last segment whose first_chunk <= chunk
            const int mid = (lo + hi + 1) >> 1;
            if ((segs[mid].first_chunk & ~kCastF32Flag) <= chunk) lo = mid; else hi = mid - 1;
        }
        const CastSeg sg = segs[lo];
        const bool to_f32 = (sg.first_chunk & kCastF32Flag) != 0;
        const long long off = (chunk - (sg.first_chunk & ~kCastF32Flag)) * kCastChunk;
        const long long n = min(kCastChunk, sg.n - off);
        const float* src = sg.src + off;
        if (to_f32) {
            float* dst = reinterpret_cast<float*>(sg.dst) + off;
            for (long long i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
        } else {
            bf16* dst = reinterpret_cast<bf16*>(sg.dst) + off;
            if (sg.dst_lo != nullptr) {
                bf16* dlo = reinterpret_cast<bf16*>(sg.dst_lo) + off;
                for (long long i = threadIdx.x; i < n; i += blockDim.x) {
                    const bf16 hi = __float2bfloat16_rn(src[i]);
                    dst[i] = hi;
                    dlo[i] = __float2bfloat16_rn(src[i] - __bfloat162float(hi));
                }
            } else if ((((uintptr_t)src & 15) | ((uintptr_t)dst & 7)) == 0) {
                const long long nv = n >> 2;
                for (long long i = threadIdx.x; i < nv; i += blockDim.x) {
                    const float4 v = reinterpret_cast<const float4*>(src)[i];
                    uint2 w;
                    w.x = pack_bf16x2(v.x, v.y);
                    w.y = pack_bf16x2(v.z, v.w);
                    reinterpret_cast<uint2*>(dst)[i] = w;
                }
                for (long long i = (nv << 2) + threadIdx.x; i < n; i += blockDim.x)
                    dst[i] = __float2bfloat16_rn(src[i]);
            } else {
                for (long long i = threadIdx.x; i < n; i += blockDim.x) dst[i] = __float2bfloat16_rn(src[i]);
            }
        }
    }
}

// Debug / experiment helper: occupies `gridDim.x` SMs (one CTA each, large dynamic smem) for
// `cycles` clock cycles.  Used by tools/contention_bench.py to emulate a co-running collective.
__global__ void hog_kernel(long long cycles, int* sink) {
    extern __shared__ uint8_t hog_smem[];
    const long long t0 = clock64();
    int x = 0;
    while (clock64() - t0 < cycles) x += hog_smem[threadIdx.x & 63];
    if (x == 123456789) *sink = x;
}

// out = act > 0 ? dy * scale : 0   (backward through ReLU + dropout using the saved activation)
__global__ void __launch_bounds__(256)
gate_bf16_kernel(const float* __restrict__ dy, const bf16* __restrict__ act, float scale,
                 bf16* __restrict__ out, long long n) {
    pdl_launch_dependents();
    pdl_wait();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = __float2bfloat16_rn(__bfloat162float(act[i]) > 0.f ? dy[i] * scale : 0.f);
}

// out[c] += sum_r x[r,c].  Block = 32 column groups x 8 row lanes; grid.y slices the rows.
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ x, long long rows, long long cols, long long ld,
              float* __restrict__ out, int vec_ok) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float s_part[8][32 * VEC + 1];
    const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const long long c0 = ((long long)blockIdx.x * 32 + cg) * VEC;
    float acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
    if (c0 < cols) {
        const long long rstride = (long long)gridDim.y * 8;
        const bool vec = vec_ok && (c0 + VEC <= cols);
        for (long long r = (long long)blockIdx.y * 8 + rl; r < rows; r += 4 * rstride) {
            if (vec) {
                // 4 independent 16-byte loads in flight per thread
                if constexpr (sizeof(T) == 2) {
                    uint4 v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        v[u] = (r + u * rstride < rows)
                                   ? *reinterpret_cast<const uint4*>(x + (r + u * rstride) * ld + c0)
                                   : make_uint4(0, 0, 0, 0);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        acc[0] += bf16_lo_to_f(v[u].x); acc[1] += bf16_hi_to_f(v[u].x);
                        acc[2] += bf16_lo_to_f(v[u].y); acc[3] += bf16_hi_to_f(v[u].y);
                        acc[4] += bf16_lo_to_f(v[u].z); acc[5] += bf16_hi_to_f(v[u].z);
                        acc[6] += bf16_lo_to_f(v[u].w); acc[7] += bf16_hi_to_f(v[u].w);
                    }
                } else {
                    float4 v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        v[u] = (r + u * rstride < rows)
                                   ? *reinterpret_cast<const float4*>(x + (r + u * rstride) * ld + c0)
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        acc[0] += v[u].x; acc[1] += v[u].y; acc[2] += v[u].z; acc[3] += v[u].w;
                    }
                }
            } else {
                for (int u = 0; u < 4; ++u) {
                    if (r + u * rstride >= rows) break;
                    const T* p = x + (r + u * rstride) * ld + c0;
                    for (int j = 0; j < VEC; ++j)
                        if (c0 + j < cols) acc[j] += (float)p[j];
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) s_part[rl][cg * VEC + j] = acc[j];
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * VEC; i += blockDim.x) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) s += s_part[r][i];
        const long long c = (long long)blockIdx.x * 32 * VEC + i;
        if (c < cols) atomicAdd(out + c, s);
    }
}

template <typename T, int VEC>
static int launch_colsum(const T* x, int64_t rows, int64_t cols, int64_t ld, float* out, void* stream) {
    MCAN_REQUIRE(x && out && rows > 0 && cols > 0, "mcan_colsum: bad args");
    const int vec_ok = (ld % VEC == 0 && ((uintptr_t)x & 15) == 0) ? 1 : 0;
    const int sms = device_num_sms();
    MCAN_REQUIRE(sms > 0, "mcan_colsum: no CUDA device");
    const int gx = (int)((cols + 32 * VEC - 1) / (32 * VEC));
    long long gy = (2LL * sms + gx - 1) / gx;
    const long long maxy = (rows + 63) / 64;
    if (gy > maxy) gy = maxy;
    if (gy < 1) gy = 1;
    MCAN_CHECK_CUDA(launch_kernel(colsum_kernel<T, VEC>, dim3(gx, (unsigned)gy), dim3(256), 0,
                                  reinterpret_cast<cudaStream_t>(stream), x, (long long)rows, (long long)cols,
                                  (long long)ld, out, vec_ok));
    return 0;
}

}  // namespace mcan

using namespace mcan;

extern "C" int mcan_cast_bf16(const float* x, int64_t n, void* hi, void* lo, void* stream) {
    MCAN_REQUIRE(x && hi && n > 0, "mcan_cast_bf16: bad args");
    MCAN_REQUIRE(((uintptr_t)x & 15) == 0 && (((uintptr_t)hi | (uintptr_t)lo) & 7) == 0, "mcan_cast_bf16: alignment");
    const int sms = device_num_sms();
    MCAN_REQUIRE(sms > 0, "mcan_cast_bf16: no CUDA device");
    long long blocks = ((n >> 2) + 255) / 256;
    if (blocks > 8LL * sms) blocks = 8LL * sms;
    if (blocks < 1) blocks = 1;
    MCAN_CHECK_CUDA(launch_kernel(cast_bf16_kernel, dim3((unsigned)blocks), dim3(256), 0,
                                  reinterpret_cast<cudaStream_t>(stream), x, (long long)n,
                                  reinterpret_cast<bf16*>(hi), reinterpret_cast<bf16*>(lo)));
    return 0;
}

extern "C" int mcan_colsum_bf16(const void* x, int64_t rows, int64_t cols, int64_t ld, float* out,
                                void* stream) {
    return launch_colsum<bf16, 8>(reinterpret_cast<const bf16*>(x), rows, cols, ld, out, stream);
}

extern "C" int mcan_colsum_f32(const float* x, int64_t rows, int64_t cols, int64_t ld, float* out,
                               void* stream) {
    return launch_colsum<float, 4>(x, rows, cols, ld, out, stream);
}

extern "C" int mcan_gate_bf16(const float* dy, const void* act, float scale, void* out, int64_t n,
                              void* stream) {
    MCAN_REQUIRE(dy && act && out && n > 0, "mcan_gate_bf16: bad args");
    const int sms = device_num_sms();
    MCAN_REQUIRE(sms > 0, "mcan_gate_bf16: no CUDA device");
    long long blocks = (n + 255) / 256;
    if (blocks > 16LL * sms) blocks = 16LL * sms;
    MCAN_CHECK_CUDA(launch_kernel(gate_bf16_kernel, dim3((unsigned)blocks), dim3(256), 0,
                                  reinterpret_cast<cudaStream_t>(stream), dy, reinterpret_cast<const bf16*>(act),
                                  scale, reinterpret_cast<bf16*>(out), (long long)n));
    return 0;
}

extern "C" int mcan_cast_multi(const void* seg_table_dev, int32_t num_segments, int64_t total_chunks,
                               void* stream) {
    MCAN_REQUIRE(seg_table_dev && num_segments > 0 && total_chunks > 0, "mcan_cast_multi: bad args");
    const int sms = device_num_sms();
    MCAN_REQUIRE(sms > 0, "mcan_cast_multi: no CUDA device");
    long long blocks = total_chunks < 16LL * sms ? total_chunks : 16LL * sms;
    MCAN_CHECK_CUDA(launch_kernel(cast_multi_kernel, dim3((unsigned)blocks), dim3(256), 0,
                                  reinterpret_cast<cudaStream_t>(stream),
                                  reinterpret_cast<const CastSeg*>(seg_table_dev), (int)num_segments,
                                  (long long)total_chunks));
    return 0;
}

extern "C" int mcan_debug_hog(int32_t ctas, int64_t cycles, int32_t smem_bytes, void* stream) {
    static int* sink = nullptr;
    if (!sink) MCAN_CHECK_CUDA(cudaMalloc(&sink, sizeof(int)));
    static int configured = 0;
    if (smem_bytes > configured) {
        MCAN_CHECK_CUDA(cudaFuncSetAttribute(hog_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        configured = smem_bytes;
    }
    hog_kernel<<<ctas, 64, smem_bytes, reinterpret_cast<cudaStream_t>(stream)>>>(cycles, sink);
    MCAN_CHECK_CUDA(cudaGetLastError());
    return 0;
}

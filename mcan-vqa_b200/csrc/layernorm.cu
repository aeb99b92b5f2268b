// L1: MCAN's custom LayerNorm (net_utils.py:48-60), forward and backward.
//
//   y = a_2 * (x - mean) / (std + eps) + b_2,   std = UNBIASED (N-1) standard deviation,
//   eps added to std (not to the variance) -- this is not F.layer_norm.
//
// Memory bound: one warp owns one row, the row lives in registers (float4 per lane), the
// reductions are warp shuffles, every global access is a coalesced 16-byte vector.
// The forward also emits the bf16 (and optional bf16 "lo") copy the next tcgen05 GEMM
// consumes; the backward also emits the dropout-gated bf16 gradient that feeds the
// dgrad/wgrad GEMMs of the layer that produced x, and the column sums for a_2, b_2 and
// that layer's bias.
#include "../../include/mcan_b200.h"
#include "common.cuh"

namespace mcan {

constexpr int kLnWarps = 8;

__device__ __forceinline__ void store_bf16x4(bf16* dst, float a, float b, float c, float d) {
    uint2 w;
    w.x = pack_bf16x2(a, b);
    w.y = pack_bf16x2(c, d);
    *reinterpret_cast<uint2*>(dst) = w;
}
__device__ __forceinline__ float bf16_round(float v) {
    return __bfloat162float(__float2bfloat16_rn(v));
}

// MAXV = float4 slots per lane; h <= MAXV * 128, h % 4 == 0.
template <int MAXV>
__global__ void __launch_bounds__(kLnWarps * 32)
ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ x2, float* __restrict__ s_out,
              long long rows, int h, const float* __restrict__ a2,
              const float* __restrict__ b2, float eps, float* __restrict__ y32,
              bf16* __restrict__ ybf, bf16* __restrict__ ylo, float* __restrict__ mean_out,
              float* __restrict__ sigma_out) {
    pdl_launch_dependents();
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * kLnWarps + warp;
    if (row >= rows) return;
    const int nv = h >> 2;
    const float4* xr = reinterpret_cast<const float4*>(x + row * h);
    float4 v[MAXV];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
        const int i = lane + 32 * j;
        v[j] = (i < nv) ? xr[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        if (x2 != nullptr && i < nv) {      // the normalised row is the sum of two inputs (proj_norm(lang + img))
            const float4 w = reinterpret_cast<const float4*>(x2 + row * h)[i];
            v[j].x += w.x; v[j].y += w.y; v[j].z += w.z; v[j].w += w.w;
            if (s_out != nullptr) reinterpret_cast<float4*>(s_out + row * h)[i] = v[j];
        }
        sum += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    }
    const float mean = warp_sum(sum) / (float)h;
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
        const int i = lane + 32 * j;
        if (i < nv) {
            const float cx = v[j].x - mean, cy = v[j].y - mean, cz = v[j].z - mean, cw = v[j].w - mean;
            sq += (cx * cx + cy * cy) + (cz * cz + cw * cw);
        }
    }
    const float sigma = sqrtf(warp_sum(sq) / (float)(h - 1));
    const float inv = 1.f / (sigma + eps);
    if (lane == 0) {
        if (mean_out) mean_out[row] = mean;
        if (sigma_out) sigma_out[row] = sigma;
    }
    const float4* ar = reinterpret_cast<const float4*>(a2);
    const float4* br = reinterpret_cast<const float4*>(b2);
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
        const int i = lane + 32 * j;
        if (i < nv) {
            const float4 a = __ldg(ar + i), b = __ldg(br + i);
            float4 y;
            y.x = a.x * (v[j].x - mean) * inv + b.x;
            y.y = a.y * (v[j].y - mean) * inv + b.y;
            y.z = a.z * (v[j].z - mean) * inv + b.z;
            y.w = a.w * (v[j].w - mean) * inv + b.w;
            if (y32) reinterpret_cast<float4*>(y32 + row * h)[i] = y;
            if (ybf) store_bf16x4(ybf + row * h + 4 * i, y.x, y.y, y.z, y.w);
            if (ylo)
                store_bf16x4(ylo + row * h + 4 * i, y.x - bf16_round(y.x), y.y - bf16_round(y.y),
                             y.z - bf16_round(y.z), y.w - bf16_round(y.w));
        }
    }
}

// Backward.  ghat = dy*a_2, c = x-mean, s = sigma+eps:
//   dx = (ghat - mean(ghat))/s - c * sum(ghat*c) / (s^2 * sigma * (N-1))
//   da_2 += dy*c/s ; db_2 += dy ; dbias += gated dx
// Single pass over HBM (x and dy are read exactly once): a warp owns one row at a time, the row
// lives in registers (like the forward), the two row statistics are warp shuffles.  The three
// column sums (a_2, b_2, bias of the producing layer) are accumulated in a PRIVATE shared-memory
// slice per warp (plain vector load/add/store, no atomics), summed over the warps of the CTA at
// the end and added to global memory with one atomicAdd per column per CTA.  The grid is
// persistent (2 CTAs per SM for h <= 1024), warps stride over the rows.
constexpr int kLnBwdWarps = 8;
constexpr int kLnBwdThreads = kLnBwdWarps * 32;

template <int MAXV>
__global__ void __launch_bounds__(kLnBwdThreads)
ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
              const float* __restrict__ mean_in, const float* __restrict__ sigma_in,
              const float* __restrict__ a2, float eps, long long rows, int h,
              float* __restrict__ dx32, bf16* __restrict__ dxbf, uint32_t drop_thr,
              float drop_scale, uint32_t drop_seed_in, const uint32_t* __restrict__ drop_seed_dev,
              float* __restrict__ da2, float* __restrict__ db2, float* __restrict__ dbias) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ __align__(16) float s_ln[];   // [warp][3][h]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nv = h >> 2;
    const uint32_t drop_seed =
        drop_seed_in ^ ((drop_thr != 0 && drop_seed_dev != nullptr) ? __ldg(drop_seed_dev) : 0U);
    const float4* ar = reinterpret_cast<const float4*>(a2);
    float4* acc_a = reinterpret_cast<float4*>(s_ln + (size_t)warp * 3 * h);
    float4* acc_b = acc_a + nv;
    float4* acc_bias = acc_b + nv;
    const bool want_ab = da2 != nullptr || db2 != nullptr;
    const bool want_bias = dbias != nullptr;
    bool first = true;
    const long long wstride = (long long)gridDim.x * kLnBwdWarps;
    for (long long row = (long long)blockIdx.x * kLnBwdWarps + warp; row < rows; row += wstride) {
        const float4* xr = reinterpret_cast<const float4*>(x + row * h);
        const float4* gr = reinterpret_cast<const float4*>(dy + row * h);
        float4 c[MAXV], g[MAXV];
#pragma unroll
        for (int j = 0; j < MAXV; ++j) {
            const int i = lane + 32 * j;
            c[j] = (i < nv) ? xr[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < MAXV; ++j) {
            const int i = lane + 32 * j;
            g[j] = (i < nv) ? gr[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const float mean = mean_in[row], sigma = sigma_in[row];
        const float s = sigma + eps, inv_s = 1.f / s;
        float sum_g = 0.f, sum_gc = 0.f;
#pragma unroll
        for (int j = 0; j < MAXV; ++j) {
            const int i = lane + 32 * j;
            if (i < nv) {
                const float4 a = __ldg(ar + i);
                c[j].x -= mean; c[j].y -= mean; c[j].z -= mean; c[j].w -= mean;
                if (want_ab) {
                    float4 ta, tb;
                    ta.x = g[j].x * c[j].x * inv_s; ta.y = g[j].y * c[j].y * inv_s;
                    ta.z = g[j].z * c[j].z * inv_s; ta.w = g[j].w * c[j].w * inv_s;
                    tb = g[j];
                    if (!first) {
                        const float4 pa = acc_a[i], pb = acc_b[i];
                        ta.x += pa.x; ta.y += pa.y; ta.z += pa.z; ta.w += pa.w;
                        tb.x += pb.x; tb.y += pb.y; tb.z += pb.z; tb.w += pb.w;
                    }
                    acc_a[i] = ta;
                    acc_b[i] = tb;
                }
                g[j].x *= a.x; g[j].y *= a.y; g[j].z *= a.z; g[j].w *= a.w;     // ghat
                sum_g += (g[j].x + g[j].y) + (g[j].z + g[j].w);
                sum_gc += (g[j].x * c[j].x + g[j].y * c[j].y) + (g[j].z * c[j].z + g[j].w * c[j].w);
            }
        }
        const float mg = warp_sum(sum_g) / (float)h;
        // sigma == 0: the reference's autograd gives NaN; emit the finite limit instead.
        const float sgc = warp_sum(sum_gc);
        const float k2 = (sigma > 0.f) ? sgc / (s * s * sigma * (float)(h - 1)) : 0.f;
#pragma unroll
        for (int j = 0; j < MAXV; ++j) {
            const int i = lane + 32 * j;
            if (i < nv) {
                float4 d;
                d.x = (g[j].x - mg) * inv_s - c[j].x * k2;
                d.y = (g[j].y - mg) * inv_s - c[j].y * k2;
                d.z = (g[j].z - mg) * inv_s - c[j].z * k2;
                d.w = (g[j].w - mg) * inv_s - c[j].w * k2;
                if (dx32) reinterpret_cast<float4*>(dx32 + row * h)[i] = d;
                if (dxbf != nullptr || want_bias) {
                    if (drop_thr != 0) {
                        const uint32_t base = (uint32_t)(row * h + 4 * i);  // multiple of 4
                        const uint32_t r0 = dropout_bits_pair(base >> 1, drop_seed);
                        const uint32_t r1 = dropout_bits_pair((base >> 1) + 1, drop_seed);
                        d.x = ((r0 & 0xFFFFU) >= drop_thr) ? d.x * drop_scale : 0.f;
                        d.y = ((r0 >> 16) >= drop_thr) ? d.y * drop_scale : 0.f;
                        d.z = ((r1 & 0xFFFFU) >= drop_thr) ? d.z * drop_scale : 0.f;
                        d.w = ((r1 >> 16) >= drop_thr) ? d.w * drop_scale : 0.f;
                    }
                    if (dxbf) store_bf16x4(dxbf + row * h + 4 * i, d.x, d.y, d.z, d.w);
                    if (want_bias) {
                        if (!first) {
                            const float4 pb = acc_bias[i];
                            d.x += pb.x; d.y += pb.y; d.z += pb.z; d.w += pb.w;
                        }
                        acc_bias[i] = d;
                    }
                }
            }
        }
        first = false;
    }
    if (!want_ab && !want_bias) return;
    // a warp that processed no row contributes zeros
    if (first) {
        for (int i = lane; i < 3 * nv; i += 32) acc_a[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    for (int col = threadIdx.x; col < h; col += kLnBwdThreads) {
        float ta = 0.f, tb = 0.f, tc = 0.f;
#pragma unroll
        for (int w = 0; w < kLnBwdWarps; ++w) {
            const float* base = s_ln + (size_t)w * 3 * h;
            ta += base[col];
            tb += base[h + col];
            tc += base[2 * h + col];
        }
        if (da2) atomicAdd(da2 + col, ta);
        if (db2) atomicAdd(db2 + col, tb);
        if (want_bias) atomicAdd(dbias + col, tc);
    }
}

int device_num_sms();

}  // namespace mcan

using namespace mcan;

static int layernorm_fwd_impl(const float* x, const float* x2, float* s_out, int64_t rows, int64_t h, const float* a2,
                              const float* b2, float eps, float* y_f32, void* y_bf16,
                              void* y_bf16_lo, float* mean, float* sigma, void* stream) {
    MCAN_REQUIRE(x && a2 && b2, "mcan_layernorm_fwd: null input");
    MCAN_REQUIRE((((uintptr_t)x2 | (uintptr_t)s_out) & 15) == 0, "mcan_layernorm_add_fwd: alignment");
    MCAN_REQUIRE(rows > 0 && h >= 8 && h % 4 == 0 && h <= 2048, "mcan_layernorm_fwd: rows=%lld h=%lld (h%%4==0, 8..2048)",
                 (long long)rows, (long long)h);
    MCAN_REQUIRE(rows * h < (1LL << 32), "mcan_layernorm_fwd: too large");
    MCAN_REQUIRE((((uintptr_t)x | (uintptr_t)a2 | (uintptr_t)b2 | (uintptr_t)y_f32) & 15) == 0 &&
                     (((uintptr_t)y_bf16 | (uintptr_t)y_bf16_lo) & 7) == 0,
                 "mcan_layernorm_fwd: alignment");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int grid = (int)((rows + kLnWarps - 1) / kLnWarps);
    bf16* ybf = reinterpret_cast<bf16*>(y_bf16);
    bf16* ylo = reinterpret_cast<bf16*>(y_bf16_lo);
#define LN_FWD(MV) MCAN_CHECK_CUDA(launch_kernel(ln_fwd_kernel<MV>, dim3(grid), dim3(kLnWarps * 32), 0, st, x, x2, s_out, (long long)rows, (int)h, a2, b2, eps, y_f32, ybf, ylo, mean, sigma))
    if (h <= 512) LN_FWD(4);
    else if (h <= 1024) LN_FWD(8);
    else LN_FWD(16);
#undef LN_FWD
    MCAN_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mcan_layernorm_fwd(const float* x, int64_t rows, int64_t h, const float* a2,
                                  const float* b2, float eps, float* y_f32, void* y_bf16,
                                  void* y_bf16_lo, float* mean, float* sigma, void* stream) {
    return layernorm_fwd_impl(x, nullptr, nullptr, rows, h, a2, b2, eps, y_f32, y_bf16, y_bf16_lo, mean, sigma, stream);
}

extern "C" int mcan_layernorm_add_fwd(const float* x, const float* x2, float* s_out, int64_t rows, int64_t h,
                                      const float* a2, const float* b2, float eps, float* y_f32, void* y_bf16,
                                      void* y_bf16_lo, float* mean, float* sigma, void* stream) {
    MCAN_REQUIRE(x2 != nullptr, "mcan_layernorm_add_fwd: null x2");
    return layernorm_fwd_impl(x, x2, s_out, rows, h, a2, b2, eps, y_f32, y_bf16, y_bf16_lo, mean, sigma, stream);
}

extern "C" int mcan_layernorm_bwd(const float* dy, const float* x, const float* mean,
                                  const float* sigma, const float* a2, float eps, int64_t rows,
                                  int64_t h, float* dx_f32, void* dx_bf16, float dropout_p,
                                  uint32_t dropout_seed, const uint32_t* dropout_seed_dev,
                                  float* da2, float* db2, float* dbias, void* stream) {
    MCAN_REQUIRE(dy && x && mean && sigma && a2, "mcan_layernorm_bwd: null input");
    MCAN_REQUIRE(rows > 0 && h >= 8 && h % 4 == 0 && h <= 2048, "mcan_layernorm_bwd: rows=%lld h=%lld",
                 (long long)rows, (long long)h);
    MCAN_REQUIRE(rows * h < (1LL << 32), "mcan_layernorm_bwd: too large");
    MCAN_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "mcan_layernorm_bwd: dropout_p=%f", dropout_p);
    MCAN_REQUIRE((((uintptr_t)dy | (uintptr_t)x | (uintptr_t)a2 | (uintptr_t)dx_f32) & 15) == 0 &&
                     ((uintptr_t)dx_bf16 & 7) == 0,
                 "mcan_layernorm_bwd: alignment");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int sms = device_num_sms();
    MCAN_REQUIRE(sms > 0, "mcan_layernorm_bwd: no CUDA device");
    const uint32_t thr = dropout_p > 0.f ? dropout_threshold(dropout_p) : 0;
    const float scale = dropout_p > 0.f ? 1.f / (1.f - dropout_p) : 1.f;
    const size_t smem = (size_t)kLnBwdWarps * 3 * (size_t)h * sizeof(float);
    const int per_sm = h <= 1024 ? 2 : 1;
    long long grid = (rows + kLnBwdWarps - 1) / kLnBwdWarps;
    if (grid > (long long)per_sm * sms) grid = (long long)per_sm * sms;
#define LN_BWD(MV)                                                                                          \
    do {                                                                                                    \
        static size_t configured = 48 * 1024;                                                               \
        if (smem > configured) {                                                                            \
            MCAN_CHECK_CUDA(cudaFuncSetAttribute(ln_bwd_kernel<MV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                 (int)smem));                                               \
            configured = smem;                                                                              \
        }                                                                                                   \
        MCAN_CHECK_CUDA(launch_kernel(ln_bwd_kernel<MV>, dim3((unsigned)grid), dim3(kLnBwdThreads), smem, st, dy, x, \
                                      mean, sigma, a2, eps, (long long)rows, (int)h, dx_f32,                \
                                      reinterpret_cast<bf16*>(dx_bf16), thr, scale, dropout_seed,           \
                                      dropout_seed_dev, da2, db2, dbias));                                  \
    } while (0)
    if (h <= 512) LN_BWD(4);
    else if (h <= 1024) LN_BWD(8);
    else LN_BWD(16);
#undef LN_BWD
    return 0;
}

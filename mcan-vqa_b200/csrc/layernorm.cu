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
ln_fwd_kernel(const float* __restrict__ x, long long rows, int h, const float* __restrict__ a2,
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
// A CTA owns a block of up to 32 consecutive rows.  Pass A (warp per row, streaming, a handful
// of registers) reduces the two row statistics into shared memory.  Pass B re-reads the block
// from L1/L2 with the threads laid out along the columns: each thread owns one float4 column
// group for all rows of the block, so the a_2 / b_2 / bias column sums are 12 registers and
// every global access is a coalesced 16-byte vector; one atomicAdd per column per CTA.
constexpr int kLnBwdMaxRows = 32;
constexpr int kLnBwdThreads = 256;

__global__ void __launch_bounds__(kLnBwdThreads)
ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
              const float* __restrict__ mean_in, const float* __restrict__ sigma_in,
              const float* __restrict__ a2, float eps, long long rows, int h, int rows_per_cta,
              float* __restrict__ dx32, bf16* __restrict__ dxbf, uint32_t drop_thr,
              float drop_scale, uint32_t drop_seed_in, const uint32_t* __restrict__ drop_seed_dev,
              float* __restrict__ da2, float* __restrict__ db2, float* __restrict__ dbias) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float s_mean[kLnBwdMaxRows], s_invs[kLnBwdMaxRows], s_mg[kLnBwdMaxRows], s_k2[kLnBwdMaxRows];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nv = h >> 2;
    const long long row_base = (long long)blockIdx.x * rows_per_cta;
    const int nrows = (int)min((long long)rows_per_cta, rows - row_base);
    const uint32_t drop_seed =
        drop_seed_in ^ ((drop_thr != 0 && drop_seed_dev != nullptr) ? __ldg(drop_seed_dev) : 0U);
    const float4* ar = reinterpret_cast<const float4*>(a2);

    // ---- pass A: row statistics ----
    for (int r = warp; r < nrows; r += kLnBwdThreads / 32) {
        const long long row = row_base + r;
        const float mean = mean_in[row], sigma = sigma_in[row];
        const float4* xr = reinterpret_cast<const float4*>(x + row * h);
        const float4* gr = reinterpret_cast<const float4*>(dy + row * h);
        float sum_g = 0.f, sum_gc = 0.f;
        for (int i = lane; i < nv; i += 32) {
            const float4 xv = xr[i], gv = gr[i], a = __ldg(ar + i);
            const float g0 = gv.x * a.x, g1 = gv.y * a.y, g2 = gv.z * a.z, g3 = gv.w * a.w;
            sum_g += (g0 + g1) + (g2 + g3);
            sum_gc += (g0 * (xv.x - mean) + g1 * (xv.y - mean)) + (g2 * (xv.z - mean) + g3 * (xv.w - mean));
        }
        sum_g = warp_sum(sum_g);
        sum_gc = warp_sum(sum_gc);
        if (lane == 0) {
            const float s = sigma + eps;
            s_mean[r] = mean;
            s_invs[r] = 1.f / s;
            s_mg[r] = sum_g / (float)h;
            // sigma == 0: the reference's autograd gives NaN; emit the finite limit instead.
            s_k2[r] = (sigma > 0.f) ? sum_gc / (s * s * sigma * (float)(h - 1)) : 0.f;
        }
    }
    __syncthreads();

    // ---- pass B: dx and the column sums ----
    for (int c = threadIdx.x; c < nv; c += kLnBwdThreads) {
        const float4 a = __ldg(ar + c);
        float4 acc_a = make_float4(0.f, 0.f, 0.f, 0.f), acc_b = acc_a, acc_bias = acc_a;
#pragma unroll 4
        for (int r = 0; r < nrows; ++r) {
            const long long row = row_base + r;
            const float4 xv = reinterpret_cast<const float4*>(x + row * h)[c];
            const float4 gv = reinterpret_cast<const float4*>(dy + row * h)[c];
            const float mean = s_mean[r], inv_s = s_invs[r], mg = s_mg[r], k2 = s_k2[r];
            const float cx = xv.x - mean, cy = xv.y - mean, cz = xv.z - mean, cw = xv.w - mean;
            float4 d;
            d.x = (gv.x * a.x - mg) * inv_s - cx * k2;
            d.y = (gv.y * a.y - mg) * inv_s - cy * k2;
            d.z = (gv.z * a.z - mg) * inv_s - cz * k2;
            d.w = (gv.w * a.w - mg) * inv_s - cw * k2;
            acc_a.x += gv.x * cx * inv_s; acc_a.y += gv.y * cy * inv_s;
            acc_a.z += gv.z * cz * inv_s; acc_a.w += gv.w * cw * inv_s;
            acc_b.x += gv.x; acc_b.y += gv.y; acc_b.z += gv.z; acc_b.w += gv.w;
            if (dx32) reinterpret_cast<float4*>(dx32 + row * h)[c] = d;
            if (dxbf != nullptr || dbias != nullptr) {
                float4 gd = d;
                if (drop_thr != 0) {
                    const uint32_t base = (uint32_t)(row * h + 4 * c);  // multiple of 4
                    const uint32_t r0 = dropout_bits_pair(base >> 1, drop_seed);
                    const uint32_t r1 = dropout_bits_pair((base >> 1) + 1, drop_seed);
                    gd.x = ((r0 & 0xFFFFU) >= drop_thr) ? d.x * drop_scale : 0.f;
                    gd.y = ((r0 >> 16) >= drop_thr) ? d.y * drop_scale : 0.f;
                    gd.z = ((r1 & 0xFFFFU) >= drop_thr) ? d.z * drop_scale : 0.f;
                    gd.w = ((r1 >> 16) >= drop_thr) ? d.w * drop_scale : 0.f;
                }
                if (dxbf) store_bf16x4(dxbf + row * h + 4 * c, gd.x, gd.y, gd.z, gd.w);
                acc_bias.x += gd.x; acc_bias.y += gd.y; acc_bias.z += gd.z; acc_bias.w += gd.w;
            }
        }
        if (da2) {
            atomicAdd(da2 + 4 * c + 0, acc_a.x); atomicAdd(da2 + 4 * c + 1, acc_a.y);
            atomicAdd(da2 + 4 * c + 2, acc_a.z); atomicAdd(da2 + 4 * c + 3, acc_a.w);
        }
        if (db2) {
            atomicAdd(db2 + 4 * c + 0, acc_b.x); atomicAdd(db2 + 4 * c + 1, acc_b.y);
            atomicAdd(db2 + 4 * c + 2, acc_b.z); atomicAdd(db2 + 4 * c + 3, acc_b.w);
        }
        if (dbias) {
            atomicAdd(dbias + 4 * c + 0, acc_bias.x); atomicAdd(dbias + 4 * c + 1, acc_bias.y);
            atomicAdd(dbias + 4 * c + 2, acc_bias.z); atomicAdd(dbias + 4 * c + 3, acc_bias.w);
        }
    }
}

int device_num_sms();

}  // namespace mcan

using namespace mcan;

extern "C" int mcan_layernorm_fwd(const float* x, int64_t rows, int64_t h, const float* a2,
                                  const float* b2, float eps, float* y_f32, void* y_bf16,
                                  void* y_bf16_lo, float* mean, float* sigma, void* stream) {
    MCAN_REQUIRE(x && a2 && b2, "mcan_layernorm_fwd: null input");
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
#define LN_FWD(MV) MCAN_CHECK_CUDA(launch_kernel(ln_fwd_kernel<MV>, dim3(grid), dim3(kLnWarps * 32), 0, st, x, (long long)rows, (int)h, a2, b2, eps, y_f32, ybf, ylo, mean, sigma))
    if (h <= 512) LN_FWD(4);
    else if (h <= 1024) LN_FWD(8);
    else LN_FWD(16);
#undef LN_FWD
    MCAN_CHECK_CUDA(cudaGetLastError());
    return 0;
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
    // rows per CTA: 32 when that still gives >= 1 CTA per SM, fewer for short inputs
    int rpc = kLnBwdMaxRows;
    while (rpc > 4 && (rows + rpc - 1) / rpc < sms) rpc >>= 1;
    const int grid = (int)((rows + rpc - 1) / rpc);
    const uint32_t thr = dropout_p > 0.f ? dropout_threshold(dropout_p) : 0;
    const float scale = dropout_p > 0.f ? 1.f / (1.f - dropout_p) : 1.f;
    MCAN_CHECK_CUDA(launch_kernel(ln_bwd_kernel, dim3(grid), dim3(kLnBwdThreads), 0, st, dy, x, mean, sigma, a2,
                                  eps, (long long)rows, (int)h, rpc, dx_f32, reinterpret_cast<bf16*>(dx_bf16), thr,
                                  scale, dropout_seed, dropout_seed_dev, da2, db2, dbias));
    return 0;
}

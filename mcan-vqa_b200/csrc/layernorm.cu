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
// Each warp walks rows with stride gridDim*warps and keeps per-lane column partial sums in
// registers; the CTA folds them through shared memory and issues one atomicAdd per column.
template <int MAXV>
__global__ void __launch_bounds__(kLnWarps * 32)
ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
              const float* __restrict__ mean_in, const float* __restrict__ sigma_in,
              const float* __restrict__ a2, float eps, long long rows, int h,
              float* __restrict__ dx32, bf16* __restrict__ dxbf, uint32_t drop_thr,
              float drop_scale, uint32_t drop_seed_in, const uint32_t* __restrict__ drop_seed_dev,
              float* __restrict__ da2,
              float* __restrict__ db2, float* __restrict__ dbias) {
    extern __shared__ float s_red[];  // [3][h]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nv = h >> 2;
    const uint32_t drop_seed =
        drop_seed_in ^ ((drop_thr != 0 && drop_seed_dev != nullptr) ? __ldg(drop_seed_dev) : 0U);
    for (int i = threadIdx.x; i < 3 * h; i += blockDim.x) s_red[i] = 0.f;
    __syncthreads();

    float4 acc_a[MAXV], acc_b[MAXV], acc_bias[MAXV];
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
        acc_a[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        acc_b[j] = acc_a[j];
        acc_bias[j] = acc_a[j];
    }
    const float4* ar = reinterpret_cast<const float4*>(a2);
    const float inv_n = 1.f / (float)h;

    for (long long row = (long long)blockIdx.x * kLnWarps + warp; row < rows;
         row += (long long)gridDim.x * kLnWarps) {
        const float mean = mean_in[row], sigma = sigma_in[row];
        const float s = sigma + eps, inv_s = 1.f / s;
        const float4* xr = reinterpret_cast<const float4*>(x + row * h);
        const float4* gr = reinterpret_cast<const float4*>(dy + row * h);
        float4 c[MAXV], gh[MAXV];
        float sum_g = 0.f, sum_gc = 0.f;
#pragma unroll
        for (int j = 0; j < MAXV; ++j) {
            const int i = lane + 32 * j;
            if (i < nv) {
                const float4 xv = xr[i], gv = gr[i], a = __ldg(ar + i);
                c[j] = make_float4(xv.x - mean, xv.y - mean, xv.z - mean, xv.w - mean);
                gh[j] = make_float4(gv.x * a.x, gv.y * a.y, gv.z * a.z, gv.w * a.w);
                sum_g += (gh[j].x + gh[j].y) + (gh[j].z + gh[j].w);
                sum_gc += (gh[j].x * c[j].x + gh[j].y * c[j].y) + (gh[j].z * c[j].z + gh[j].w * c[j].w);
                acc_a[j].x += gv.x * c[j].x * inv_s; acc_a[j].y += gv.y * c[j].y * inv_s;
                acc_a[j].z += gv.z * c[j].z * inv_s; acc_a[j].w += gv.w * c[j].w * inv_s;
                acc_b[j].x += gv.x; acc_b[j].y += gv.y; acc_b[j].z += gv.z; acc_b[j].w += gv.w;
            } else {
                c[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                gh[j] = c[j];
            }
        }
        sum_g = warp_sum(sum_g);
        sum_gc = warp_sum(sum_gc);
        const float mg = sum_g * inv_n;
        // sigma == 0: the reference's autograd gives NaN; emit the finite limit instead.
        const float k2 = (sigma > 0.f) ? sum_gc / (s * s * sigma * (float)(h - 1)) : 0.f;
#pragma unroll
        for (int j = 0; j < MAXV; ++j) {
            const int i = lane + 32 * j;
            if (i < nv) {
                float4 d;
                d.x = (gh[j].x - mg) * inv_s - c[j].x * k2;
                d.y = (gh[j].y - mg) * inv_s - c[j].y * k2;
                d.z = (gh[j].z - mg) * inv_s - c[j].z * k2;
                d.w = (gh[j].w - mg) * inv_s - c[j].w * k2;
                if (dx32) reinterpret_cast<float4*>(dx32 + row * h)[i] = d;
                if (dxbf != nullptr || dbias != nullptr) {
                    float4 gd = d;
                    if (drop_thr != 0) {
                        const uint32_t base = (uint32_t)(row * h + 4 * i);  // multiple of 4
                        const uint32_t r0 = dropout_bits_pair(base >> 1, drop_seed);
                        const uint32_t r1 = dropout_bits_pair((base >> 1) + 1, drop_seed);
                        gd.x = ((r0 & 0xFFFFU) >= drop_thr) ? d.x * drop_scale : 0.f;
                        gd.y = ((r0 >> 16) >= drop_thr) ? d.y * drop_scale : 0.f;
                        gd.z = ((r1 & 0xFFFFU) >= drop_thr) ? d.z * drop_scale : 0.f;
                        gd.w = ((r1 >> 16) >= drop_thr) ? d.w * drop_scale : 0.f;
                    }
                    if (dxbf) store_bf16x4(dxbf + row * h + 4 * i, gd.x, gd.y, gd.z, gd.w);
                    acc_bias[j].x += gd.x; acc_bias[j].y += gd.y;
                    acc_bias[j].z += gd.z; acc_bias[j].w += gd.w;
                }
            }
        }
    }
    // fold the per-lane partial sums: registers -> shared atomics -> one global atomic per column
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
        const int i = lane + 32 * j;
        if (i < nv) {
            float* sa = s_red + 4 * i;
            atomicAdd(sa + 0, acc_a[j].x); atomicAdd(sa + 1, acc_a[j].y);
            atomicAdd(sa + 2, acc_a[j].z); atomicAdd(sa + 3, acc_a[j].w);
            float* sb = s_red + h + 4 * i;
            atomicAdd(sb + 0, acc_b[j].x); atomicAdd(sb + 1, acc_b[j].y);
            atomicAdd(sb + 2, acc_b[j].z); atomicAdd(sb + 3, acc_b[j].w);
            if (dbias) {
                float* sc = s_red + 2 * h + 4 * i;
                atomicAdd(sc + 0, acc_bias[j].x); atomicAdd(sc + 1, acc_bias[j].y);
                atomicAdd(sc + 2, acc_bias[j].z); atomicAdd(sc + 3, acc_bias[j].w);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < h; i += blockDim.x) {
        if (da2) atomicAdd(da2 + i, s_red[i]);
        if (db2) atomicAdd(db2 + i, s_red[h + i]);
        if (dbias) atomicAdd(dbias + i, s_red[2 * h + i]);
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
#define LN_FWD(MV) ln_fwd_kernel<MV><<<grid, kLnWarps * 32, 0, st>>>(x, rows, (int)h, a2, b2, eps, y_f32, ybf, ylo, mean, sigma)
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
    long long want = (rows + kLnWarps - 1) / kLnWarps;
    const int grid = (int)(want < 2LL * sms ? want : 2LL * sms);
    const uint32_t thr = dropout_p > 0.f ? dropout_threshold(dropout_p) : 0;
    const float scale = dropout_p > 0.f ? 1.f / (1.f - dropout_p) : 1.f;
    const size_t smem = (size_t)3 * h * sizeof(float);
    bf16* dbf = reinterpret_cast<bf16*>(dx_bf16);
#define LN_BWD(MV) ln_bwd_kernel<MV><<<grid, kLnWarps * 32, smem, st>>>(dy, x, mean, sigma, a2, eps, rows, (int)h, dx_f32, dbf, thr, scale, dropout_seed, dropout_seed_dev, da2, db2, dbias)
    if (h <= 512) LN_BWD(4);
    else if (h <= 1024) LN_BWD(8);
    else LN_BWD(16);
#undef LN_BWD
    MCAN_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// The two ends of the hot path:
//
//   * image side input: fp32 region features -> bf16 GEMM operand AND the zero-row mask in ONE pass
//     (reference core/model/net.py:100, 135-137: make_mask = (sum(|feature|, -1) == 0), followed by
//     img_feat_linear on the same tensor) -- the features are read from HBM exactly once;
//   * output head: sigmoid + BCELoss(reduction='sum') forward, and its backward straight into the bf16
//     operand of the proj wgrad / dgrad GEMMs plus the proj bias gradient
//     (reference core/model/net.py:127-129, core/exec.py:67,178).
//
// All three kernels are HBM/latency bound element-wise passes; coalesced 16-byte accesses where the
// leading dimensions allow it.
#include "../../include/mcan_b200.h"
#include "common.cuh"

namespace mcan {

int device_num_sms();

// One warp per feature row: cast to bf16 (hi [+ lo]) and test the row for "all zero".
// sum_j |x_j| == 0  <=>  every x_j == +-0 (NaN / inf rows are not masked, like the reference's sum).
constexpr int kRmWarps = 8;

__global__ void __launch_bounds__(kRmWarps * 32)
rowmask_cast_kernel(const float* __restrict__ x, long long rows, int cols, bf16* __restrict__ hi,
                    bf16* __restrict__ lo, long long ld, uint8_t* __restrict__ mask) {
    pdl_launch_dependents();
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * kRmWarps + warp;
    if (row >= rows) return;
    const float* xr = x + row * cols;
    bf16* hr = hi + row * ld;
    bf16* lr = lo ? lo + row * ld : nullptr;
    bool nonzero = false;
    const bool vec = (cols % 4 == 0) && (ld % 4 == 0);
    if (vec) {
        const int nv = cols >> 2;
        for (int i = lane; i < nv; i += 32) {
            const float4 v = reinterpret_cast<const float4*>(xr)[i];
            nonzero |= !(v.x == 0.f && v.y == 0.f && v.z == 0.f && v.w == 0.f);
            uint2 w;
            w.x = pack_bf16x2(v.x, v.y);
            w.y = pack_bf16x2(v.z, v.w);
            reinterpret_cast<uint2*>(hr)[i] = w;
            if (lr) {
                uint2 l;
                l.x = pack_bf16x2(v.x - bf16_lo_to_f(w.x), v.y - bf16_hi_to_f(w.x));
                l.y = pack_bf16x2(v.z - bf16_lo_to_f(w.y), v.w - bf16_hi_to_f(w.y));
                reinterpret_cast<uint2*>(lr)[i] = l;
            }
        }
    } else {
        for (int i = lane; i < cols; i += 32) {
            const float v = xr[i];
            nonzero |= !(v == 0.f);
            const bf16 h = __float2bfloat16_rn(v);
            hr[i] = h;
            if (lr) lr[i] = __float2bfloat16_rn(v - __bfloat162float(h));
        }
    }
    const uint32_t any = __ballot_sync(0xffffffffU, nonzero);
    if (lane == 0 && mask != nullptr) mask[row] = any ? 0 : 1;
}

// ---------------------------------------------------------------------------------------------
// sigmoid + BCE(sum).  Thread = one answer column, blockIdx.y = a group of rows; the loss partials of
// the CTAs are summed in a FIXED order by the last CTA to finish (ticket in the workspace), so the loss is
// bit-reproducible run to run.
//   p    = 1 / (1 + exp(-z))
//   loss = sum -( t * max(log p, -100) + (1 - t) * max(log1p(-p), -100) )      (torch BCELoss clamps)
// workspace: [0] ticket (uint32, zero on entry, reset by the last CTA), [1 ...] partials.
// ---------------------------------------------------------------------------------------------
constexpr int kHeadThreads = 256;
constexpr int kHeadMaxCtas = 1024;

__global__ void __launch_bounds__(kHeadThreads)
sigmoid_bce_fwd_kernel(const float* __restrict__ logits, long long ld, const float* __restrict__ target,
                       int rows, int cols, int rows_per_cta, float* __restrict__ probs,
                       float* __restrict__ loss, float* __restrict__ workspace) {
    pdl_launch_dependents();
    pdl_wait();
    const int c = blockIdx.x * kHeadThreads + threadIdx.x;
    const int r0 = blockIdx.y * rows_per_cta;
    const int r1 = min(rows, r0 + rows_per_cta);
    float part = 0.f;
    if (c < cols) {
        for (int r = r0; r < r1; ++r) {
            const float z = logits[(long long)r * ld + c];
            const float p = 1.f / (1.f + expf(-z));
            probs[(long long)r * cols + c] = p;
            if (target != nullptr) {
                const float t = target[(long long)r * cols + c];
                part -= t * fmaxf(logf(p), -100.f) + (1.f - t) * fmaxf(log1pf(-p), -100.f);
            }
        }
    }
    if (loss == nullptr) return;
    __shared__ float s_part[kHeadThreads / 32];
    __shared__ bool s_last;
    part = warp_sum(part);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = part;
    __syncthreads();
    const int ncta = gridDim.x * gridDim.y;
    const int cta = blockIdx.y * gridDim.x + blockIdx.x;
    if (threadIdx.x == 0) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kHeadThreads / 32; ++w) s += s_part[w];
        workspace[1 + cta] = s;
        __threadfence();
        const uint32_t ticket = atomicAdd(reinterpret_cast<uint32_t*>(workspace), 1U);
        s_last = (ticket == (uint32_t)ncta - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    float s = 0.f;
    for (int i = threadIdx.x; i < ncta; i += kHeadThreads) s += __ldcg(workspace + 1 + i);
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < kHeadThreads / 32; ++w) tot += s_part[w];
        *loss = tot;
        *reinterpret_cast<uint32_t*>(workspace) = 0U;    // ready for the next launch
    }
}

// Backward.  With `target`: d loss / d z for loss = BCE(sum) scaled by the device scalar *gscale (1 if NULL),
//   dz = g * (p - t) / max((1 - p) p, 1e-12) * (1 - p) p          (torch's BCELoss x sigmoid backward);
// without: the plain sigmoid backward of an element-wise incoming gradient gout[rows, cols],
//   dz = gout * (1 - p) p.
// Written as bf16 with leading dimension ld (the wgrad / dgrad operand; pad columns are zeroed) and summed
// over the rows into dbias (the proj bias gradient, fp32 atomics on a zeroed buffer).
__global__ void __launch_bounds__(kHeadThreads)
sigmoid_bce_bwd_kernel(const float* __restrict__ probs, const float* __restrict__ target,
                       const float* __restrict__ gout, const float* __restrict__ gscale, int rows, int cols,
                       int rows_per_cta, bf16* __restrict__ dz, long long ld, float* __restrict__ dbias) {
    pdl_launch_dependents();
    pdl_wait();
    const int c = blockIdx.x * kHeadThreads + threadIdx.x;
    const int r0 = blockIdx.y * rows_per_cta;
    const int r1 = min(rows, r0 + rows_per_cta);
    if (c >= ld) return;
    if (c >= cols) {
        for (int r = r0; r < r1; ++r) dz[(long long)r * ld + c] = __float2bfloat16_rn(0.f);
        return;
    }
    const float g = (target != nullptr && gscale != nullptr) ? __ldg(gscale) : 1.f;
    float colsum = 0.f;
    for (int r = r0; r < r1; ++r) {
        const long long i = (long long)r * cols + c;
        const float p = probs[i];
        const float pq = (1.f - p) * p;
        float d;
        if (target != nullptr)
            d = g * (p - target[i]) / fmaxf(pq, 1e-12f) * pq;
        else
            d = gout[i] * pq;
        const bf16 b = __float2bfloat16_rn(d);
        dz[(long long)r * ld + c] = b;
        colsum += __bfloat162float(b);
    }
    if (dbias != nullptr) atomicAdd(dbias + c, colsum);
}

static void head_grid(int rows, int cols_padded, int sms, dim3* grid, int* rows_per_cta) {
    const int gx = (cols_padded + kHeadThreads - 1) / kHeadThreads;
    int gy = (2 * sms + gx - 1) / gx;
    if (gy > rows) gy = rows;
    if (gy < 1) gy = 1;
    int rpc = (rows + gy - 1) / gy;
    gy = (rows + rpc - 1) / rpc;
    while ((long long)gx * gy > kHeadMaxCtas) {
        ++rpc;
        gy = (rows + rpc - 1) / rpc;
    }
    *grid = dim3((unsigned)gx, (unsigned)gy);
    *rows_per_cta = rpc;
}

}  // namespace mcan

using namespace mcan;

extern "C" int mcan_rowmask_cast(const float* x, int64_t rows, int64_t cols, void* hi, void* lo, int64_t ld,
                                 uint8_t* mask, void* stream) {
    MCAN_REQUIRE(x && hi && rows > 0 && cols > 0 && ld >= cols, "mcan_rowmask_cast: bad args");
    MCAN_REQUIRE(cols < (1LL << 31), "mcan_rowmask_cast: too wide");
    MCAN_REQUIRE(((uintptr_t)x & 15) == 0 && (((uintptr_t)hi | (uintptr_t)lo) & 7) == 0, "mcan_rowmask_cast: alignment");
    const long long grid = (rows + kRmWarps - 1) / kRmWarps;
    MCAN_CHECK_CUDA(launch_kernel(rowmask_cast_kernel, dim3((unsigned)grid), dim3(kRmWarps * 32), 0,
                                  reinterpret_cast<cudaStream_t>(stream), x, (long long)rows, (int)cols,
                                  reinterpret_cast<bf16*>(hi), reinterpret_cast<bf16*>(lo), (long long)ld, mask));
    return 0;
}

extern "C" int mcan_sigmoid_bce_fwd(const float* logits, int64_t ld, const float* target, int32_t rows, int32_t cols,
                                    float* probs, float* loss, float* workspace, void* stream) {
    MCAN_REQUIRE(logits && probs && rows > 0 && cols > 0 && ld >= cols, "mcan_sigmoid_bce_fwd: bad args");
    MCAN_REQUIRE((loss == nullptr) || (target != nullptr && workspace != nullptr),
                 "mcan_sigmoid_bce_fwd: the loss needs a target and a workspace");
    const int sms = device_num_sms();
    MCAN_REQUIRE(sms > 0, "mcan_sigmoid_bce_fwd: no CUDA device");
    dim3 grid;
    int rpc;
    head_grid(rows, cols, sms, &grid, &rpc);
    MCAN_CHECK_CUDA(launch_kernel(sigmoid_bce_fwd_kernel, grid, dim3(kHeadThreads), 0,
                                  reinterpret_cast<cudaStream_t>(stream), logits, (long long)ld, target, (int)rows,
                                  (int)cols, rpc, probs, loss, workspace));
    return 0;
}

extern "C" int mcan_sigmoid_bce_bwd(const float* probs, const float* target, const float* gout, const float* gscale_dev,
                                    int32_t rows, int32_t cols, void* dz_bf16, int64_t ld, float* dbias,
                                    void* stream) {
    MCAN_REQUIRE(probs && dz_bf16 && rows > 0 && cols > 0 && ld >= cols, "mcan_sigmoid_bce_bwd: bad args");
    MCAN_REQUIRE((target != nullptr) != (gout != nullptr),
                 "mcan_sigmoid_bce_bwd: exactly one of target (fused BCE) and gout (plain sigmoid backward)");
    const int sms = device_num_sms();
    MCAN_REQUIRE(sms > 0, "mcan_sigmoid_bce_bwd: no CUDA device");
    dim3 grid;
    int rpc;
    head_grid(rows, (int)ld, sms, &grid, &rpc);
    MCAN_CHECK_CUDA(launch_kernel(sigmoid_bce_bwd_kernel, grid, dim3(kHeadThreads), 0,
                                  reinterpret_cast<cudaStream_t>(stream), probs, target, gout, gscale_dev, (int)rows,
                                  (int)cols, rpc, reinterpret_cast<bf16*>(dz_bf16), (long long)ld, dbias));
    return 0;
}

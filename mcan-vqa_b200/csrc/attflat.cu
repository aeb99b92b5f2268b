// F1: AttFlat attention pooling (net.py:38-55), forward and backward.
//
// The H -> flat_mlp_size projection with ReLU/dropout (86 % of AttFlat's FLOPs) is a tcgen05
// GEMM (mcan_gemm, relu + dropout epilogue) that leaves hmid in bf16.  Everything after it --
// the flat_mlp_size -> glimpses projection, masked_fill(-1e9), the softmax over the SEQUENCE
// dimension and the glimpse-weighted sums -- is this one kernel: logits and attention weights
// stay in shared memory, x is streamed once with coalesced float4 loads.
//
// Grid = (slices, batch).  A batch of 64 samples alone would leave 84 of the 148 SMs idle (42 / 61 us
// per launch), so every sample is worked on by `slices` CTAs:
//   forward : each CTA recomputes the S x G logits and their softmax (hmid is 100 KB per sample and
//             comes from L2 after the first reader) and owns H / slices columns of the pooled sums;
//   backward: each CTA owns S / slices rows.  The only quantity of the softmax backward that couples
//             the rows, sum_s att[s,g] * datt[s,g], equals dpooled[g,:] . pooled[g,:] (the forward's
//             fp32 output), so no exchange between the CTAs of a sample is needed.
#include "../../include/mcan_b200.h"
#include "common.cuh"
#include <stdlib.h>

namespace mcan {

int device_num_sms();

constexpr int kFlatThreads = 512;          // upper bound; the launch picks the block size (flat_threads())
constexpr int kFlatMaxSeq = 128;
constexpr int kFlatMaxGlimpses = 8;

template <int kRows>
__global__ void __launch_bounds__(kFlatThreads)
attflat_pool_fwd_kernel(const bf16* __restrict__ hmid, const bf16* __restrict__ hmid_lo,
                        const float* __restrict__ w2,
                        const float* __restrict__ b2, const uint8_t* __restrict__ mask,
                        const float* __restrict__ x, int S, int H, int M, int G,
                        float* __restrict__ att_w, float* __restrict__ pooled32,
                        bf16* __restrict__ pooledbf) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float s_att[kFlatMaxSeq * kFlatMaxGlimpses];
    __shared__ __align__(16) float4 s_red[kFlatThreads * 4];
    const int b = blockIdx.y, slice = blockIdx.x, nslices = gridDim.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;

    // logits[s,g] = hmid[s,:] . w2[g,:] + b2[g]; masked -> -1e9.  A warp works on kRows rows at a time so that
    // their loads are in flight together (the phase is latency bound: one row per warp at a time took 26 us).
    for (int s0 = warp * kRows; s0 < S; s0 += nwarps * kRows) {
        float acc[kRows][kFlatMaxGlimpses];
#pragma unroll
        for (int r = 0; r < kRows; ++r)
#pragma unroll
            for (int g = 0; g < kFlatMaxGlimpses; ++g) acc[r][g] = 0.f;
        for (int c = lane * 8; c < M; c += 256) {
            uint4 hv[kRows], lv[kRows];
#pragma unroll
            for (int r = 0; r < kRows; ++r) {
                const long long off = ((long long)b * S + s0 + r) * M + c;
                hv[r] = (s0 + r < S) ? *reinterpret_cast<const uint4*>(hmid + off) : make_uint4(0, 0, 0, 0);
                lv[r] = (hmid_lo != nullptr && s0 + r < S) ? *reinterpret_cast<const uint4*>(hmid_lo + off)
                                                           : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int g = 0; g < kFlatMaxGlimpses; ++g) {
                if (g < G) {
                    const float4 w0 = __ldg(reinterpret_cast<const float4*>(w2 + (long long)g * M + c));
                    const float4 w1 = __ldg(reinterpret_cast<const float4*>(w2 + (long long)g * M + c + 4));
#pragma unroll
                    for (int r = 0; r < kRows; ++r) {
                        acc[r][g] += bf16_lo_to_f(hv[r].x) * w0.x + bf16_hi_to_f(hv[r].x) * w0.y +
                                     bf16_lo_to_f(hv[r].y) * w0.z + bf16_hi_to_f(hv[r].y) * w0.w +
                                     bf16_lo_to_f(hv[r].z) * w1.x + bf16_hi_to_f(hv[r].z) * w1.y +
                                     bf16_lo_to_f(hv[r].w) * w1.z + bf16_hi_to_f(hv[r].w) * w1.w;
                        if (hmid_lo != nullptr)
                            acc[r][g] += bf16_lo_to_f(lv[r].x) * w0.x + bf16_hi_to_f(lv[r].x) * w0.y +
                                         bf16_lo_to_f(lv[r].y) * w0.z + bf16_hi_to_f(lv[r].y) * w0.w +
                                         bf16_lo_to_f(lv[r].z) * w1.x + bf16_hi_to_f(lv[r].z) * w1.y +
                                         bf16_lo_to_f(lv[r].w) * w1.z + bf16_hi_to_f(lv[r].w) * w1.w;
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
#pragma unroll
            for (int g = 0; g < kFlatMaxGlimpses; ++g) {
                if (g < G && s0 + r < S) {      // warp-uniform
                    const float v = warp_sum(acc[r][g]);
                    const bool masked = mask != nullptr && mask[(long long)b * S + s0 + r] != 0;
                    if (lane == 0) s_att[(s0 + r) * G + g] = masked ? -1e9f : v + b2[g];
                }
            }
        }
    }
    __syncthreads();
    // softmax over the sequence, one warp per glimpse
    for (int g = warp; g < G; g += nwarps) {
        float mx = -INFINITY;
        for (int s = lane; s < S; s += 32) mx = fmaxf(mx, s_att[s * G + g]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int s = lane; s < S; s += 32) {
            const float e = __expf(s_att[s * G + g] - mx);
            s_att[s * G + g] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        const float inv = 1.f / sum;
        for (int s = lane; s < S; s += 32) {
            const float pv = s_att[s * G + g] * inv;
            s_att[s * G + g] = pv;
            if (slice == 0) att_w[((long long)b * S + s) * G + g] = pv;
        }
    }
    __syncthreads();
    // pooled[g, cols of this slice] = sum_s att[s,g] x[s,:].  Threads = (column, row group): every row group
    // sums its share of the rows in a FIXED order, the groups are then added in a fixed order (bit-reproducible).
    const float4* xb = reinterpret_cast<const float4*>(x + (long long)b * S * H);
    const int hv = H >> 2;
    const int cols = hv / nslices;                  // float4 columns of this slice (host: hv % nslices == 0)
    const int c_first = slice * cols;
    const int nthreads = blockDim.x;
    const int width = cols < nthreads ? cols : nthreads;             // columns per pass
    const int ngroups = nthreads / width;                            // row groups working on one pass
    const int grp = threadIdx.x / width, cl = threadIdx.x - grp * width;
    for (int c0 = 0; c0 < cols; c0 += width) {
        const int c = c0 + cl;
        const bool valid = grp < ngroups && c < cols;
        for (int g0 = 0; g0 < G; g0 += 4) {  // 4 glimpses per pass over x
            float4 acc[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) {
#pragma unroll 4
                for (int s = grp; s < S; s += ngroups) {
                    const float4 xv = xb[(long long)s * hv + c_first + c];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (g0 + q < G) {
                            const float w = s_att[s * G + g0 + q];
                            acc[q].x += w * xv.x; acc[q].y += w * xv.y;
                            acc[q].z += w * xv.z; acc[q].w += w * xv.w;
                        }
                    }
                }
            }
            if (ngroups > 1) {
                __syncthreads();
#pragma unroll
                for (int q = 0; q < 4; ++q) s_red[q * nthreads + threadIdx.x] = acc[q];
                __syncthreads();
                if (valid && grp == 0) {
                    for (int o = 1; o < ngroups; ++o) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 v = s_red[q * nthreads + o * width + cl];
                            acc[q].x += v.x; acc[q].y += v.y; acc[q].z += v.z; acc[q].w += v.w;
                        }
                    }
                }
            }
            if (valid && grp == 0) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (g0 + q < G) {
                        const long long o = ((long long)b * G + g0 + q) * H + 4 * (c_first + c);
                        if (pooled32) *reinterpret_cast<float4*>(pooled32 + o) = acc[q];
                        if (pooledbf) {
                            uint2 w;
                            w.x = pack_bf16x2(acc[q].x, acc[q].y);
                            w.y = pack_bf16x2(acc[q].z, acc[q].w);
                            *reinterpret_cast<uint2*>(pooledbf + o) = w;
                        }
                    }
                }
            }
        }
    }
}

__global__ void __launch_bounds__(kFlatThreads)
attflat_pool_bwd_kernel(const float* __restrict__ dpooled, const float* __restrict__ pooled,
                        const bf16* __restrict__ hmid,
                        const float* __restrict__ w2, const uint8_t* __restrict__ mask,
                        const float* __restrict__ x, const float* __restrict__ att_w, int S, int H,
                        int M, int G, float gate_scale, float* __restrict__ dx,
                        bf16* __restrict__ dhmid, float* __restrict__ dw2,
                        float* __restrict__ db2) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float s_att[kFlatMaxSeq * kFlatMaxGlimpses];
    __shared__ float s_dl[kFlatMaxSeq * kFlatMaxGlimpses];  // d att_w, then d logit (rows of this CTA)
    __shared__ float s_dot[kFlatMaxGlimpses][kFlatThreads / 32];
    const int b = blockIdx.y, slice = blockIdx.x, nslices = gridDim.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int hv = H >> 2;
    const int rows_per = (S + nslices - 1) / nslices;
    const int s_lo = slice * rows_per, s_hi = min(S, s_lo + rows_per);

    for (int i = threadIdx.x; i < S * G; i += blockDim.x) s_att[i] = att_w[(long long)b * S * G + i];
    // dot[g] = sum_s att[s,g] datt[s,g] = dpooled[g,:] . pooled[g,:]
    const float4* dp = reinterpret_cast<const float4*>(dpooled + (long long)b * G * H);
    const float4* pp = reinterpret_cast<const float4*>(pooled + (long long)b * G * H);
    for (int g = 0; g < G; ++g) {
        float acc = 0.f;
        for (int c = threadIdx.x; c < hv; c += blockDim.x) {
            const float4 d = __ldg(dp + (long long)g * hv + c), q = __ldg(pp + (long long)g * hv + c);
            acc += (d.x * q.x + d.y * q.y) + (d.z * q.z + d.w * q.w);
        }
        acc = warp_sum(acc);
        if (lane == 0) s_dot[g][warp] = acc;
    }
    __syncthreads();

    // d att_w[s,g] = dpooled[g,:] . x[s,:] ;  dx[s,:] = sum_g att[s,g] dpooled[g,:]
    for (int s = s_lo + warp; s < s_hi; s += nwarps) {
        const float4* xr = reinterpret_cast<const float4*>(x + ((long long)b * S + s) * H);
        float4* dxr = reinterpret_cast<float4*>(dx + ((long long)b * S + s) * H);
        float dots[kFlatMaxGlimpses];
#pragma unroll
        for (int g = 0; g < kFlatMaxGlimpses; ++g) dots[g] = 0.f;
#pragma unroll 4
        for (int c = lane; c < hv; c += 32) {
            const float4 xv = xr[c];
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int g = 0; g < kFlatMaxGlimpses; ++g) {
                if (g < G) {
                    const float4 d = __ldg(dp + (long long)g * hv + c);
                    dots[g] += (d.x * xv.x + d.y * xv.y) + (d.z * xv.z + d.w * xv.w);
                    const float w = s_att[s * G + g];
                    o.x += w * d.x; o.y += w * d.y; o.z += w * d.z; o.w += w * d.w;
                }
            }
            dxr[c] = o;
        }
        const bool masked = mask != nullptr && mask[(long long)b * S + s] != 0;
#pragma unroll
        for (int g = 0; g < kFlatMaxGlimpses; ++g) {
            if (g < G) {
                const float v = warp_sum(dots[g]);
                float dot = 0.f;
                for (int w = 0; w < nwarps; ++w) dot += s_dot[g][w];
                // softmax backward over the sequence; masked positions get no gradient
                if (lane == 0) s_dl[s * G + g] = masked ? 0.f : s_att[s * G + g] * (v - dot);
            }
        }
    }
    __syncthreads();
    if (db2 != nullptr) {
        for (int g = warp; g < G; g += nwarps) {
            float bsum = 0.f;
            for (int s = s_lo + lane; s < s_hi; s += 32) bsum += s_dl[s * G + g];
            bsum = warp_sum(bsum);
            if (lane == 0) atomicAdd(db2 + g, bsum);
        }
    }
    // dhmid[s,m] = (hmid > 0) * gate_scale * sum_g dlogit[s,g] w2[g,m]
    const int mc = M >> 3;
    const int nrows = max(s_hi - s_lo, 0);
    for (int i = threadIdx.x; i < nrows * mc; i += blockDim.x) {
        const int s = s_lo + i / mc, c = (i % mc) * 8;
        const long long off = ((long long)b * S + s) * M + c;
        const uint4 hvv = *reinterpret_cast<const uint4*>(hmid + off);
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        for (int g = 0; g < G; ++g) {
            const float dl = s_dl[s * G + g];
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(w2 + (long long)g * M + c));
            const float4 w1 = __ldg(reinterpret_cast<const float4*>(w2 + (long long)g * M + c + 4));
            acc[0] += dl * w0.x; acc[1] += dl * w0.y; acc[2] += dl * w0.z; acc[3] += dl * w0.w;
            acc[4] += dl * w1.x; acc[5] += dl * w1.y; acc[6] += dl * w1.z; acc[7] += dl * w1.w;
        }
        const uint32_t hw[4] = {hvv.x, hvv.y, hvv.z, hvv.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            acc[2 * j] = bf16_lo_to_f(hw[j]) > 0.f ? acc[2 * j] * gate_scale : 0.f;
            acc[2 * j + 1] = bf16_hi_to_f(hw[j]) > 0.f ? acc[2 * j + 1] * gate_scale : 0.f;
        }
        uint4 o;
        o.x = pack_bf16x2(acc[0], acc[1]);
        o.y = pack_bf16x2(acc[2], acc[3]);
        o.z = pack_bf16x2(acc[4], acc[5]);
        o.w = pack_bf16x2(acc[6], acc[7]);
        *reinterpret_cast<uint4*>(dhmid + off) = o;
    }
    // dw2[g,m] += sum_s dlogit[s,g] hmid[s,m]
    if (dw2 != nullptr) {
        for (int m = threadIdx.x * 2; m < M; m += blockDim.x * 2) {
            float acc0[kFlatMaxGlimpses], acc1[kFlatMaxGlimpses];
#pragma unroll
            for (int g = 0; g < kFlatMaxGlimpses; ++g) { acc0[g] = 0.f; acc1[g] = 0.f; }
#pragma unroll 4
            for (int s = s_lo; s < s_hi; ++s) {
                const uint32_t hvv = *reinterpret_cast<const uint32_t*>(hmid + ((long long)b * S + s) * M + m);
                const float h0 = bf16_lo_to_f(hvv), h1 = bf16_hi_to_f(hvv);
#pragma unroll
                for (int g = 0; g < kFlatMaxGlimpses; ++g) {
                    if (g < G) {
                        const float dl = s_dl[s * G + g];
                        acc0[g] += dl * h0;
                        acc1[g] += dl * h1;
                    }
                }
            }
#pragma unroll
            for (int g = 0; g < kFlatMaxGlimpses; ++g) {
                if (g < G) {
                    atomicAdd(dw2 + (long long)g * M + m, acc0[g]);
                    atomicAdd(dw2 + (long long)g * M + m + 1, acc1[g]);
                }
            }
        }
    }
}

// debug / tuning switches (read once): block size and rows in flight per warp of the forward's logit phase
static int flat_env(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}
// Defaults from the sweep in profiles/r02_attflat_pool_sweep.txt (B200, batch 64, L2 flushed): both kernels are bound by
// the per-CTA chain of dependent phases, not by bandwidth; MORE slices than this made them slower (the forward's
// slices all recompute the logits), fewer left SMs idle.   forward 42 -> 24.6 us, backward 61 -> 41 us (image side).
static int flat_threads_fwd() { static int v = flat_env("MCAN_FLAT_THREADS", 512); return v; }
static int flat_threads_bwd() { static int v = flat_env("MCAN_FLAT_THREADS_BWD", 256); return v; }
static int flat_rows() { static int v = flat_env("MCAN_FLAT_ROWS", 4); return v; }
static int flat_max_slices_fwd() { static int v = flat_env("MCAN_FLAT_SLICES", 2); return v; }
static int flat_max_slices_bwd() { static int v = flat_env("MCAN_FLAT_SLICES_BWD", 4); return v; }

// CTAs per sample: enough to give every SM about two CTAs, a power of two <= 8 that divides `divisible`.
static int flat_slices(int batch, int divisible, int limit, int max_slices) {
    const int sms = device_num_sms();
    int c = 1;
    while (c < max_slices && batch * c < 2 * sms && divisible % (2 * c) == 0 && 2 * c <= limit) c *= 2;
    return c;
}

static int check_flat(int batch, int s, int h, int mlp, int g, const char* who) {
    MCAN_REQUIRE(batch >= 1 && s >= 1 && s <= kFlatMaxSeq, "%s: batch=%d s=%d (s<=128)", who, batch, s);
    MCAN_REQUIRE(g >= 1 && g <= kFlatMaxGlimpses, "%s: glimpses=%d (1..8)", who, g);
    MCAN_REQUIRE(h % 4 == 0 && mlp % 8 == 0, "%s: h=%d (%%4) mlp=%d (%%8)", who, h, mlp);
    return 0;
}

}  // namespace mcan

using namespace mcan;

extern "C" int mcan_attflat_pool_fwd(const void* hmid, const void* hmid_lo, const float* w2,
                                     const float* b2, const uint8_t* mask, const float* x,
                                     int32_t batch, int32_t s, int32_t h, int32_t mlp,
                                     int32_t glimpses, float* att_w, float* pooled_f32,
                                     void* pooled_bf16, void* stream) {
    MCAN_REQUIRE(hmid && w2 && b2 && x && att_w, "mcan_attflat_pool_fwd: null input");
    if (int rc = check_flat(batch, s, h, mlp, glimpses, "mcan_attflat_pool_fwd")) return rc;
    MCAN_REQUIRE((((uintptr_t)hmid | (uintptr_t)w2 | (uintptr_t)x | (uintptr_t)pooled_f32) & 15) == 0 &&
                     ((uintptr_t)pooled_bf16 & 7) == 0,
                 "mcan_attflat_pool_fwd: alignment");
    const int slices = flat_slices(batch, h / 4, 8, flat_max_slices_fwd());      // every slice owns (h/4)/slices float4 columns
#define FLAT_FWD(R)                                                                                              \
    MCAN_CHECK_CUDA(launch_kernel(attflat_pool_fwd_kernel<R>, dim3(slices, batch), dim3(flat_threads_fwd()), 0,      \
                                  reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<const bf16*>(hmid),   \
                                  reinterpret_cast<const bf16*>(hmid_lo), w2, b2, mask, x, s, h, mlp, glimpses,  \
                                  att_w, pooled_f32, reinterpret_cast<bf16*>(pooled_bf16)))
    if (flat_rows() >= 4) FLAT_FWD(4);
    else if (flat_rows() >= 2) FLAT_FWD(2);
    else FLAT_FWD(1);
#undef FLAT_FWD
    return 0;
}

extern "C" int mcan_attflat_pool_bwd(const float* dpooled, const float* pooled, const void* hmid, const float* w2,
                                     const uint8_t* mask, const float* x, const float* att_w,
                                     int32_t batch, int32_t s, int32_t h, int32_t mlp,
                                     int32_t glimpses, float gate_scale, float* dx, void* dhmid,
                                     float* dw2, float* db2, void* stream) {
    MCAN_REQUIRE(dpooled && pooled && hmid && w2 && x && att_w && dx && dhmid, "mcan_attflat_pool_bwd: null input");
    if (int rc = check_flat(batch, s, h, mlp, glimpses, "mcan_attflat_pool_bwd")) return rc;
    MCAN_REQUIRE((((uintptr_t)dpooled | (uintptr_t)pooled | (uintptr_t)hmid | (uintptr_t)w2 | (uintptr_t)x | (uintptr_t)dx |
                   (uintptr_t)dhmid) & 15) == 0,
                 "mcan_attflat_pool_bwd: alignment");
    const int slices = flat_slices(batch, 8, s, flat_max_slices_bwd());      // rows of a sample are split evenly, any power of two <= 8
    MCAN_CHECK_CUDA(launch_kernel(attflat_pool_bwd_kernel, dim3(slices, batch), dim3(flat_threads_bwd()), 0,
                                  reinterpret_cast<cudaStream_t>(stream), dpooled, pooled,
                                  reinterpret_cast<const bf16*>(hmid), w2, mask, x, att_w, s, h, mlp, glimpses,
                                  gate_scale, dx, reinterpret_cast<bf16*>(dhmid), dw2, db2));
    return 0;
}

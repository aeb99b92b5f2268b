// F1: AttFlat attention pooling (net.py:38-55), forward and backward, one CTA per sample.
//
// The H -> flat_mlp_size projection with ReLU/dropout (86 % of AttFlat's FLOPs) is a tcgen05
// GEMM (mcan_gemm, relu + dropout epilogue) that leaves hmid in bf16.  Everything after it --
// the flat_mlp_size -> glimpses projection, masked_fill(-1e9), the softmax over the SEQUENCE
// dimension and the glimpse-weighted sums -- is this one kernel: logits and attention weights
// stay in shared memory, x is streamed once with coalesced float4 loads.
#include "../../include/mcan_b200.h"
#include "common.cuh"

namespace mcan {

constexpr int kFlatThreads = 256;
constexpr int kFlatMaxSeq = 128;
constexpr int kFlatMaxGlimpses = 8;

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
    const int b = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;

    // logits[s,g] = hmid[s,:] . w2[g,:] + b2[g]; masked -> -1e9
    for (int s = warp; s < S; s += nwarps) {
        const bf16* hr = hmid + ((long long)b * S + s) * M;
        const bool masked = mask != nullptr && mask[(long long)b * S + s] != 0;
        for (int g = 0; g < G; ++g) {
            const float* wr = w2 + (long long)g * M;
            float acc = 0.f;
            for (int c = lane * 8; c < M; c += 256) {
                const uint4 hv = *reinterpret_cast<const uint4*>(hr + c);
                const float4 w0 = __ldg(reinterpret_cast<const float4*>(wr + c));
                const float4 w1 = __ldg(reinterpret_cast<const float4*>(wr + c + 4));
                acc += bf16_lo_to_f(hv.x) * w0.x + bf16_hi_to_f(hv.x) * w0.y +
                       bf16_lo_to_f(hv.y) * w0.z + bf16_hi_to_f(hv.y) * w0.w +
                       bf16_lo_to_f(hv.z) * w1.x + bf16_hi_to_f(hv.z) * w1.y +
                       bf16_lo_to_f(hv.w) * w1.z + bf16_hi_to_f(hv.w) * w1.w;
                if (hmid_lo != nullptr) {
                    const uint4 lv = *reinterpret_cast<const uint4*>(hmid_lo + ((long long)b * S + s) * M + c);
                    acc += bf16_lo_to_f(lv.x) * w0.x + bf16_hi_to_f(lv.x) * w0.y +
                           bf16_lo_to_f(lv.y) * w0.z + bf16_hi_to_f(lv.y) * w0.w +
                           bf16_lo_to_f(lv.z) * w1.x + bf16_hi_to_f(lv.z) * w1.y +
                           bf16_lo_to_f(lv.w) * w1.z + bf16_hi_to_f(lv.w) * w1.w;
                }
            }
            acc = warp_sum(acc);
            if (lane == 0) s_att[s * G + g] = masked ? -1e9f : acc + b2[g];
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
            att_w[((long long)b * S + s) * G + g] = pv;
        }
    }
    __syncthreads();
    // pooled[g,:] = sum_s att[s,g] x[s,:]
    const float4* xb = reinterpret_cast<const float4*>(x + (long long)b * S * H);
    const int hv = H >> 2;
    for (int c = threadIdx.x; c < hv; c += blockDim.x) {
        for (int g0 = 0; g0 < G; g0 += 4) {  // 4 glimpses per pass over x
            float4 acc[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int s = 0; s < S; ++s) {
                const float4 xv = xb[(long long)s * hv + c];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (g0 + q < G) {
                        const float w = s_att[s * G + g0 + q];
                        acc[q].x += w * xv.x; acc[q].y += w * xv.y;
                        acc[q].z += w * xv.z; acc[q].w += w * xv.w;
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (g0 + q < G) {
                    const long long o = ((long long)b * G + g0 + q) * H + 4 * c;
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

__global__ void __launch_bounds__(kFlatThreads)
attflat_pool_bwd_kernel(const float* __restrict__ dpooled, const bf16* __restrict__ hmid,
                        const float* __restrict__ w2, const uint8_t* __restrict__ mask,
                        const float* __restrict__ x, const float* __restrict__ att_w, int S, int H,
                        int M, int G, float gate_scale, float* __restrict__ dx,
                        bf16* __restrict__ dhmid, float* __restrict__ dw2,
                        float* __restrict__ db2) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float s_att[kFlatMaxSeq * kFlatMaxGlimpses];
    __shared__ float s_dl[kFlatMaxSeq * kFlatMaxGlimpses];  // d att_w, then d logit
    const int b = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int hv = H >> 2;

    for (int i = threadIdx.x; i < S * G; i += blockDim.x) s_att[i] = att_w[(long long)b * S * G + i];
    __syncthreads();

    // d att_w[s,g] = dpooled[g,:] . x[s,:] ;  dx[s,:] = sum_g att[s,g] dpooled[g,:]
    const float4* dp = reinterpret_cast<const float4*>(dpooled + (long long)b * G * H);
    for (int s = warp; s < S; s += nwarps) {
        const float4* xr = reinterpret_cast<const float4*>(x + ((long long)b * S + s) * H);
        float4* dxr = reinterpret_cast<float4*>(dx + ((long long)b * S + s) * H);
        float dots[kFlatMaxGlimpses];
#pragma unroll
        for (int g = 0; g < kFlatMaxGlimpses; ++g) dots[g] = 0.f;
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
#pragma unroll
        for (int g = 0; g < kFlatMaxGlimpses; ++g) {
            if (g < G) {
                const float v = warp_sum(dots[g]);
                if (lane == 0) s_dl[s * G + g] = v;
            }
        }
    }
    __syncthreads();
    // softmax backward over the sequence; masked positions get no gradient
    for (int g = warp; g < G; g += nwarps) {
        float dot = 0.f;
        for (int s = lane; s < S; s += 32) dot += s_att[s * G + g] * s_dl[s * G + g];
        dot = warp_sum(dot);
        float bsum = 0.f;
        for (int s = lane; s < S; s += 32) {
            const bool masked = mask != nullptr && mask[(long long)b * S + s] != 0;
            const float dl = masked ? 0.f : s_att[s * G + g] * (s_dl[s * G + g] - dot);
            s_dl[s * G + g] = dl;
            bsum += dl;
        }
        bsum = warp_sum(bsum);
        if (lane == 0 && db2 != nullptr) atomicAdd(db2 + g, bsum);
    }
    __syncthreads();
    // dhmid[s,m] = (hmid > 0) * gate_scale * sum_g dlogit[s,g] w2[g,m]
    const int mc = M >> 3;
    for (int i = threadIdx.x; i < S * mc; i += blockDim.x) {
        const int s = i / mc, c = (i % mc) * 8;
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
            for (int s = 0; s < S; ++s) {
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
    MCAN_CHECK_CUDA(launch_kernel(attflat_pool_fwd_kernel, dim3(batch), dim3(kFlatThreads), 0,
                                  reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<const bf16*>(hmid),
                                  reinterpret_cast<const bf16*>(hmid_lo), w2, b2, mask, x, s, h, mlp, glimpses,
                                  att_w, pooled_f32, reinterpret_cast<bf16*>(pooled_bf16)));
    return 0;
}

extern "C" int mcan_attflat_pool_bwd(const float* dpooled, const void* hmid, const float* w2,
                                     const uint8_t* mask, const float* x, const float* att_w,
                                     int32_t batch, int32_t s, int32_t h, int32_t mlp,
                                     int32_t glimpses, float gate_scale, float* dx, void* dhmid,
                                     float* dw2, float* db2, void* stream) {
    MCAN_REQUIRE(dpooled && hmid && w2 && x && att_w && dx && dhmid, "mcan_attflat_pool_bwd: null input");
    if (int rc = check_flat(batch, s, h, mlp, glimpses, "mcan_attflat_pool_bwd")) return rc;
    MCAN_REQUIRE((((uintptr_t)dpooled | (uintptr_t)hmid | (uintptr_t)w2 | (uintptr_t)x | (uintptr_t)dx |
                   (uintptr_t)dhmid) & 15) == 0,
                 "mcan_attflat_pool_bwd: alignment");
    MCAN_CHECK_CUDA(launch_kernel(attflat_pool_bwd_kernel, dim3(batch), dim3(kFlatThreads), 0,
                                  reinterpret_cast<cudaStream_t>(stream), dpooled,
                                  reinterpret_cast<const bf16*>(hmid), w2, mask, x, att_w, s, h, mlp, glimpses,
                                  gate_scale, dx, reinterpret_cast<bf16*>(dhmid), dw2, db2));
    return 0;
}

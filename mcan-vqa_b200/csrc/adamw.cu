// Fused multi-tensor AdamW for the fp32 master parameters (reference: core/model/optim.py:58-64,
// torch.optim.AdamW(lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4) driven by
// WarmupOptimizer.step, optim.py:26-34).
//
// ONE launch per optimiser step updates every parameter and, in the same pass, re-emits the bf16
// GEMM-operand copy of each weight (and the fp32 concatenation of stacked biases) that the next
// forward consumes -- the separate fp32 -> bf16 refresh pass over all weights disappears.
// HBM bound: 16 B read + 12 B (+2 B) written per parameter.  The segment table, the learning rate
// and the step counter live in device memory, so the launch is CUDA-graph replayable.
//
//   p   <- p * (1 - lr * wd)
//   m   <- b1 * m + (1 - b1) * g
//   v   <- b2 * v + (1 - b2) * g * g
//   p   <- p - lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
#include "../../include/mcan_b200.h"
#include "common.cuh"

namespace mcan {

int device_num_sms();

struct AdamSeg {
    float* p;
    const float* g;
    float* m;
    float* v;
    void* shadow;           // optional copy kept in sync with p: bf16, or fp32 when bit 62 of first_chunk is set
    long long n;            // elements
    long long first_chunk;  // index of this segment's first chunk (bits 0..60)
    void* shadow2;          // optional second copy (a weight can sit in two stacked operand buffers); fp32 flag = bit 61
};
constexpr long long kAdamChunk = 4096;
constexpr long long kAdamF32Flag = 1LL << 62;
constexpr long long kAdamF32Flag2 = 1LL << 61;
constexpr long long kAdamBf16Grad = 1LL << 60;    // g points to bf16 (the all-reduced bf16 staging buffer of dp.py)
constexpr long long kAdamChunkMask = ~(kAdamF32Flag | kAdamF32Flag2 | kAdamBf16Grad);

struct AdamCoef {
    float lr_wd;   // 1 - lr * wd
    float b1, b2, one_m_b1, one_m_b2;
    float step_size;       // lr / (1 - b1^t)
    float inv_sqrt_bc2;    // 1 / sqrt(1 - b2^t)
    float eps;
};

__device__ __forceinline__ float adam_one(float p, float g, float& m, float& v, const AdamCoef& c) {
    m = c.b1 * m + c.one_m_b1 * g;
    v = c.b2 * v + c.one_m_b2 * g * g;
    const float denom = sqrtf(v) * c.inv_sqrt_bc2 + c.eps;
    return p * c.lr_wd - c.step_size * (m / denom);
}

__global__ void __launch_bounds__(256)
adamw_multi_kernel(const AdamSeg* __restrict__ segs, int nseg, long long total_chunks,
                   const float* __restrict__ lr_dev, const float* __restrict__ step_dev, float beta1,
                   float beta2, float eps, float weight_decay) {
    pdl_launch_dependents();
    pdl_wait();
    AdamCoef c;
    {
        const float lr = __ldg(lr_dev), t = __ldg(step_dev);
        c.lr_wd = 1.f - lr * weight_decay;
        c.b1 = beta1;
        c.b2 = beta2;
        c.one_m_b1 = 1.f - beta1;
        c.one_m_b2 = 1.f - beta2;
        c.step_size = lr / (1.f - powf(beta1, t));
        c.inv_sqrt_bc2 = rsqrtf(1.f - powf(beta2, t));
        c.eps = eps;
    }
    for (long long chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
        int lo = 0, hi = nseg - 1;
        while (lo < hi) {   // last segment whose first_chunk <= chunk
            const int mid = (lo + hi + 1) >> 1;
            if ((segs[mid].first_chunk & kAdamChunkMask) <= chunk) lo = mid; else hi = mid - 1;
        }
        const AdamSeg sg = segs[lo];
        const bool shadow_f32 = (sg.first_chunk & kAdamF32Flag) != 0;
        const bool shadow2_f32 = (sg.first_chunk & kAdamF32Flag2) != 0;
        const bool g_bf16 = (sg.first_chunk & kAdamBf16Grad) != 0;
        const long long off = (chunk - (sg.first_chunk & kAdamChunkMask)) * kAdamChunk;
        const long long n = min(kAdamChunk, sg.n - off);
        float* p = sg.p + off;
        const float* g = g_bf16 ? nullptr : sg.g + off;
        const bf16* g16 = g_bf16 ? reinterpret_cast<const bf16*>(sg.g) + off : nullptr;
        float* m = sg.m + off;
        float* v = sg.v + off;
        float* sh32 = (sg.shadow != nullptr && shadow_f32) ? reinterpret_cast<float*>(sg.shadow) + off : nullptr;
        bf16* sh16 = (sg.shadow != nullptr && !shadow_f32) ? reinterpret_cast<bf16*>(sg.shadow) + off : nullptr;
        float* sh32b = (sg.shadow2 != nullptr && shadow2_f32) ? reinterpret_cast<float*>(sg.shadow2) + off : nullptr;
        bf16* sh16b = (sg.shadow2 != nullptr && !shadow2_f32) ? reinterpret_cast<bf16*>(sg.shadow2) + off : nullptr;
        const bool vec = ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v | (uintptr_t)sh32 | (uintptr_t)sh32b) & 15) == 0) &&
                         ((((uintptr_t)sh16 | (uintptr_t)sh16b | (uintptr_t)g16) & 7) == 0);
        long long done = 0;
        if (vec) {
            const long long nv = n >> 2;
            for (long long i = threadIdx.x; i < nv; i += blockDim.x) {
                float4 pv = reinterpret_cast<float4*>(p)[i];
                float4 gv;
                if (g_bf16) {
                    const uint2 gw = reinterpret_cast<const uint2*>(g16)[i];
                    gv = make_float4(bf16_lo_to_f(gw.x), bf16_hi_to_f(gw.x), bf16_lo_to_f(gw.y), bf16_hi_to_f(gw.y));
                } else {
                    gv = reinterpret_cast<const float4*>(g)[i];
                }
                float4 mv = reinterpret_cast<float4*>(m)[i];
                float4 vv = reinterpret_cast<float4*>(v)[i];
                pv.x = adam_one(pv.x, gv.x, mv.x, vv.x, c);
                pv.y = adam_one(pv.y, gv.y, mv.y, vv.y, c);
                pv.z = adam_one(pv.z, gv.z, mv.z, vv.z, c);
                pv.w = adam_one(pv.w, gv.w, mv.w, vv.w, c);
                reinterpret_cast<float4*>(p)[i] = pv;
                reinterpret_cast<float4*>(m)[i] = mv;
                reinterpret_cast<float4*>(v)[i] = vv;
                if (sh16 || sh16b) {
                    uint2 w;
                    w.x = pack_bf16x2(pv.x, pv.y);
                    w.y = pack_bf16x2(pv.z, pv.w);
                    if (sh16) reinterpret_cast<uint2*>(sh16)[i] = w;
                    if (sh16b) reinterpret_cast<uint2*>(sh16b)[i] = w;
                }
                if (sh32) reinterpret_cast<float4*>(sh32)[i] = pv;
                if (sh32b) reinterpret_cast<float4*>(sh32b)[i] = pv;
            }
            done = nv << 2;
        }
        for (long long i = done + threadIdx.x; i < n; i += blockDim.x) {
            float mm = m[i], vv = v[i];
            const float pn = adam_one(p[i], g_bf16 ? __bfloat162float(g16[i]) : g[i], mm, vv, c);
            p[i] = pn;
            m[i] = mm;
            v[i] = vv;
            if (sh16) sh16[i] = __float2bfloat16_rn(pn);
            if (sh32) sh32[i] = pn;
            if (sh16b) sh16b[i] = __float2bfloat16_rn(pn);
            if (sh32b) sh32b[i] = pn;
        }
    }
}

}  // namespace mcan

using namespace mcan;

extern "C" int mcan_adamw_multi(const void* seg_table_dev, int32_t num_segments, int64_t total_chunks,
                                const float* lr_dev, const float* step_dev, float beta1, float beta2,
                                float eps, float weight_decay, int32_t flags, void* stream) {
    MCAN_REQUIRE(seg_table_dev && lr_dev && step_dev && num_segments > 0 && total_chunks > 0,
                 "mcan_adamw_multi: bad args");
    MCAN_REQUIRE(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f,
                 "mcan_adamw_multi: betas (%f, %f) eps %g", beta1, beta2, eps);
    const int sms = device_num_sms();
    MCAN_REQUIRE(sms > 0, "mcan_adamw_multi: no CUDA device");
    // flags bit 0: one 4096-element chunk per CTA (short-lived CTAs) instead of a persistent grid -- for an update
    // that runs on a low-priority stream NEXT TO latency-bound kernels (the encoder backward): those get an SM
    // as soon as any of these CTAs retires, instead of waiting behind a grid that owns every SM to its end
    long long blocks = (flags & 1) ? total_chunks : (total_chunks < 16LL * sms ? total_chunks : 16LL * sms);
    MCAN_REQUIRE(blocks < (1LL << 31), "mcan_adamw_multi: too many chunks");
    MCAN_CHECK_CUDA(launch_kernel(adamw_multi_kernel, dim3((unsigned)blocks), dim3(256), 0,
                                  reinterpret_cast<cudaStream_t>(stream),
                                  reinterpret_cast<const AdamSeg*>(seg_table_dev), (int)num_segments,
                                  (long long)total_chunks, lr_dev, step_dev, beta1, beta2, eps, weight_decay));
    return 0;
}

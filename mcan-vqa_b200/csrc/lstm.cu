// Question encoder: embedding lookup + single-layer LSTM, forward and backward
// (reference core/model/net.py:66-78, 96-104: nn.Embedding -> nn.LSTM(batch_first, zero initial state), pads run
// through the LSTM; the token mask is make_mask(ques_ix) = (ques_ix == 0), net.py:99,135-137).
//
// The recurrence is 14 dependent steps of a [batch, H] x [H, 4H] product -- 0.5 GFLOP each, far too small for a
// GEMM launch per step (cuDNN: 29 TF32 GEMM launches + 45 element-wise launches per training step).  Here ONE
// persistent kernel runs all time steps:
//   * the hidden units are partitioned over the CTAs (U units = 4U gate rows of W_hh each); a CTA keeps ITS slice of
//     W_hh in REGISTERS for the whole sequence (mma.sync B fragments, 64 registers per thread for H = 1024) -- the
//     weight is read from HBM/L2 exactly once per launch instead of once per step;
//   * per step every CTA pulls h_{t-1} (bf16 [batch, H], 128 KB) from L2 into shared memory, the 8 warps split the
//     contraction (K) between them, partial sums meet in shared memory, then the gate non-linearities and the cell
//     update run on the CTA's own units (the cell state stays in registers across the steps);
//   * the steps are separated by a grid barrier (one atomic counter; the grid is at most 128 CTAs, one per SM).
// The input projection x W_ih^T for ALL time steps is one tcgen05 GEMM in front of the kernel, the weight gradients
// dW_ih / dW_hh are one grouped tcgen05 launch behind the backward kernel, which produces the pre-activation gradients
// dA (bf16) with the transposed recurrence dh_{t-1} = dA_t W_hh in the same persistent structure.
//
// Row layout of every per-token buffer: row(b, s) = b * (T + 1) + s with one spare slot per sample, so that
//   hbuf[b, s] = h_{s-1} (slot 0 = the zero initial state) is at the same row as dA[b, s = t] and x[b, s = t]:
// the GEMMs dW_hh = dA^T hbuf and dW_ih = dA^T x read plain row-major matrices with K = batch * (T + 1) rows
// (dA slot T and x slot T are zero).  bf16 operands, fp32 accumulation and fp32 cell state.
#include "../../include/mcan_b200.h"
#include "common.cuh"

namespace mcan {

int device_num_sms();

constexpr int kLstmThreads = 256;
constexpr int kLstmWarps = 8;
constexpr int kLstmMaxBatch = 64;     // rows of one launch (4 m16 tiles); larger batches are split by the host

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                          uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
        "{%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Grid barrier: every CTA adds 1, then waits until the counter reaches `target` (monotonic over the steps of one
// launch; the last CTA to leave the kernel resets it).  All CTAs are co-resident (grid <= #SMs, one CTA per SM).
__device__ __forceinline__ void lstm_grid_sync(uint32_t* bar, uint32_t target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1U);
        long long t0 = clock64();
        uint32_t spins = 0;
        while (ld_acquire_u32(bar) < target) {
            if ((++spins & 0x3FFU) == 0 && (clock64() - t0) > 8000000000LL) {
                printf("mcan: LSTM grid barrier timed out (block %d)\n", (int)blockIdx.x);
                __trap();
            }
        }
        __threadfence();
    }
    __syncthreads();
}
__device__ __forceinline__ void lstm_grid_exit(uint32_t* bar) {
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t n = atomicAdd(bar + 1, 1U);
        if (n == gridDim.x - 1) {      // every CTA has passed its last barrier: leave the counters ready for the next launch
            bar[0] = 0U;
            bar[1] = 0U;
            __threadfence();
        }
    }
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

struct LstmParams {
    const float* xw;        // [R, 4H] x W_ih^T + b_ih  (R = batch * (T + 1), row(b, s) = b * (T + 1) + s)
    const bf16* w_hh;       // [4H, H]
    const float* b_hh;      // [4H]
    int batch, steps, hidden;
    bf16* hbuf;             // [R, H]: slot s holds h_{s-1}; slot 0 is written with zeros by the forward kernel
    float* h_out;           // [batch * T, H] fp32, row b * T + t: the module output
    float* cbuf;            // [R, H] fp32 cell states c_t at slot t (backward)
    float* gates;           // [R, 4H] fp32 activated gates (i, f, g, o) at slot t (backward); may be NULL (inference)
    uint32_t* bar;          // 2 words, zero before the first launch
    // backward
    const float* dout;      // [batch * T, H]
    bf16* da;               // [R, 4H] bf16 pre-activation gradients at slot t; slot T is written with zeros
};

// ------------------------------------------------------------------------------------------------------------
// forward.  KT = k16 tiles per warp (H = 128 KT), NT = n8 tiles per CTA (U = 2 NT hidden units, 8 NT gate columns)
// ------------------------------------------------------------------------------------------------------------
template <int KT, int NT>
__global__ void __launch_bounds__(kLstmThreads, 1) lstm_fwd_kernel(const LstmParams p) {
    constexpr int H = KT * 128;
    constexpr int U = NT * 2;
    constexpr int NC = NT * 8;              // gate columns of this CTA: local column n = gate * U + u
    constexpr int HS = H + 8;               // shared row stride of the h tile (bf16): 16-byte aligned, conflict-free ldmatrix
    constexpr int RS = NC + 4;              // row stride of the partial-sum tiles (fp32)
    constexpr int PAIRS = (kLstmMaxBatch * U + kLstmThreads - 1) / kLstmThreads;
    extern __shared__ __align__(16) uint8_t lstm_smem[];
    bf16* hs = reinterpret_cast<bf16*>(lstm_smem);                                  // [64][HS]
    float* red = reinterpret_cast<float*>(lstm_smem + (size_t)kLstmMaxBatch * HS * 2);   // [8 warps][64][RS]
    pdl_launch_dependents();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int j = blockIdx.x;               // units [j * U, (j + 1) * U)
    const int B = p.batch, T = p.steps, S1 = T + 1;
    pdl_wait();

    // this CTA's slice of W_hh as mma B fragments, resident for the whole sequence
    uint32_t wreg[NT][KT][2];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const int n = nt * 8 + g;
        const long long row = (long long)(n / U) * H + j * U + (n % U);
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) {
            const int k0 = (warp * KT + kt) * 16 + 2 * t4;
            wreg[nt][kt][0] = *reinterpret_cast<const uint32_t*>(p.w_hh + row * H + k0);
            wreg[nt][kt][1] = *reinterpret_cast<const uint32_t*>(p.w_hh + row * H + k0 + 8);
        }
    }
    // rows >= batch of the h tile stay zero; slot 0 of hbuf = the zero initial state (this CTA's columns)
    for (int i = threadIdx.x; i < kLstmMaxBatch * (HS / 8); i += kLstmThreads)
        reinterpret_cast<uint4*>(hs)[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < B * U; i += kLstmThreads)
        p.hbuf[((long long)(i / U) * S1) * H + j * U + (i % U)] = __float2bfloat16_rn(0.f);
    float c_reg[PAIRS];
#pragma unroll
    for (int i = 0; i < PAIRS; ++i) c_reg[i] = 0.f;
    __syncthreads();

    for (int t = 0; t < T; ++t) {
        // input projection + recurrent bias of this thread's (sample, unit) pairs: fetched early, used after the MMAs
        float pre[PAIRS][4];
#pragma unroll
        for (int i = 0; i < PAIRS; ++i) {
            const int idx = threadIdx.x + i * kLstmThreads;
            const int b = idx / U, u = idx % U;
            if (idx < kLstmMaxBatch * U && b < B) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    pre[i][q] = p.xw[((long long)b * S1 + t) * (4 * H) + q * H + j * U + u] + __ldg(p.b_hh + q * H + j * U + u);
            }
        }
        if (t > 0) {
            // h_{t-1}: [batch, H] bf16 from L2 (written by all CTAs in the previous step) -> shared memory
            constexpr int CPR = H / 8;      // 16-byte chunks per row
            for (int i = threadIdx.x; i < B * CPR; i += kLstmThreads) {
                const int b = i / CPR, c8 = i % CPR;
                cp_async16(hs + b * HS + c8 * 8, p.hbuf + ((long long)b * S1 + t) * H + c8 * 8);
            }
            cp_async_commit();
            cp_async_wait<0>();
            __syncthreads();
            float acc[4][NT][4];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
#pragma unroll
            for (int kt = 0; kt < KT; ++kt) {
                const int k0 = (warp * KT + kt) * 16;
#pragma unroll
                for (int mt = 0; mt < 4; ++mt) {
                    uint32_t a0, a1, a2, a3;
                    ldsm_x4(smem_u32(hs + (mt * 16 + (lane & 15)) * HS + k0 + (lane >> 4) * 8), a0, a1, a2, a3);
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) mma_16816(acc[mt][nt], a0, a1, a2, a3, wreg[nt][kt][0], wreg[nt][kt][1]);
                }
            }
            float* rw = red + (size_t)warp * kLstmMaxBatch * RS;
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    *reinterpret_cast<float2*>(rw + (mt * 16 + g) * RS + nt * 8 + 2 * t4) = make_float2(acc[mt][nt][0], acc[mt][nt][1]);
                    *reinterpret_cast<float2*>(rw + (mt * 16 + g + 8) * RS + nt * 8 + 2 * t4) = make_float2(acc[mt][nt][2], acc[mt][nt][3]);
                }
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < PAIRS; ++i) {
            const int idx = threadIdx.x + i * kLstmThreads;
            const int b = idx / U, u = idx % U;
            if (idx < kLstmMaxBatch * U && b < B) {
                float a[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float s = pre[i][q];
                    if (t > 0) {
#pragma unroll
                        for (int w = 0; w < kLstmWarps; ++w) s += red[((size_t)w * kLstmMaxBatch + b) * RS + q * U + u];
                    }
                    a[q] = s;
                }
                const float ig = sigmoidf_(a[0]), fg = sigmoidf_(a[1]), gg = tanhf(a[2]), og = sigmoidf_(a[3]);
                const float c = fg * c_reg[i] + ig * gg;
                const float h = og * tanhf(c);
                c_reg[i] = c;
                const int col = j * U + u;
                const long long r = (long long)b * S1 + t;
                p.hbuf[(r + 1) * H + col] = __float2bfloat16_rn(h);
                p.h_out[((long long)b * T + t) * H + col] = h;
                if (p.gates != nullptr) {
                    p.cbuf[r * H + col] = c;
                    float* gp = p.gates + r * (4 * H) + col;
                    gp[0] = ig; gp[H] = fg; gp[2 * H] = gg; gp[3 * H] = og;
                }
            }
        }
        if (t + 1 < T) lstm_grid_sync(p.bar, (uint32_t)(t + 1) * gridDim.x);
    }
    lstm_grid_exit(p.bar);
}

// ------------------------------------------------------------------------------------------------------------
// backward: dA_t for all steps.  CTA j owns the 8 hidden units [8j, 8j + 8): per step it needs
//   dh_rec[b, u] = sum_r dA_{t+1}[b, r] W_hh[r, 8j + u]       (contraction over all 4H gate rows)
// so it streams dA_{t+1} ([batch, 4H] bf16 from L2) through a double-buffered shared-memory tile in chunks of 512
// columns, against ITS 8 columns of W_hh held in registers (KC chunks x 4 k16 tiles per warp).
// ------------------------------------------------------------------------------------------------------------
template <int KC>     // chunks of 512 gate rows: 4H = 512 KC
__global__ void __launch_bounds__(kLstmThreads, 1) lstm_bwd_kernel(const LstmParams p) {
    constexpr int H = KC * 128;
    constexpr int G4 = 4 * H;
    constexpr int U = 8;
    constexpr int CW = 512;                 // chunk width (gate rows = contraction columns)
    constexpr int CS = CW + 8;              // shared row stride (bf16)
    constexpr int RS = U + 4;
    constexpr int PAIRS = (kLstmMaxBatch * U) / kLstmThreads;      // 2
    extern __shared__ __align__(16) uint8_t lstm_smem[];
    bf16* ds = reinterpret_cast<bf16*>(lstm_smem);                                  // [2][64][CS]
    float* red = reinterpret_cast<float*>(lstm_smem + (size_t)2 * kLstmMaxBatch * CS * 2);   // [8][64][RS]
    pdl_launch_dependents();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int j = blockIdx.x;
    const int B = p.batch, T = p.steps, S1 = T + 1;
    pdl_wait();

    // B fragments of W_hh[:, 8j .. 8j+8) (k = gate row, n = unit): chunk c, k16 tile kk of this warp covers gate rows
    // c * 512 + warp * 64 + kk * 16 + ...
    uint32_t wreg[KC][4][2];
    {
        const unsigned short* w16 = reinterpret_cast<const unsigned short*>(p.w_hh);
        const int col = j * U + g;
#pragma unroll
        for (int c = 0; c < KC; ++c)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const long long k0 = (long long)c * CW + warp * 64 + kk * 16 + 2 * t4;
                const uint32_t lo0 = w16[k0 * H + col], hi0 = w16[(k0 + 1) * H + col];
                const uint32_t lo1 = w16[(k0 + 8) * H + col], hi1 = w16[(k0 + 9) * H + col];
                wreg[c][kk][0] = lo0 | (hi0 << 16);
                wreg[c][kk][1] = lo1 | (hi1 << 16);
            }
    }
    for (int i = threadIdx.x; i < 2 * kLstmMaxBatch * (CS / 8); i += kLstmThreads)
        reinterpret_cast<uint4*>(ds)[i] = make_uint4(0, 0, 0, 0);
    // slot T of dA (this CTA's 4 x 8 gate columns) is zero: the weight-gradient GEMMs contract over all T + 1 slots
    for (int i = threadIdx.x; i < B * 4 * U; i += kLstmThreads) {
        const int b = i / (4 * U), q = (i / U) % 4, u = i % U;
        p.da[((long long)b * S1 + T) * G4 + q * H + j * U + u] = __float2bfloat16_rn(0.f);
    }
    float dc_reg[PAIRS];
#pragma unroll
    for (int i = 0; i < PAIRS; ++i) dc_reg[i] = 0.f;
    __syncthreads();

    auto load_chunk = [&](int buf, int t_next, int c) {
        constexpr int CPR = CW / 8;
        bf16* dst = ds + (size_t)buf * kLstmMaxBatch * CS;
        for (int i = threadIdx.x; i < B * CPR; i += kLstmThreads) {
            const int b = i / CPR, c8 = i % CPR;
            cp_async16(dst + b * CS + c8 * 8, p.da + ((long long)b * S1 + t_next) * G4 + (long long)c * CW + c8 * 8);
        }
        cp_async_commit();
    };

    for (int t = T - 1; t >= 0; --t) {
        // operands of the element-wise part, fetched before the matrix product
        float dh[PAIRS], gi[PAIRS], gf[PAIRS], gg[PAIRS], go[PAIRS], cc[PAIRS], cp[PAIRS];
#pragma unroll
        for (int i = 0; i < PAIRS; ++i) {
            const int idx = threadIdx.x + i * kLstmThreads;
            const int b = idx / U, u = idx % U;
            if (b < B) {
                const int col = j * U + u;
                const long long r = (long long)b * S1 + t;
                dh[i] = p.dout[((long long)b * T + t) * H + col];
                const float* gp = p.gates + r * G4 + col;
                gi[i] = gp[0]; gf[i] = gp[H]; gg[i] = gp[2 * H]; go[i] = gp[3 * H];
                cc[i] = p.cbuf[r * H + col];
                cp[i] = t > 0 ? p.cbuf[(r - 1) * H + col] : 0.f;
            }
        }
        if (t < T - 1) {
            float acc[4][4];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[mt][e] = 0.f;
            load_chunk(0, t + 1, 0);
#pragma unroll
            for (int c = 0; c < KC; ++c) {
                if (c + 1 < KC) {
                    load_chunk((c + 1) & 1, t + 1, c + 1);
                    cp_async_wait<1>();
                } else {
                    cp_async_wait<0>();
                }
                __syncthreads();
                const bf16* src = ds + (size_t)(c & 1) * kLstmMaxBatch * CS;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const int k0 = warp * 64 + kk * 16;
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt) {
                        uint32_t a0, a1, a2, a3;
                        ldsm_x4(smem_u32(src + (mt * 16 + (lane & 15)) * CS + k0 + (lane >> 4) * 8), a0, a1, a2, a3);
                        mma_16816(acc[mt], a0, a1, a2, a3, wreg[c][kk][0], wreg[c][kk][1]);
                    }
                }
                __syncthreads();        // the buffer is refilled two chunks later
            }
            float* rw = red + (size_t)warp * kLstmMaxBatch * RS;
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) {
                *reinterpret_cast<float2*>(rw + (mt * 16 + g) * RS + 2 * t4) = make_float2(acc[mt][0], acc[mt][1]);
                *reinterpret_cast<float2*>(rw + (mt * 16 + g + 8) * RS + 2 * t4) = make_float2(acc[mt][2], acc[mt][3]);
            }
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < PAIRS; ++i) {
            const int idx = threadIdx.x + i * kLstmThreads;
            const int b = idx / U, u = idx % U;
            if (b < B) {
                float d = dh[i];
                if (t < T - 1) {
#pragma unroll
                    for (int w = 0; w < kLstmWarps; ++w) d += red[((size_t)w * kLstmMaxBatch + b) * RS + u];
                }
                const float tc = tanhf(cc[i]);
                const float dc = d * go[i] * (1.f - tc * tc) + dc_reg[i];
                const float dai = dc * gg[i] * gi[i] * (1.f - gi[i]);
                const float daf = dc * cp[i] * gf[i] * (1.f - gf[i]);
                const float dag = dc * gi[i] * (1.f - gg[i] * gg[i]);
                const float dao = d * tc * go[i] * (1.f - go[i]);
                dc_reg[i] = dc * gf[i];
                bf16* o = p.da + ((long long)b * S1 + t) * G4 + j * U + u;
                o[0] = __float2bfloat16_rn(dai);
                o[H] = __float2bfloat16_rn(daf);
                o[2 * H] = __float2bfloat16_rn(dag);
                o[3 * H] = __float2bfloat16_rn(dao);
            }
        }
        if (t > 0) lstm_grid_sync(p.bar, (uint32_t)(T - t) * gridDim.x);
    }
    lstm_grid_exit(p.bar);
}

// ------------------------------------------------------------------------------------------------------------
// embedding lookup (net.py:103) + token mask (net.py:99): x[row(b, t)] = bf16(table[token]), zero in slot T and in
// the pad columns; mask[b * T + t] = (token == 0).  One warp per output row.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
embed_gather_kernel(const long long* __restrict__ tokens, const float* __restrict__ table, int vocab, int E,
                    int batch, int steps, bf16* __restrict__ x, bf16* __restrict__ xlo, int ldx, uint8_t* __restrict__ mask) {
    pdl_launch_dependents();
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + warp;
    const int S1 = steps + 1;
    if (row >= batch * S1) return;
    const int b = row / S1, s = row % S1;
    bf16* xr = x + (long long)row * ldx;
    bf16* lr = xlo != nullptr ? xlo + (long long)row * ldx : nullptr;      // low-order halves: x = hi + lo to ~16 bits
    if (s == steps) {
        for (int c = lane; c < ldx; c += 32) {
            xr[c] = __float2bfloat16_rn(0.f);
            if (lr) lr[c] = __float2bfloat16_rn(0.f);
        }
        return;
    }
    long long tok = tokens[(long long)b * steps + s];
    if (lane == 0 && mask != nullptr) mask[b * steps + s] = tok == 0 ? 1 : 0;
    if (tok < 0 || tok >= vocab) tok = 0;       // (torch raises for an out-of-range index; never reached with valid data)
    const float* tr = table + tok * E;
    for (int c = lane; c < ldx; c += 32) {
        const float v = c < E ? tr[c] : 0.f;
        const bf16 h = __float2bfloat16_rn(v);
        xr[c] = h;
        if (lr) lr[c] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
}

// dTable[token[b, t], :] += dx[row(b, t), :E]   (fp32 atomics on a zero-initialised gradient; slot T is skipped)
__global__ void __launch_bounds__(256)
embed_scatter_kernel(const long long* __restrict__ tokens, const float* __restrict__ dx, int lddx, int vocab, int E,
                     int batch, int steps, float* __restrict__ dtable) {
    pdl_launch_dependents();
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + warp;
    if (i >= batch * steps) return;
    const int b = i / steps, t = i % steps;
    const long long tok = tokens[i];
    if (tok < 0 || tok >= vocab) return;
    const float* src = dx + ((long long)b * (steps + 1) + t) * lddx;
    float* dst = dtable + tok * E;
    for (int c = lane; c < E; c += 32) atomicAdd(dst + c, src[c]);
}

template <typename K>
static int lstm_launch(K kernel, int grid, size_t smem, const LstmParams& p, cudaStream_t st) {
    static size_t configured[64] = {0};
    int dev = 0;
    MCAN_CHECK_CUDA(cudaGetDevice(&dev));
    MCAN_REQUIRE(dev >= 0 && dev < 64, "device index %d", dev);
    // (one flag per device and kernel instance would be exact; setting the attribute again is harmless)
    MCAN_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    (void)configured;
    MCAN_CHECK_CUDA(launch_kernel(kernel, dim3(grid), dim3(kLstmThreads), smem, st, p));
    return 0;
}

static int lstm_check(const mcan_lstm_args* a, const char* who) {
    MCAN_REQUIRE(a != nullptr, "%s: null args", who);
    MCAN_REQUIRE(a->batch >= 1 && a->batch <= kLstmMaxBatch, "%s: batch=%d (1..64 rows per launch)", who, a->batch);
    MCAN_REQUIRE(a->steps >= 1 && a->steps <= 4096, "%s: steps=%d", who, a->steps);
    MCAN_REQUIRE(a->hidden == 128 || a->hidden == 256 || a->hidden == 512 || a->hidden == 1024,
                 "%s: hidden=%d (128, 256, 512 or 1024)", who, a->hidden);
    MCAN_REQUIRE(a->w_hh && a->hbuf && a->barrier, "%s: null pointer", who);
    MCAN_REQUIRE((a->gates == nullptr) == (a->cbuf == nullptr), "%s: gates and cbuf go together", who);
    MCAN_REQUIRE((((uintptr_t)a->w_hh | (uintptr_t)a->hbuf | (uintptr_t)a->da) & 15) == 0, "%s: alignment", who);
    return 0;
}

static void lstm_params(const mcan_lstm_args* a, LstmParams* p) {
    p->xw = a->xw;
    p->w_hh = reinterpret_cast<const bf16*>(a->w_hh);
    p->b_hh = a->b_hh;
    p->batch = a->batch;
    p->steps = a->steps;
    p->hidden = a->hidden;
    p->hbuf = reinterpret_cast<bf16*>(a->hbuf);
    p->h_out = a->h_out;
    p->cbuf = a->cbuf;
    p->gates = a->gates;
    p->bar = a->barrier;
    p->dout = a->dout;
    p->da = reinterpret_cast<bf16*>(a->da);
}

}  // namespace mcan

using namespace mcan;

extern "C" int mcan_lstm_fwd(const mcan_lstm_args* a) {
    if (int rc = lstm_check(a, "mcan_lstm_fwd")) return rc;
    MCAN_REQUIRE(a->xw && a->b_hh && a->h_out, "mcan_lstm_fwd: null pointer");
    MCAN_REQUIRE(device_num_sms() >= 128, "mcan_lstm_fwd: needs >= 128 SMs (the persistent grid must be co-resident)");
    LstmParams p;
    lstm_params(a, &p);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(a->stream);
    const int H = a->hidden;
#define LSTM_FWD(KT, NT)                                                                                         \
    do {                                                                                                         \
        const size_t smem = (size_t)kLstmMaxBatch * (KT * 128 + 8) * 2 + (size_t)kLstmWarps * kLstmMaxBatch * (NT * 8 + 4) * 4; \
        return lstm_launch(lstm_fwd_kernel<KT, NT>, H / (NT * 2), smem, p, st);                                  \
    } while (0)
    if (H == 1024) LSTM_FWD(8, 4);      // 128 CTAs x 8 units
    if (H == 512) LSTM_FWD(4, 2);       // 128 CTAs x 4 units
    if (H == 256) LSTM_FWD(2, 1);       // 128 CTAs x 2 units
    LSTM_FWD(1, 1);                     // H = 128: 64 CTAs x 2 units
#undef LSTM_FWD
}

extern "C" int mcan_lstm_bwd(const mcan_lstm_args* a) {
    if (int rc = lstm_check(a, "mcan_lstm_bwd")) return rc;
    MCAN_REQUIRE(a->dout && a->da && a->gates && a->cbuf, "mcan_lstm_bwd: null pointer");
    MCAN_REQUIRE(device_num_sms() >= 128, "mcan_lstm_bwd: needs >= 128 SMs (the persistent grid must be co-resident)");
    LstmParams p;
    lstm_params(a, &p);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(a->stream);
    const int H = a->hidden;
    const size_t smem = (size_t)2 * kLstmMaxBatch * (512 + 8) * 2 + (size_t)kLstmWarps * kLstmMaxBatch * (8 + 4) * 4;
    if (H == 1024) return lstm_launch(lstm_bwd_kernel<8>, H / 8, smem, p, st);
    if (H == 512) return lstm_launch(lstm_bwd_kernel<4>, H / 8, smem, p, st);
    if (H == 256) return lstm_launch(lstm_bwd_kernel<2>, H / 8, smem, p, st);
    return lstm_launch(lstm_bwd_kernel<1>, H / 8, smem, p, st);
}

extern "C" int mcan_embed_gather(const int64_t* tokens, const float* table, int32_t vocab, int32_t embed, int32_t batch,
                                 int32_t steps, void* x_bf16, void* x_bf16_lo, int32_t ldx, uint8_t* mask, void* stream) {
    MCAN_REQUIRE(tokens && table && x_bf16 && vocab > 0 && embed > 0 && batch > 0 && steps > 0 && ldx >= embed,
                 "mcan_embed_gather: bad args");
    const int rows = batch * (steps + 1);
    MCAN_CHECK_CUDA(launch_kernel(embed_gather_kernel, dim3((rows + 7) / 8), dim3(256), 0,
                                  reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<const long long*>(tokens), table,
                                  (int)vocab, (int)embed, (int)batch, (int)steps, reinterpret_cast<bf16*>(x_bf16),
                                  reinterpret_cast<bf16*>(x_bf16_lo), (int)ldx, mask));
    return 0;
}

extern "C" int mcan_embed_scatter_add(const int64_t* tokens, const float* dx, int32_t lddx, int32_t vocab, int32_t embed,
                                      int32_t batch, int32_t steps, float* dtable, void* stream) {
    MCAN_REQUIRE(tokens && dx && dtable && vocab > 0 && embed > 0 && batch > 0 && steps > 0 && lddx >= embed,
                 "mcan_embed_scatter_add: bad args");
    const int rows = batch * steps;
    MCAN_CHECK_CUDA(launch_kernel(embed_scatter_kernel, dim3((rows + 7) / 8), dim3(256), 0,
                                  reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<const long long*>(tokens), dx,
                                  (int)lddx, (int)vocab, (int)embed, (int)batch, (int)steps, dtable));
    return 0;
}

// Persistent LSTM recurrence for SLSTM (dac/model/encodec.py:272-288 -> nn.LSTM(C, C, 2)).
//
// The input projections x W_ih^T + b_ih + b_hh for all T steps are one GEMM (conv_simt.cu);
// what remains is the serial chain  gates_t = xg_t + h_{t-1} W_hh^T  for T steps.  One
// cooperative launch runs all T steps of one layer: G = H/U CTAs, CTA c owns hidden units
// [c*U, c*U+U) (all four gates), keeps their cell state in shared memory, and exchanges h_t
// through a [2][H][32] buffer in L2 with one device-wide barrier per step.
//
// Per step each CTA computes a [32 batch] x [4U gate rows] x [H] product on the FMA pipe:
// the 8 warps split H eight ways and stream their slice of W_hh (pre-packed per CTA as
// [H][4U], so the copy is linear) and of h_{t-1} through private cp.async double buffers;
// partial sums meet in shared memory, then 32*U threads apply the gate math
// (PyTorch gate order i, f, g, o).
#include <cooperative_groups.h>
#include "common.cuh"
#include "kernels.h"

namespace fac {

constexpr int LSTM_BT = 32;     // batch tile (columns of hT)
constexpr int LSTM_WARPS = 8;
constexpr int LSTM_KS = 16;     // k rows per cp.async sub-chunk

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <int U>
__global__ void __launch_bounds__(LSTM_WARPS * 32, 1) lstm_rec_kernel(LstmParams p) {
    constexpr int R = 4 * U;
    constexpr int RP = R + 1;   // padded row of the reduction buffer (bank-conflict-free reads)
    constexpr int STAGE_F = LSTM_KS * (LSTM_BT + R);  // floats per stage per warp
    extern __shared__ __align__(16) float smem[];
    float* stage_base = smem;                                   // [8 warps][2][STAGE_F]
    float* red = smem + LSTM_WARPS * 2 * STAGE_F;               // [8][32][R]
    float* cstate = red + LSTM_WARPS * LSTM_BT * RP;            // [32][U]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cta = blockIdx.x;
    const int j0 = cta * U;
    const int H = p.H;
    const int kslice = H / LSTM_WARPS;
    const int k_begin = warp * kslice;
    const int nsub = kslice / LSTM_KS;
    const int bg = lane >> 2, rg = lane & 3;   // 8 batch groups x 4 gates

    for (int i = tid; i < LSTM_BT * U; i += blockDim.x) cstate[i] = 0.f;
    __syncthreads();

    const float* wsrc = p.whh_p + (size_t)cta * H * R;
    float* my_stage = stage_base + warp * 2 * STAGE_F;

    constexpr int PAIRS = (LSTM_BT * U + LSTM_WARPS * 32 - 1) / (LSTM_WARPS * 32);
    for (int t = 0; t < p.T; ++t) {
        // ---- prefetch this step's input-projection gates (independent of the barrier) ----
        float xgv[PAIRS][4];
#pragma unroll
        for (int pi = 0; pi < PAIRS; ++pi) {
            int idx = tid + pi * LSTM_WARPS * 32;
            int b = idx % LSTM_BT, u = idx / LSTM_BT;
#pragma unroll
            for (int g = 0; g < 4; ++g)
                xgv[pi][g] = (idx < LSTM_BT * U && b < p.B)
                                 ? __ldg(p.xg + ((size_t)b * p.T + t) * (4 * H) + (size_t)g * H + j0 + u)
                                 : 0.f;
        }
        // ---- wait until every CTA has published h_{t-1} ----
        if (t > 0) {
            if (tid == 0) {
                unsigned target = (unsigned)p.G * (unsigned)t;
                unsigned v;
                do {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p.bar));
                } while (v < target);
            }
            __syncthreads();
        }
        const float* hprev = p.hT + (size_t)((t + 1) & 1) * H * LSTM_BT;  // parity of t-1

        float acc[4][U];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < U; ++j) acc[i][j] = 0.f;

        auto issue = [&](int sub, int buf) {
            float* hs = my_stage + buf * STAGE_F;
            float* wsm = hs + LSTM_KS * LSTM_BT;
            const float* hg = hprev + (size_t)(k_begin + sub * LSTM_KS) * LSTM_BT;
            const float* wg = wsrc + (size_t)(k_begin + sub * LSTM_KS) * R;
#pragma unroll
            for (int i = lane; i < LSTM_KS * LSTM_BT / 4; i += 32) cp_async16(hs + i * 4, hg + i * 4);
#pragma unroll
            for (int i = lane; i < LSTM_KS * R / 4; i += 32) cp_async16(wsm + i * 4, wg + i * 4);
            cp_async_commit();
        };

        issue(0, 0);
        for (int sub = 0; sub < nsub; ++sub) {
            int buf = sub & 1;
            if (sub + 1 < nsub) {
                issue(sub + 1, buf ^ 1);
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncwarp();
            const float* hs = my_stage + buf * STAGE_F;
            const float* wsm = hs + LSTM_KS * LSTM_BT;
#pragma unroll
            for (int k = 0; k < LSTM_KS; ++k) {
                float4 hv = *reinterpret_cast<const float4*>(hs + k * LSTM_BT + bg * 4);
                float hb[4] = {hv.x, hv.y, hv.z, hv.w};
                float wv[U];
#pragma unroll
                for (int j = 0; j < U; j += 4) {
                    float4 w4 = *reinterpret_cast<const float4*>(wsm + k * R + rg * U + j);
                    wv[j] = w4.x; wv[j + 1] = w4.y; wv[j + 2] = w4.z; wv[j + 3] = w4.w;
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < U; ++j) acc[i][j] = fmaf(hb[i], wv[j], acc[i][j]);
            }
            __syncwarp();   // everyone done with buf before it is refilled two iterations later
        }
        // ---- cross-warp reduction through shared memory ----
        float* myred = red + warp * LSTM_BT * RP;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < U; ++j) myred[(bg * 4 + i) * RP + rg * U + j] = acc[i][j];
        __syncthreads();

        float* hcur = p.hT + (size_t)(t & 1) * H * LSTM_BT;
#pragma unroll
        for (int pi = 0; pi < PAIRS; ++pi) {
            int idx = tid + pi * LSTM_WARPS * 32;
            if (idx >= LSTM_BT * U) break;
            int b = idx % LSTM_BT, u = idx / LSTM_BT;   // consecutive threads -> consecutive b (hT row)
            float g4[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < LSTM_WARPS; ++w) s += red[w * LSTM_BT * RP + b * RP + g * U + u];
                g4[g] = s + xgv[pi][g];
            }
            float ig = sigmoid_f(g4[0]), fg = sigmoid_f(g4[1]), gg = tanhf(g4[2]), og = sigmoid_f(g4[3]);
            float c = fg * cstate[b * U + u] + ig * gg;
            cstate[b * U + u] = c;
            float h = og * tanhf(c);
            hcur[(size_t)(j0 + u) * LSTM_BT + b] = h;
            if (b < p.B) {
                size_t o = ((size_t)b * p.T + t) * H + j0 + u;
                p.y[o] = p.skip ? h + p.skip[o] : h;
            }
        }
        // ---- publish ----
        __threadfence();
        __syncthreads();
        if (tid == 0) atomicAdd(p.bar, 1u);
    }
}

int lstm_units_per_cta(int H) {
    if (H % 12 == 0 && H / 12 <= 132 && (H / LSTM_WARPS) % LSTM_KS == 0) return 12;
    if (H % 8 == 0 && H / 8 <= 132 && (H / LSTM_WARPS) % LSTM_KS == 0) return 8;
    return 0;
}

template <int U>
static cudaError_t launch_u(const LstmParams& p, cudaStream_t st) {
    constexpr int R = 4 * U;
    size_t smem = sizeof(float) * (LSTM_WARPS * 2 * LSTM_KS * (LSTM_BT + R) + LSTM_WARPS * LSTM_BT * (R + 1) + LSTM_BT * U);
    cudaError_t e = cudaFuncSetAttribute(lstm_rec_kernel<U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(p.bar, 0, sizeof(unsigned int), st);
    if (e != cudaSuccess) return e;
    // h_{-1} = 0 lives in parity slot 1
    e = cudaMemsetAsync(p.hT + (size_t)p.H * LSTM_BT, 0, sizeof(float) * p.H * LSTM_BT, st);
    if (e != cudaSuccess) return e;
    LstmParams pp = p;
    void* args[] = {&pp};
    return cudaLaunchCooperativeKernel((void*)lstm_rec_kernel<U>, dim3(p.G), dim3(LSTM_WARPS * 32), args, smem, st);
}

cudaError_t launch_lstm_layer(const LstmParams& p, cudaStream_t st) {
    if (p.B > LSTM_BT || p.B <= 0) return cudaErrorInvalidValue;
    if (p.U == 8) return launch_u<8>(p, st);
    if (p.U == 12) return launch_u<12>(p, st);
    return cudaErrorInvalidValue;
}

}  // namespace fac

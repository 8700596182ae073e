// Alias-free activation: Activation1d.forward (alias_free_torch/act.py:24-29)
//   UpSample1d(2, 12)   resample.py:28-37 : replicate-pad 5 | depthwise conv_transpose1d stride 2
//                                            with the 12-tap Kaiser-sinc filter, x2 gain | crop 15/15
//   act                 SnakeBeta (modules/quantize.py:78-88) or identity
//   DownSample1d(2, 12) resample.py:54-57 / filter.py:88-96 : replicate-pad (5, 6) | depthwise
//                                            conv1d stride 2 with the same filter
// fused in one pass over [B, C, T] (the layout of its only callers, the predictor heads): the 2x
// oversampled signal never leaves shared memory.
#include "common.cuh"
#include "kernels.h"

namespace fac {

constexpr int AA_TILE = 256;

__global__ void __launch_bounds__(AA_TILE) alias_free_act_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                 int C, int T, const float* __restrict__ filt,
                                                                 const float* __restrict__ alpha,
                                                                 const float* __restrict__ beta) {
    __shared__ float f[12];
    __shared__ float xs[AA_TILE + 16];          // x[t0-8 .. t0+AA_TILE+8)
    __shared__ float us[2 * AA_TILE + 16];      // u[2 t0 - 5 .. 2 t0 + 2 AA_TILE + 6]
    const int ntile = (T + AA_TILE - 1) / AA_TILE;
    const long long blk = blockIdx.x;           // row-major over (b * C + c, time tile): no 65535 limit on B * C
    const long long row = blk / ntile;          // b * C + c
    const int c = (int)(row % C);
    const int t0 = (int)(blk - row * ntile) * AA_TILE;
    const float* xr = x + (size_t)row * T;
    if (threadIdx.x < 12) f[threadIdx.x] = filt[threadIdx.x];
    for (int i = threadIdx.x; i < AA_TILE + 16; i += AA_TILE) {
        int t = t0 - 8 + i;
        t = t < 0 ? 0 : (t > T - 1 ? T - 1 : t);       // replicate padding
        xs[i] = xr[t];
    }
    __syncthreads();
    const float al = alpha ? alpha[c] : 0.f;
    const float ib = beta ? 1.0f / (beta[c] + 1e-9f) : 0.f;
    const int U = 2 * T;
    for (int i = threadIdx.x; i < 2 * AA_TILE + 12; i += AA_TILE) {
        int m = 2 * t0 - 5 + i;
        m = m < 0 ? 0 : (m > U - 1 ? U - 1 : m);        // replicate padding of the upsampled signal
        // u[m] = 2 * sum_{k = (m+15) mod 2, step 2} xp[(m + 15 - k) / 2] f[k],  xp[i] = x[clamp(i - 5)]
        int n = m + 15;
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            int k = (n & 1) + 2 * j;
            int xi = (n - k) / 2 - 5;                   // index into x (before clamping)
            int sidx = xi - (t0 - 8);
            sidx = sidx < 0 ? 0 : (sidx > AA_TILE + 15 ? AA_TILE + 15 : sidx);
            acc = fmaf(xs[sidx], f[k], acc);
        }
        float u = 2.0f * acc;
        if (alpha) {
            float s = sinf(u * al);
            u = u + ib * (s * s);
        }
        us[i] = u;
    }
    __syncthreads();
    int t = t0 + threadIdx.x;
    if (t < T) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 12; ++k) acc = fmaf(us[2 * threadIdx.x + k], f[k], acc);
        y[(size_t)row * T + t] = acc;
    }
}

cudaError_t launch_alias_free_act(const float* x, float* y, int B, int C, int T, const float* filt12, const float* alpha,
                                  const float* beta, cudaStream_t st) {
    const long long nblk = (long long)((T + AA_TILE - 1) / AA_TILE) * B * C;
    if (nblk <= 0) return cudaSuccess;
    if (nblk > 0x7fffffffLL) return cudaErrorInvalidValue;
    alias_free_act_kernel<<<(unsigned)nblk, AA_TILE, 0, st>>>(x, y, C, T, filt12, alpha, beta);
    return cudaGetLastError();
}

}  // namespace fac

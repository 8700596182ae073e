// Shared device helpers for the FAcodec sm_100a hot path.
// Activations are CHANNELS-LAST fp32: a tensor the reference calls [B, C, T] lives in HBM as
// [B][T][C] ("frames x channels"); see DESIGN.md "Data layout in HBM".
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fac {

// ---- reflect / zero padding index map --------------------------------------------------
// Restates encodec.py:96-113 (pad1d) for one padded position p in [-pad_left, L + pad_right):
// reflect mode with the short-input branch (x zero-extended to length max_pad+1 before the
// reflection, then truncated).  Returns the source row, or -1 when the padded value is 0.
struct PadMap {
    int L;        // valid input length (rows)
    int Le;       // L, or max_pad + 1 when L <= max_pad (reflect mode)
    int reflect;  // 1 = reflect, 0 = zero pad
    __host__ __device__ static PadMap make(int L, int pad_left, int pad_right, int reflect) {
        PadMap m;
        m.L = L;
        m.reflect = reflect;
        int mp = pad_left > pad_right ? pad_left : pad_right;
        m.Le = (reflect && L <= mp) ? mp + 1 : L;
        return m;
    }
    __device__ __forceinline__ int src(int p) const {
        if (p >= 0 && p < L) return p;
        if (!reflect) return -1;
        int q = p < 0 ? -p : (p < Le ? p : 2 * Le - 2 - p);
        return (q >= 0 && q < L) ? q : -1;
    }
};

// ---- Snake activation, dac/nn/layers.py:17-24 -------------------------------------------
// x + (alpha + 1e-9)^-1 * sin(alpha x)^2 ; inv_alpha is precomputed on the host in fp32.
__device__ __forceinline__ float snake_f(float x, float alpha, float inv_alpha) {
    float s = sinf(alpha * x);
    return x + inv_alpha * (s * s);
}

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

// x * tanh(softplus(x)), modules/style_encoder.py:6-10 ; F.softplus threshold 20
__device__ __forceinline__ float mish_f(float x) {
    float sp = x > 20.0f ? x : log1pf(expf(x));
    return x * tanhf(sp);
}

enum OutAct { ACT_NONE = 0, ACT_TANH = 1, ACT_MISH = 2, ACT_SNAKE = 3 };

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace fac

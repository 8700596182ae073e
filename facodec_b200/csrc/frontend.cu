// Mel front-end tail: |STFT|^2 -> HTK mel filterbank -> log -> affine.
//
// Reference: FAquantizer.preprocess modules/quantize.py:239-242 over
// torchaudio.transforms.MelSpectrogram(sample_rate=24000, n_fft=2048, win_length=1200,
// hop_length=300, n_mels=80) (modules/quantize.py:228-230).
//
// The windowed DFT itself runs in conv_simt.cu as a K=1200-tap, stride-300, Cin=1 "conv" of the
// waveform against a [1200][2*1025] cos/-sin basis with the Hann window folded in (engine.cu
// builds the basis in fp64); reflect padding of the centre=True STFT is the conv's index map.
// This kernel consumes its output: spec[b][f][2*bin] = Re, [2*bin+1] = Im.
#include "common.cuh"
#include "kernels.h"

namespace fac {

constexpr int MEL_BINS = 1025;
constexpr int MEL_N = 80;

__global__ void __launch_bounds__(128) mel_from_spec_kernel(const float* __restrict__ spec, int ldspec,
                                                            const float* __restrict__ fb, float* __restrict__ mel,
                                                            int F, int Tm) {
    __shared__ float pw[MEL_BINS + 3];
    const int b = blockIdx.y, f = blockIdx.x;
    const float* row = spec + ((size_t)b * F + f) * ldspec;
    for (int i = threadIdx.x; i < MEL_BINS; i += blockDim.x) {
        float2 c = *reinterpret_cast<const float2*>(row + 2 * i);
        pw[i] = c.x * c.x + c.y * c.y;   // spec.abs().pow(2)
    }
    __syncthreads();
    const int m = threadIdx.x;
    if (m < MEL_N) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int i = 0;
        for (; i + 3 < MEL_BINS; i += 4) {
            a0 = fmaf(pw[i], __ldg(fb + (size_t)i * MEL_N + m), a0);
            a1 = fmaf(pw[i + 1], __ldg(fb + (size_t)(i + 1) * MEL_N + m), a1);
            a2 = fmaf(pw[i + 2], __ldg(fb + (size_t)(i + 2) * MEL_N + m), a2);
            a3 = fmaf(pw[i + 3], __ldg(fb + (size_t)(i + 3) * MEL_N + m), a3);
        }
        for (; i < MEL_BINS; ++i) a0 = fmaf(pw[i], __ldg(fb + (size_t)i * MEL_N + m), a0);
        float v = (a0 + a1) + (a2 + a3);
        // (log(1e-5 + mel) - mean) / std with mean=-4, std=4
        mel[((size_t)b * Tm + f) * MEL_N + m] = (logf(1e-5f + v) + 4.0f) / 4.0f;
    }
}

// Frame gather for the tensor-core DFT: frames[b][f][j] = wave[b][reflect(f*hop - pad + j)], j < win.  The hop (300) is
// not a multiple of the 16-channel chunk of the tcgen05 conv, so the strided "conv" view of the STFT cannot feed it
// directly; the explicit [B*F][1200] matrix (49 MB at B=32) can, as a plain K=1 GEMM against the folded basis.
__global__ void __launch_bounds__(256) stft_frames_kernel(const float* __restrict__ wave, float* __restrict__ frames, int T,
                                                          int F, int hop, int win, int pad) {
    const int b = blockIdx.y, f = blockIdx.x;
    const PadMap pm = PadMap::make(T, pad, pad, 1);
    const float* w = wave + (size_t)b * T;
    float* o = frames + ((size_t)b * F + f) * win;
    for (int j = threadIdx.x; j < win; j += blockDim.x) {
        const int src = pm.src(f * hop - pad + j);
        o[j] = src >= 0 ? __ldg(w + src) : 0.f;
    }
}
cudaError_t launch_stft_frames(const float* wave, float* frames, int B, int T, int F, int hop, int win, int pad, cudaStream_t st) {
    if (B <= 0 || F <= 0) return cudaSuccess;
    stft_frames_kernel<<<dim3(F, B), 256, 0, st>>>(wave, frames, T, F, hop, win, pad);
    return cudaGetLastError();
}

cudaError_t launch_mel_from_spec(const float* spec, int ldspec, const float* fb, float* mel, int B, int F, int Tm,
                                 cudaStream_t st) {
    if (B <= 0 || Tm <= 0) return cudaSuccess;
    dim3 grid(Tm, B);
    mel_from_spec_kernel<<<grid, 128, 0, st>>>(spec, ldspec, fb, mel, F, Tm);
    return cudaGetLastError();
}

}  // namespace fac

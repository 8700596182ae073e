/* facodec_b200 -- kernel-level test hooks, host-only packing/plan probes and per-launch profiling.
 *
 * NOT part of the drop-in surface (that is include/facodec_b200.h = SURVEY.md section 8b): these entry points exist
 * for tests/test_gpu_kernels.py, tests/test_host.py, scripts/ and bench.py's roofline pass.  Same library, same
 * conventions (plain pointers and sizes, negative fac_status on error).
 */
#ifndef FACODEC_B200_DEBUG_H
#define FACODEC_B200_DEBUG_H

#include "facodec_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Kernel-level test hooks (used by tests/test_gpu_kernels.py; not part of the drop-in surface).
 * fac_debug_conv runs the generic channels-last conv kernel on one layer: x [B,Tin,Cin] and
 * y [B,Tout,Cout] are DEVICE channels-last buffers, w_host is a HOST nn.Conv1d weight
 * [Cout,Cin,K] (already weight-normed), bias/in_alpha/out_alpha HOST vectors or NULL,
 * res a DEVICE tensor like y or NULL; act: 0 none, 1 tanh, 2 mish.
 * fac_debug_slstm runs SLSTM (2 layers + skip) on x [B,T,H] DEVICE with HOST nn.LSTM weights
 * w[8] = {w_ih_l0, w_hh_l0, b_ih_l0, b_hh_l0, w_ih_l1, ...}. */
int fac_debug_conv(fac_handle* h, const float* x, const float* w_host, const float* bias_host, int B, int Tin,
                   int Cin, int Cout, int K, int dil, int stride, int pad_left, int pad_right, int reflect,
                   const float* in_alpha_host, const float* out_alpha_host, int act, const float* res,
                   float* y, int Tout, void* stream);
/* Same contract as fac_debug_conv, forced through the tcgen05 kernels: promoted = 0 -> conv_tc_kernel
 * (3xTF32, accumulates in TMEM only), 1 -> conv_tcp_kernel (3xTF32, TMEM accumulators promoted to fp32
 * registers every ~48 MMAs; the variant used upstream of the VQ), 2 -> conv_tc_kernel with the bf16 hi/lo
 * split (decoder-only precision class), 3 -> conv_tcp_kernel with the fp16 hi + 2^11-scaled fp16 lo split
 * (experimental "encoder_f16x2" class).  Returns FAC_ERR_UNSUPPORTED when
 * the layer geometry is not eligible (Cin % 16, Cout % 16, stride). */
int fac_debug_conv_tc(fac_handle* h, const float* x, const float* w_host, const float* bias_host, int B, int Tin,
                      int Cin, int Cout, int K, int dil, int stride, int pad_left, int pad_right, int reflect,
                      const float* in_alpha_host, const float* out_alpha_host, int act, const float* res,
                      float* y, int Tout, int promoted, void* stream);
/* One ResidualUnit (dac/model/dac.py:25-42) y = x + conv1(snake(conv7_d(snake(x)))) on DEVICE channels-last
 * x, y [B,T,C] with HOST folded weights w7 [C,C,7], w1 [C,C,1].  mode 0: fp32 FMA kernels, 1: two tcgen05
 * launches, 2: the single fused tcgen05 launch (FAC_ERR_UNSUPPORTED if the geometry cannot be fused);
 * 3 / 4: as 1 / 2 with the bf16 hi/lo split. */
int fac_debug_resunit(fac_handle* h, const float* x, const float* w7_host, const float* b7_host, const float* w1_host,
                      const float* b1_host, const float* alpha1_host, const float* alpha2_host, int B, int T, int C,
                      int dil, int mode, float* y, void* stream);
/* clock64() phase timestamps written by one probe CTA of the most recent conv_tc_kernel launch:
 * [0] start, [1] all activation chunks produced, [2] GEMM 1 retired, [3] GEMM-2 operand produced (fused),
 * [4] GEMM 2 retired (fused), [5] epilogue done.  Kernel-tuning aid. */
int fac_debug_tc_phase_clocks(fac_handle* h, long long* out8);
/* Probe producer thread of the same launch, cycles summed over the tile's chunks: [0] waiting for a free operand buffer,
 * [1] waiting for the chunk's global loads, [2] Snake + split + stores + arrive, [3] number of chunks. */
int fac_debug_tc_producer_clocks(fac_handle* h, long long* out4);
/* Per-chunk timeline of the same probe CTA, out80 = [5][16] absolute clock64 values for chunks 0..15: [0] MMA warp saw the
 * chunk's operands, [1] MMA warp finished issuing the chunk, [2] producer thread 0 saw the operand buffer free,
 * [3] producer thread 0 arrived (chunk stored), [4] cycles the MMA warp waited for weights inside the chunk. */
int fac_debug_tc_trace(fac_handle* h, long long* out80);
/* Host-only: the recurrent-weight packing of lstm_rec_kernel for one nn.LSTM weight_hh [4H][H] (HOST, gate order
 * i,f,g,o): bf16 = 0 -> fp32 [G][H][4U] (row r = gate*U + u of the CTA owning hidden units g*U..g*U+U-1);
 * bf16 = 1 -> [G][H/16][hi|lo][8 k-pairs][4U] words of two bf16 (even k in the low half), lo = rn_bf16(w - hi).
 * bf16 = 2 / 3 -> the resident-W kernel's layouts (lstm2.cu): one fp16 plane [G][H/16][8][4U] / fp16 hi + 2^11-scaled lo planes
 * [G][H/16][hi|lo][8][4U], column of row r of k pair k2 = r ^ swizzle(k2) (conflict-free fragment loads).
 * Returns the number of 32-bit words (G*H*4U) also when out is NULL/too small; info3 = {U, G, 4U}. */
long long fac_debug_lstm_pack(const float* whh_host, int H, int bf16, float* out, long long capacity_floats, int* info3);
/* Host-only: the padding index map every conv kernel applies instead of materialising a padded copy
 * (dac/model/encodec.py:96-113 pad1d incl. the short-input branch): out[i] = source row of padded position
 * i - pad_left, or -1 where the padded value is zero; n must be pad_left + L + pad_right. */
int fac_debug_pad_map(int L, int pad_left, int pad_right, int reflect, int* out, int n);
/* Host-only (no GPU, no handle): the tile plan of the tcgen05 conv kernels for one layer geometry.  mode: 0 conv_tc TF32,
 * 1 conv_tcp (promoted) TF32, 2 conv_tc bf16, 3 conv_tcp fp16 hi + scaled lo, 4 fused ResidualUnit bf16, 5 fused TF32,
 * 6 conv_tt (transposed: out8[0] = output channels per CTA tile -- 128, or 256 in PAIR mode where two weight tiles share one
 * produced operand --, out8[1] = time steps per tile).
 * Tout may be 0 (unknown).  out8 = {N, MT, K chunks, weight-ring stages, TMEM columns, dynamic shared-memory bytes,
 * padded rows of the operand buffer, chunks per promotion}.  FAC_ERR_UNSUPPORTED when the layer is not eligible. */
int fac_debug_tc_plan(int Cin, int Cout, int K, int dil, int stride, int Tout, int mode, int occ2_maxn, int* out8);
/* Host-only: packs nn.Conv1d weights [Cout][Cin][K] (HOST) into the tensor-core blob of mode 0..3 (see
 * fac_debug_tc_plan): [Cout/N][K chunks][taps][hi|lo][k-groups][N][16 bytes], hi|lo = TF32 pair (4 k-groups of 4 fp32
 * words), bf16 pair or fp16 hi / 2^11-scaled lo (2 k-groups of 8 halves).  Returns the blob size in 32-bit words (also
 * when blob_out is NULL or too small), or a negative status. */
long long fac_debug_tc_pack(const float* w_host, int Cin, int Cout, int K, int stride, int mode, float* blob_out,
                            long long capacity_floats);
/* clock64() totals of CTA 0 of the most recent lstm_rec_kernel launch, summed over all steps:
 * [0] grid-barrier wait, [1] W_hh/h streaming + MMAs, [2] cross-warp reduce + gate math, [3] publish. */
int fac_debug_lstm_phase_clocks(fac_handle* h, long long* out4);
/* Host-only: see engine.cu -- the conv form of a ConvTranspose1d(k = 2s, stride s) weight [Cin][Cout][2s]: causal -> 2 taps,
 * non-causal -> 3 taps, out [taps][Cin][s*Cout] (phase-major channels); returns the float count. */
long long fac_debug_convtr_pack(const float* w_host, int Cin, int Cout, int stride, int causal, float* out,
                                long long capacity_floats);
int fac_debug_slstm(fac_handle* h, const float* x, const float* const* w_host, int B, int T, int H, float* y,
                    void* stream);
/* Registers (dst != NULL) or clears a named tap: the next forward copies that channels-last
 * intermediate into dst (DEVICE, up to capacity_floats).  Names: enc_conv0, enc_block1..4,
 * enc_lstm, mel80, f0_input, gamma_beta, dec_conv0, dec_lstm, dec_block1..4. */
int fac_debug_tap(fac_handle* h, const char* name, float* dst, size_t capacity_floats);

/* Per-kernel-family device timing for bench.py's roofline object: when enabled, every launch of
 * the forward paths is bracketed by CUDA events on the launching stream.  Families: "conv"
 * (conv_cl_kernel, fp32 FMA), "conv_tc" (conv_tc_kernel, tcgen05 3xTF32), "lstm_rec", "fa_quantize".  fac_profile_get returns
 * the accumulated device milliseconds, ALGORITHMIC flops (2*MACs) and bytes (in + out + weights
 * once) and launch count since the last fac_profile_reset (it synchronises the device). */
int fac_profile_enable(fac_handle* h, int on);
int fac_profile_reset(fac_handle* h);
int fac_profile_get(fac_handle* h, const char* family, double* ms, double* flops, double* bytes,
                    long long* launches);
/* Per-call-site breakdown as text lines "key<TAB>ms<TAB>GFLOP<TAB>GB<TAB>launches"; returns the
 * buffer size needed (call with buf = NULL first). */
size_t fac_profile_dump(fac_handle* h, char* buf, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* FACODEC_B200_DEBUG_H */

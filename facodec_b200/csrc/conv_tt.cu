// conv_tt_kernel: the TRANSPOSED tcgen05 formulation of the channels-last conv for everything upstream of the VQ.
//
// Same contract and the same fp32-faithful arithmetic class as conv_tcp_kernel<true> (fp16 hi + 2^11-scaled fp16 lo
// split, kind::f16, register-promoted accumulation), but the roles of the MMA operands are swapped:
//
//   D[co][t] = sum_tap sum_j  W[tap][j][co] * X[t - PLr + tap*dil][j]
//     A (M = 128 rows)  = weight tile  [128 output channels][16 k]   streamed by bulk TMA (zero rows beyond Cout)
//     B (N = NT <= 256) = activations  [NT time rows][16 k]          produced once per chunk, a tap = a row offset of
//                                                                    the descriptor start address (as before)
//     D                 = [128 TMEM lanes = output channels][NT columns = time steps]
//
// Why (measured on this chip with scripts/mma_probe.cu, profiles/r02/mma_probe_r02.log): one MMA stream pays a fixed
// ~110-130 cycles per tcgen05.mma whatever its size; a 128 x 64 x 16 MMA (the old C = 64 tile) keeps the tensor pipe 25 %
// busy, 128 x 128 45 %, 128 x 256 75 %.  With time as N every layer gets N = 256 regardless of its channel count, so
// the same product needs 4x (C = 64) / 2x (C = 128) fewer MMAs than with channels as N.
//
// TMEM: D0 (hi*hi, columns [0, NT)) is promoted into fp32 registers every <= 48 MMAs, D1 (the two 2^11-scaled cross
// terms, columns [256, 256 + NT)) lives for the whole tile and is added once, times 2^-11.  Both are single-buffered
// (2 x 256 columns is all of TMEM): the MMA warp idles while the accumulator warps drain a group (~1 k cycles per
// <= 24 k-cycle group).
// Epilogue: tcgen05.ld hands lane l of warp quarter q output channel co = 32 q + l and 16 consecutive time steps, so
// for a fixed time step the 32 lanes of a warp already cover 128 contiguous bytes of the channels-last row: results
// and residuals move fully coalesced without the shared-memory transpose the other kernels need; bias and Snake
// parameters are per-thread scalars.
//
// Roles (16 warps launched with 128 registers, re-allocated with setmaxnreg): warps 0-7 accumulators (192 registers: 128
// fp32 accumulators per thread + an epilogue that keeps its pointers in registers), warps 8-13 activation producers (64
// registers, fp16 split), warp 14 weight TMA, warp 15 MMA issue (the two single-warp critical roles get the highest warp
// ids: the scheduler prefers them).  Persistent: one CTA per SM walks the tile list,
// channel tile fastest so CTAs running at the same time share the activation rows in L2.
// With N = 256 the MMAs are cheap enough that the kernel is bound by the SIMT work around them (ncu, first version: 49
// thread-instructions per input element in the producers, 56 per output element in the epilogue, stall_no_inst and
// register-starved address re-materialisation on top): hence the interior-tile producer without any index map or
// bound checks, the 32-bit offset arithmetic and the one-range-check-per-slab Snake below.
#include <cstring>
#include <mutex>

#include "common.cuh"
#include "conv_tc_common.cuh"
#include "kernels.h"

namespace fac {

namespace tc {
constexpr int kThreadsT = 512;      // warps 0-7 accumulators (192 registers), 8-13 producers, 14 weight TMA, 15 MMA issue (64 registers)
constexpr int kProdT = 192;         // producer threads
constexpr int kStagesT = 4;        // weight ring depth (8 KB per stage)
constexpr int kABufT = 4;          // activation-operand ring depth (20 KB per buffer at NT = 256): lets the producers run ahead of
                                   // the MMAs on layers whose chunks are cheap for the tensor core (K = 1 / 2 taps)
constexpr int kPfT = 3;            // raw-activation staging ring depth (cp.async issued kPfT - 1 chunks ahead), 21 KB per stage
constexpr int kMT = 128;           // MMA M = output channels per tile
constexpr int kAccT = 128;         // fp32 register accumulators per accumulator thread (NT / 2 time steps)
struct SmemT {
    uint64_t b_full[kStagesT];
    uint64_t b_empty[kStagesT];
    uint64_t a_full[kABufT];
    uint64_t a_empty[kABufT];
    uint64_t acc_ready;
    uint64_t acc_free;
    uint32_t tmem_base;
    uint32_t pad;
};
static_assert(sizeof(SmemT) <= kSmemHdr, "SmemT header");
}  // namespace tc

// sin(x)^2 for the rare |x| > 4096 (beyond the two-constant Cody-Waite range of sin2_poly): reduction mod pi in double
// precision (error ~|x| * 2^-52), then the same polynomial.  A handful of instructions and no local memory, unlike the
// inlined sinf() slow path (Payne-Hanek) the older kernels carry: conv_tt's code must stay small (instruction cache).
__device__ __forceinline__ float sin2_wide(float x) {
    const double kd = rint((double)x * 0.31830988618379067154);
    const float r = (float)fma(-kd, 3.14159265358979323846, (double)x);
    const float r2 = r * r;
    float q = fmaf(r2, 1.60590438368216146e-10f, -2.50521083854417188e-8f);
    q = fmaf(q, r2, 2.75573192239858907e-6f);
    q = fmaf(q, r2, -1.98412698412698413e-4f);
    q = fmaf(q, r2, 8.33333333333333333e-3f);
    q = fmaf(q, r2, -1.66666666666666667e-1f);
    q = fmaf(q * r2, r, r);
    return q * q;
}
__device__ __forceinline__ float4 snake4_w(float4 x, float4 al, float4 ia) {
    const float y0 = al.x * x.x, y1 = al.y * x.y, y2 = al.z * x.z, y3 = al.w * x.w;
    const float m = fmaxf(fmaxf(fabsf(y0), fabsf(y1)), fmaxf(fabsf(y2), fabsf(y3)));
    float4 o;
    if (m > 4096.0f) {
        o.x = fmaf(ia.x, sin2_wide(y0), x.x); o.y = fmaf(ia.y, sin2_wide(y1), x.y);
        o.z = fmaf(ia.z, sin2_wide(y2), x.z); o.w = fmaf(ia.w, sin2_wide(y3), x.w);
    } else {
        o.x = fmaf(ia.x, sin2_poly(y0), x.x); o.y = fmaf(ia.y, sin2_poly(y1), x.y);
        o.z = fmaf(ia.z, sin2_poly(y2), x.z); o.w = fmaf(ia.w, sin2_poly(y3), x.w);
    }
    return o;
}
__device__ __forceinline__ void split_store_h(float4 x4, uint8_t* ahi, uint8_t* alo, uint32_t off) {
    __half2 h01 = __floats2half2_rn(x4.x, x4.y), h23 = __floats2half2_rn(x4.z, x4.w);
    float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    __half2 l01 = __floats2half2_rn((x4.x - f01.x) * tc::kLoScale, (x4.y - f01.y) * tc::kLoScale);
    __half2 l23 = __floats2half2_rn((x4.z - f23.x) * tc::kLoScale, (x4.w - f23.y) * tc::kLoScale);
    uint2 hv, lv;
    hv.x = *reinterpret_cast<uint32_t*>(&h01); hv.y = *reinterpret_cast<uint32_t*>(&h23);
    lv.x = *reinterpret_cast<uint32_t*>(&l01); lv.y = *reinterpret_cast<uint32_t*>(&l23);
    *reinterpret_cast<uint2*>(ahi + off) = hv;
    *reinterpret_cast<uint2*>(alo + off) = lv;
}

// Edge tiles (padding index map, partial tiles): the producer of conv_tc_common.cuh restated with the compact Snake.
template <int NTHR>
__device__ __forceinline__ void produce_chunk_edge(const TcConvParams& p, const PadMap& pm, const float* __restrict__ xb, int c,
                                                   int t0, int R, int Rpad, uint8_t* ahi, uint8_t* alo, int ptid) {
    constexpr int RSTEP = NTHR / 4;
    const int pc = ptid & 3;
    const int j = c * tc::kChunk + pc * 4;
    const int soff = j / p.Cin, ci = j - soff * p.Cin;
    const bool has_alpha = p.in_alpha != nullptr;
    float4 al = make_float4(0.f, 0.f, 0.f, 0.f), ia = al;
    if (has_alpha) {
        al = __ldg(reinterpret_cast<const float4*>(p.in_alpha + ci));
        ia = __ldg(reinterpret_cast<const float4*>(p.in_inv_alpha + ci));
    }
    const int row_limit = p.Tout + (p.Kr - 1) * p.dil;
    const int vrow0 = t0 - p.PLr;
    const float* __restrict__ xcol = xb + ci;
#pragma unroll 1
    for (int r = ptid >> 2; r < R; r += RSTEP * 2) {
        float4 v[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int rr = r + u * RSTEP, vrow = vrow0 + rr;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (rr < R && vrow < row_limit) {
                const int src = pm.src(vrow * p.vf + soff);
                if (src >= 0) v[u] = __ldg(reinterpret_cast<const float4*>(xcol + (size_t)src * p.ldx));
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int rr = r + u * RSTEP;
            if (rr < R) {
                float4 x4 = v[u];
                if (has_alpha) x4 = snake4_w(x4, al, ia);       // snake(0) == 0: padded zeros stay zero
                split_store_h(x4, ahi, alo, ((uint32_t)(pc >> 1) * Rpad + rr) * 16 + (uint32_t)(pc & 1) * 8);
            }
        }
    }
}

// Waiting without burning issue slots.  With N = 256 MMAs this kernel is bound by SIMT instruction issue (ncu: 2.8 warp
// instructions per cycle per SM), and back-to-back try_wait polling by the 6 producer and 8 accumulator warps was ~35 % of all
// issued instructions: every wait except the MMA warp's parks the thread with a suspend-time hint instead (NS nanoseconds;
// short, because a parked thread may only wake at the end of the hint).
template <int NS>
__device__ __forceinline__ void mbar_wait_park(uint64_t* bar, uint32_t parity) {
    long long t0 = 0;
    for (int spin = 0;; ++spin) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(tc::smem_u32(bar)), "r"(parity), "r"((uint32_t)NS) : "memory");
        if (ok) return;
        if (spin == 64) t0 = clock64();
        else if (spin > 64 && (spin & 1023) == 0 && clock64() - t0 > 4000000000LL) __trap();   // protocol bug: abort, do not hang
    }
}

// Interior tiles (every row the taps touch exists: no reflect / zero padding, no partial tile): thread `ptid` owns the
// 16-byte piece pc = ptid & 3 (4 channels) of rows ptid/4 + k*48 -- one pointer + a constant stride.  The raw fp32
// pieces travel global -> shared memory with cp.async into a thread-private slot of a kPfT-deep staging ring, issued
// kPfT - 1 chunks ahead (across tile boundaries), so the HBM latency of a chunk hides behind the transform of the
// previous ones without holding the loads in registers; the transform (Snake, fp16 hi / scaled-lo split) then reads
// its own pieces back.  Edge tiles take produce_chunk_edge (PadMap, bound checks, direct loads).
__device__ __forceinline__ void cp_async16_tt(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tc::smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit_tt() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_tt() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NTHR, int NB>
__device__ __forceinline__ void stage_chunk_interior(const TcConvParams& p, const float* __restrict__ xb, int c, int vrow0, int R,
                                                     uint8_t* stage /* [NB][NTHR][16 B] */, int ptid) {
    constexpr int RSTEP = NTHR / 4;
    const int pc = ptid & 3;
    const int j = c * tc::kChunk + pc * 4;
    const int soff = j / p.Cin, ci = j - soff * p.Cin;
    const int r0 = ptid >> 2;
    const float* __restrict__ src = xb + (size_t)((vrow0 + r0) * p.vf + soff) * p.ldx + ci;
    const size_t step = (size_t)RSTEP * p.vf * p.ldx;
#pragma unroll
    for (int u = 0; u < NB; ++u)
        if (r0 + u * RSTEP < R) cp_async16_tt(stage + ((size_t)u * NTHR + ptid) * 16, src + u * step);
}

template <int NTHR, int NB>
__device__ __forceinline__ void transform_chunk_interior(const TcConvParams& p, int c, int R, int Rpad, const uint8_t* stage,
                                                         uint8_t* ahi, uint8_t* alo, int ptid) {
    constexpr int RSTEP = NTHR / 4;
    const int pc = ptid & 3;
    const int r0 = ptid >> 2;
    const bool has_alpha = p.in_alpha != nullptr;
    const bool mufu = p.snake_mufu != 0;
    float4 al = make_float4(0.f, 0.f, 0.f, 0.f), ia = al;
    if (has_alpha) {
        const int j = c * tc::kChunk + pc * 4;
        const int ci = j - (j / p.Cin) * p.Cin;
        al = __ldg(reinterpret_cast<const float4*>(p.in_alpha + ci));
        ia = __ldg(reinterpret_cast<const float4*>(p.in_inv_alpha + ci));
    }
    const uint32_t off0 = ((uint32_t)(pc >> 1) * Rpad + r0) * 16 + (uint32_t)(pc & 1) * 8;
    const float4* __restrict__ st = reinterpret_cast<const float4*>(stage) + ptid;
#pragma unroll
    for (int u = 0; u < NB; ++u) {
        if (r0 + u * RSTEP < R) {
            float4 x4 = st[u * NTHR];
            if (has_alpha) x4 = mufu ? snake4_mufu(x4, al, ia) : snake4_w(x4, al, ia);
            split_store_h(x4, ahi, alo, off0 + u * (RSTEP * 16));
        }
    }
}

// One slab of 8 consecutive time steps of one output channel: bias, activation (ONE range check for the 8 Snake
// arguments), residual, store.  `off` = 32-bit element offset of the slab's first row from yb / rb.
template <int ACTT>
__device__ __forceinline__ void epilogue_slab8(float (&v)[8], float bi, float al, float ia, int act, float* __restrict__ yb,
                                               const float* __restrict__ rb, uint32_t off, uint32_t ldy, int nvalid, bool mufu) {
    float r[8];
    if (rb) {
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = (i < nvalid) ? rb[off + i * ldy] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += bi;
    if (ACTT == ACT_SNAKE || (ACTT < 0 && act == ACT_SNAKE)) {
        float y[8], m = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { y[i] = al * v[i]; m = fmaxf(m, fabsf(y[i])); }
        if (m > 4096.0f) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fmaf(ia, sin2_wide(y[i]), v[i]);
        } else if (mufu) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fmaf(ia, sin2_mufu(y[i]), v[i]);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fmaf(ia, sin2_poly(y[i]), v[i]);
        }
    } else if (ACTT < 0 && act == ACT_TANH) {
#pragma unroll 1
        for (int i = 0; i < 8; ++i) v[i] = tanhf(v[i]);
    } else if (ACTT < 0 && act == ACT_MISH) {
#pragma unroll 1
        for (int i = 0; i < 8; ++i) v[i] = mish_f(v[i]);
    }
    if (rb) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += r[i];
    }
    if (nvalid >= 8) {
#pragma unroll
        for (int i = 0; i < 8; ++i) yb[off + i * ldy] = v[i];
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (i < nvalid) yb[off + i * ldy] = v[i];
    }
}

// role wait-time probes (fac_set_option "tt_probe" 1 + fac_debug_tc_phase_clocks): CTA 3 accumulates clock64 spans --
// [0] CTA lifetime, [1] producers waiting for a free buffer, [2]/[3]/[4] MMA warp waiting for activations / weights / drained
// accumulators, [5] accumulator warps waiting for MMAs, [6] draining TMEM, [7] in the epilogue.
__device__ long long g_tt_probe[8];

template <bool PROBE, int ACTT>      // ACTT: ACT_NONE / ACT_SNAKE compiled in, -1 = the activation is read from p.out_act
__global__ void __launch_bounds__(tc::kThreadsT, 1) conv_tt_kernel(TcConvParams p) {
    using namespace tc;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    SmemT* sm = reinterpret_cast<SmemT*>(smem_raw);
    constexpr int KG = 2;                                   // 16-byte k-groups per 16-channel chunk (fp16)
    const int NT = p.NT;                                    // time steps per tile = MMA N
    const int R = NT + (p.Kr - 1) * p.dil;                  // activation rows all taps of a tile touch
    const int Rpad = p.Rpad;
    const uint32_t a_half = (uint32_t)Rpad * 16 * KG;       // one hi (or lo) activation buffer
    constexpr uint32_t w_half = (uint32_t)kMT * 16 * KG;    // one hi (or lo) weight tile: 4 KB
    uint8_t* a_base = smem_raw + kSmemHdr;                  // [kABufT bufs][hi|lo][KG][Rpad][16 B]
    uint8_t* w_base = a_base + (size_t)kABufT * 2 * a_half;                  // [S][hi|lo][KG][128][16 B]
    // PAIR mode (p.pair): a CTA tile is TWO 128-channel weight tiles x NT <= 128 time steps instead of one x 256: the produced
    // activation operand (the SIMT work that bounds every 1-3 tap layer: Snake + split of NT + halo rows per 16 channels) is
    // shared by both, i.e. produced Cout/256 times instead of Cout/128 times per time step; a weight-ring stage holds both tiles.
    const int CT = p.pair ? 2 : 1;
    const uint32_t w_stage = (uint32_t)CT * 2 * w_half;
    uint8_t* stg_base = w_base + (size_t)kStagesT * w_stage;   // [kPfT][7][192 threads][16 B] raw fp32 pieces (producers)
    constexpr int S = kStagesT;
    const int P = p.promote_every;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nchunk = p.nchunk, Kr = p.Kr;
    const int gx = (p.Tout + NT - 1) / NT, gy = ((p.Cout + kMT - 1) / kMT) / CT;   // gy: channel tiles (pairs of tiles in PAIR mode)
    const int ntiles = gx * gy * p.B;                       // L -> (channel tile, time tile, batch), channel tile fastest
    const int G = (nchunk + P - 1) / P;

    if (tid == 0) {
        for (int i = 0; i < S; ++i) { mbar_init(&sm->b_full[i], 1); mbar_init(&sm->b_empty[i], 1); }
        for (int i = 0; i < kABufT; ++i) { mbar_init(&sm->a_full[i], kProdT); mbar_init(&sm->a_empty[i], 1); }
        mbar_init(&sm->acc_ready, 1);
        mbar_init(&sm->acc_free, 256);
        fence_mbar_init();
    }
    // Role -> warp id.  The warp scheduler favours the HIGHEST warp id of a sub-partition among eligible warps (B300 microarch
    // notes): the MMA issuer and the weight-TMA issuer are the two latency-critical single warps, so they get the top ids
    // (15 and 14: sub-partitions 3 and 2); accumulators are warps 0-7 (TMEM lane quarter = warp & 3), producers 8-13.
    constexpr int kWarpTma = 14, kWarpMma = 15;
    if (warp == kWarpMma) tmem_alloc(&sm->tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm->tmem_base;

    if (warp >= 8) reg_dec<64>();     // 8*32*64 + 8*32*192 == 512*128
    if (warp == kWarpTma) {
        // ================= weight producer: one 8 KB bulk copy per (chunk, tap) =================
        if (lane == 0) {
            int it = 0;
            for (int L = blockIdx.x; L < ntiles; L += gridDim.x) {
                const int ntile = (L % gy) * CT;
                const size_t tile_floats = (size_t)nchunk * Kr * (size_t)(2 * w_half / 4);
                const float* wsrc = p.wblob + (size_t)ntile * tile_floats;
                for (int j = 0; j < nchunk * Kr; ++j, ++it) {
                    const int s = it % S;
                    mbar_wait_park<200>(&sm->b_empty[s], ((it / S) & 1) ^ 1);
                    mbar_arrive_expect_tx(&sm->b_full[s], w_stage);
                    for (int ct = 0; ct < CT; ++ct)
                        bulk_g2s(w_base + (size_t)s * w_stage + (size_t)ct * 2 * w_half, wsrc + (size_t)ct * tile_floats + (size_t)j * (2 * w_half / 4),
                                 2 * w_half, &sm->b_full[s]);
                }
            }
        }
    } else if (warp == kWarpMma) {
        // ================= MMA issuer (converged warp, elect-predicated tcgen05 instructions) =================
        // instruction descriptor: D = f32, A = B = f16, both K-major, N = NT, M = 128
        const uint32_t idesc = (1u << 4) | ((uint32_t)(NT >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
        const uint32_t a_base16 = __shfl_sync(0xffffffffu, smem_u32(a_base), 0) >> 4;
        const uint32_t w_base16 = __shfl_sync(0xffffffffu, smem_u32(w_base), 0) >> 4;
        const uint32_t a_lbo16 = (uint32_t)Rpad, w_lbo16 = (uint32_t)kMT;
        const uint32_t a_half16 = a_half >> 4, w_half16 = w_half >> 4;
        const uint32_t d0 = tmem_u, d1 = tmem_u + 256u;
        int it = 0, cg = 0, gg = 0;
        const bool mprobe = PROBE && blockIdx.x == 3;
        long long w_a = 0, w_b = 0, w_acc = 0, tq = 0;
        for (int L = blockIdx.x; L < ntiles; L += gridDim.x) {
            for (int g = 0; g < G; ++g, ++gg) {
                // D0 (and, for g == 0, D1 of the previous tile) drained by the accumulator warps
                if (mprobe) tq = clock64();
                mbar_wait(&sm->acc_free, (gg & 1) ^ 1);
                if (mprobe) w_acc += clock64() - tq;
                tc_fence_after();
                const int c_begin = g * P, c_end = (c_begin + P < nchunk) ? c_begin + P : nchunk;
                for (int c = c_begin; c < c_end; ++c, ++cg) {
                    const int buf = cg % kABufT;
                    if (mprobe) tq = clock64();
                    mbar_wait(&sm->a_full[buf], (cg / kABufT) & 1);
                    if (mprobe) w_a += clock64() - tq;
                    const uint32_t x_hi = a_base16 + (uint32_t)buf * 2 * a_half16;
                    const uint32_t x_lo = x_hi + a_half16;
                    for (int tap = 0; tap < Kr; ++tap, ++it) {
                        const int s = it % S;
                        if (mprobe) tq = clock64();
                        mbar_wait(&sm->b_full[s], (it / S) & 1);
                        if (mprobe) w_b += clock64() - tq;
                        tc_fence_after();
                        const uint32_t row_off = (uint32_t)(tap * p.dil);
                        const uint32_t first0 = ((c - c_begin) | tap) != 0, first1 = (g | (c - c_begin) | tap) != 0;
                        for (int ct = 0; ct < CT; ++ct) {
                            const uint32_t w_hi = w_base16 + (uint32_t)s * (w_stage >> 4) + (uint32_t)ct * 2 * w_half16;
                            const uint32_t w_lo = w_hi + w_half16;
                            const uint32_t dc = (uint32_t)ct * 128u;
                            umma_bf16(d0 + dc, desc_u(w_hi, w_lbo16), desc_u(x_hi + row_off, a_lbo16), idesc, first0);
                            umma_bf16(d1 + dc, desc_u(w_hi, w_lbo16), desc_u(x_lo + row_off, a_lbo16), idesc, first1);
                            umma_bf16(d1 + dc, desc_u(w_lo, w_lbo16), desc_u(x_hi + row_off, a_lbo16), idesc, 1u);
                        }
                        umma_commit(&sm->b_empty[s]);
                    }
                    umma_commit(&sm->a_empty[buf]);
                }
                umma_commit(&sm->acc_ready);
            }
        }
        if (mprobe && lane == 0) { g_tt_probe[2] = w_a; g_tt_probe[3] = w_b; g_tt_probe[4] = w_acc; }
    } else if (warp >= 8) {
        // ================= activation producers (warps 8..13) =================
        const int wtid = tid - 256;                                 // 0..191
        const PadMap pm = PadMap::make(p.Tin, p.pad_left_s, p.pad_right_s, p.reflect);
        const bool pprobe = PROBE && blockIdx.x == 3 && wtid == 0;
        const long long t_start = pprobe ? clock64() : 0;
        long long w_ae = 0, tq = 0;
        // two cursors over the CTA's (tile, chunk) stream: `pf` issues the cp.async of chunk n + kPfT - 1, `cs` transforms chunk n
        struct Cur { int L, c, t0, vrow0; bool interior; const float* xb; };
        auto setup = [&](Cur& k) {
            k.t0 = ((k.L / gy) % gx) * NT;
            k.xb = p.x + (size_t)(k.L / (gx * gy)) * p.x_bstride;
            k.vrow0 = k.t0 - p.PLr;
            // every row of the tile's union exists and the tile is full: no index map, no bound checks (R <= 7 * 48 rows)
            k.interior = k.vrow0 >= 0 && (k.vrow0 + R) * p.vf <= p.Tin && k.t0 + NT <= p.Tout && R <= 7 * (kProdT / 4);
        };
        auto advance = [&](Cur& k) {
            if (++k.c == nchunk) { k.c = 0; k.L += gridDim.x; if (k.L < ntiles) setup(k); }
        };
        Cur pf{(int)blockIdx.x, 0, 0, 0, false, nullptr}, cs = pf;
        if (pf.L < ntiles) { setup(pf); cs = pf; }
        constexpr uint32_t kStageBytes = 7u * kProdT * 16u;
#pragma unroll 1
        for (int i = 0; i < kPfT - 1; ++i) {
            if (pf.L < ntiles) {
                if (pf.interior) stage_chunk_interior<kProdT, 7>(p, pf.xb, pf.c, pf.vrow0, R, stg_base + (size_t)i * kStageBytes, wtid);
                advance(pf);
            }
            cp_async_commit_tt();
        }
        int cg = 0;
#pragma unroll 1
        for (; cs.L < ntiles; ++cg) {
            if (pf.L < ntiles) {
                if (pf.interior) stage_chunk_interior<kProdT, 7>(p, pf.xb, pf.c, pf.vrow0, R, stg_base + (size_t)((cg + kPfT - 1) % kPfT) * kStageBytes, wtid);
                advance(pf);
            }
            cp_async_commit_tt();
            cp_async_wait_tt<kPfT - 1>();                           // this thread's pieces of chunk cg have landed
            const int buf = cg % kABufT;
            uint8_t* ahi = a_base + (size_t)buf * 2 * a_half;
            if (pprobe) tq = clock64();
            mbar_wait_park<200>(&sm->a_empty[buf], ((cg / kABufT) & 1) ^ 1);
            if (pprobe) w_ae += clock64() - tq;
            if (cs.interior) transform_chunk_interior<kProdT, 7>(p, cs.c, R, Rpad, stg_base + (size_t)(cg % kPfT) * kStageBytes, ahi, ahi + a_half, wtid);
            else produce_chunk_edge<kProdT>(p, pm, cs.xb, cs.c, cs.t0, R, Rpad, ahi, ahi + a_half, wtid);
            fence_proxy_async();
            mbar_arrive(&sm->a_full[buf]);
            advance(cs);
        }
        cp_async_wait_tt<0>();
        if (pprobe) { g_tt_probe[0] = clock64() - t_start; g_tt_probe[1] = w_ae; }
    } else {
        // ================= accumulators (warps 0..7): promote + epilogue =================
        reg_inc<192>();
        const int q = warp & 3;                                     // TMEM lane quarter = output channels 32 q .. 32 q + 31
        const int half = warp >> 2;                                 // time half of the tile
        int split = ((NT / 2 + 15) / 16) * 16;
        if (split > NT) split = NT;
        // PAIR mode: warp (q, half) owns channel tile `half` of the pair and all NT <= 128 time steps (TMEM columns half * 128 ..)
        const int mycol0 = CT == 2 ? half * 128 : (half ? split : 0);        // first TMEM column of this warp
        const int mytime0 = CT == 2 ? 0 : mycol0;                            // first time step of this warp inside the tile
        const int mycols = CT == 2 ? NT : (half ? NT - split : split);      // <= 128
        const int act = p.out_act;
        int gg = 0;
        const bool aprobe = PROBE && blockIdx.x == 3 && tid == 0;
        long long w_ar = 0, t_dr = 0, t_ep = 0, tq = 0;
        for (int L = blockIdx.x; L < ntiles; L += gridDim.x) {
            const int ntile = (L % gy) * CT + (CT == 2 ? half : 0);
            const int t0 = ((L / gy) % gx) * NT;
            const int b = L / (gx * gy);
            const int co = ntile * kMT + q * 32 + lane;
            const bool co_ok = co < p.Cout;
            float acc[kAccT];
#pragma unroll
            for (int i = 0; i < kAccT; ++i) acc[i] = 0.f;
            for (int g = 0; g < G; ++g, ++gg) {
                if (aprobe) tq = clock64();
                mbar_wait_park<100>(&sm->acc_ready, gg & 1);
                if (aprobe) { long long n = clock64(); w_ar += n - tq; tq = n; }
                tc_fence_after();
                const uint32_t tb0 = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)mycol0;
#pragma unroll
                for (int grp = 0; grp < kAccT / 16; ++grp) {
                    if (grp * 16 < mycols) {
                        uint32_t v[16];
                        tmem_ld16(tb0 + grp * 16, v);
#pragma unroll
                        for (int i = 0; i < 16; ++i) acc[grp * 16 + i] += __uint_as_float(v[i]);
                    }
                }
                if (g == G - 1) {
                    // the last group's commit covers every MMA of the tile: add the scaled cross terms
#pragma unroll
                    for (int grp = 0; grp < kAccT / 16; ++grp) {
                        if (grp * 16 < mycols) {
                            uint32_t v[16];
                            tmem_ld16(tb0 + 256u + grp * 16, v);
#pragma unroll
                            for (int i = 0; i < 16; ++i) acc[grp * 16 + i] = fmaf(__uint_as_float(v[i]), kLoUnscale, acc[grp * 16 + i]);
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(&sm->acc_free);
                if (aprobe) t_dr += clock64() - tq;
            }
            if (aprobe) tq = clock64();
            // ---- epilogue from registers (TMEM already released: the next tile's MMAs run meanwhile).  8-step slabs,
            // rolled loop (the register slab is selected by a switch) to stay inside the instruction cache.
            float* __restrict__ yb = p.y + (size_t)b * p.y_bstride + co;
            const float* __restrict__ rb = p.res ? p.res + (size_t)b * p.y_bstride + co : nullptr;
            float bi = 0.f, al = 0.f, ia = 0.f;
            if (co_ok) {
                if (p.bias) bi = __ldg(p.bias + co);
                if (act == ACT_SNAKE) { al = __ldg(p.out_alpha + co); ia = __ldg(p.out_inv_alpha + co); }
            }
            const int tbeg = t0 + mytime0;
            const uint32_t ldy = (uint32_t)p.ldy;
#pragma unroll 1
            for (int sl = 0; sl * 8 < mycols; ++sl) {
                float v[8];
#define FAC_SLAB(S0) _Pragma("unroll") for (int i = 0; i < 8; ++i) v[i] = acc[(S0) + i];
                switch (sl) {
                    case 0: FAC_SLAB(0) break;
                    case 1: FAC_SLAB(8) break;
                    case 2: FAC_SLAB(16) break;
                    case 3: FAC_SLAB(24) break;
                    case 4: FAC_SLAB(32) break;
                    case 5: FAC_SLAB(40) break;
                    case 6: FAC_SLAB(48) break;
                    case 7: FAC_SLAB(56) break;
                    case 8: FAC_SLAB(64) break;
                    case 9: FAC_SLAB(72) break;
                    case 10: FAC_SLAB(80) break;
                    case 11: FAC_SLAB(88) break;
                    case 12: FAC_SLAB(96) break;
                    case 13: FAC_SLAB(104) break;
                    case 14: FAC_SLAB(112) break;
                    default: FAC_SLAB(120) break;
                }
#undef FAC_SLAB
                const int ts = tbeg + sl * 8;
                const int nvalid = co_ok ? p.Tout - ts : 0;          // rows of this slab that exist (may exceed 8)
                if (nvalid > 0) epilogue_slab8<ACTT>(v, bi, al, ia, act, yb, rb, (uint32_t)ts * ldy, ldy, nvalid, p.snake_mufu != 0);
            }
            if (aprobe) t_ep += clock64() - tq;
        }
        if (aprobe) { g_tt_probe[5] = w_ar; g_tt_probe[6] = t_dr; g_tt_probe[7] = t_ep; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kWarpMma) tmem_dealloc(tmem, 512);
}

// ---- host side ---------------------------------------------------------------------------------
int g_tt_pair_ok = 1;   // fac_set_option "tt_pair": 0 = never plan PAIR-mode tiles (A/B aid, process-wide)

bool tt_conv_plan(TcConvParams& p) {
    // p.Cin, p.vf, p.Kr, p.dil, p.Cout (and p.Tout when known) must be set
    if ((p.Cin % 4) != 0 || ((p.Cin * p.vf) % tc::kChunk) != 0 || p.Cout < 1) return false;
    p.tt = 1; p.promoted = 1; p.f16x2 = 1; p.bf16 = 0; p.fused = 0;
    p.N = tc::kMT; p.MT = 1;
    p.nchunk = p.Cin * p.vf / tc::kChunk;
    p.nchunk2 = 0;
    // time tile: 256 unless a smaller multiple of 16 wastes clearly fewer padded columns (T' = 320 layers -> 160)
    int NT = 256;
    if (p.Tout > 0) {
        auto padded = [&](int nt) { return (long long)((p.Tout + nt - 1) / nt) * nt; };
        long long best = padded(256);
        for (int cand = 240; cand >= 128; cand -= 16)
            if (padded(cand) * 10 < best * 9) { best = padded(cand); NT = cand; }
    }
    // PAIR mode for the producer-bound short-tap layers (1x1 convs, 2-tap down convs, input projections, k = 3 conv_out)
    // with an even number of 128-channel tiles: two weight tiles share one produced operand of NT <= 128 time steps
    const int ntile_co = (p.Cout + tc::kMT - 1) / tc::kMT;
    p.pair = (g_tt_pair_ok && p.Kr <= 3 && ntile_co >= 2 && ntile_co % 2 == 0) ? 1 : 0;
    if (p.pair) {
        NT = 128;
        if (p.Tout > 0) {
            auto padded = [&](int nt) { return (long long)((p.Tout + nt - 1) / nt) * nt; };
            long long best = padded(128);
            for (int cand = 112; cand >= 64; cand -= 16)
                if (padded(cand) * 10 < best * 9) { best = padded(cand); NT = cand; }
        }
    }
    p.NT = NT;
    p.promote_every = 48 / p.Kr < 1 ? 1 : 48 / p.Kr;
    int R = NT + (p.Kr - 1) * p.dil, Rpad = R;
    while (Rpad % 8 != 2) ++Rpad;
    p.Rpad = Rpad;
    p.tmem_cols = 512;
    p.stagesB = tc::kStagesT;
    const size_t a_bytes = (size_t)tc::kABufT * 2 * Rpad * 16 * 2, w_bytes = (size_t)(p.pair ? 2 : 1) * tc::kStagesT * 2 * tc::kMT * 16 * 2;
    p.smem_bytes = tc::kSmemHdr + a_bytes + w_bytes + (size_t)tc::kPfT * 7 * tc::kProdT * 16;
    return p.smem_bytes <= 225 * 1024;
}

size_t tt_blob_floats(const TcConvParams& p) {
    return (size_t)((p.Cout + tc::kMT - 1) / tc::kMT) * p.nchunk * p.Kr * 2 * 2 * tc::kMT * 4;
}

// wp: packed generic weights [Kr * vf*Cin][ldw].  blob: [co tile][chunk][tap][hi|lo'][k8 (2)][128 rows][8 fp16],
// lo' = rn_f16((w - hi) * 2^11), rows beyond Cout are zero.
void tt_pack_blob(const TcConvParams& p, const float* wp, int ldw, float* blob) {
    const int Cw = p.Cin * p.vf;
    uint16_t* ob = reinterpret_cast<uint16_t*>(blob);
    size_t o16 = 0;
    const int ntile = (p.Cout + tc::kMT - 1) / tc::kMT;
    for (int nt = 0; nt < ntile; ++nt)
        for (int c = 0; c < p.nchunk; ++c)
            for (int tap = 0; tap < p.Kr; ++tap)
                for (int hl = 0; hl < 2; ++hl)
                    for (int k8 = 0; k8 < 2; ++k8)
                        for (int n = 0; n < tc::kMT; ++n)
                            for (int e = 0; e < 8; ++e) {
                                const int co = nt * tc::kMT + n;
                                const int kk = tap * Cw + c * tc::kChunk + k8 * 8 + e;
                                const float w = co < p.Cout ? wp[(size_t)kk * ldw + co] : 0.f;
                                __half hi = __float2half_rn(w);
                                __half v = hl == 0 ? hi : __float2half_rn((w - __half2float(hi)) * 2048.0f);
                                uint16_t bits;
                                memcpy(&bits, &v, 2);
                                ob[o16++] = bits;
                            }
}

int g_tt_probe_on = 0;
cudaError_t tt_read_probe(long long* out8) { return cudaMemcpyFromSymbol(out8, g_tt_probe, sizeof(long long) * 8); }

namespace {
struct DevCfgT { bool done = false; int sm_count = 0; };
DevCfgT g_devcfg_t[64];
std::mutex g_devcfg_t_mu;
}  // namespace

cudaError_t launch_conv_tt(const TcConvParams& p, cudaStream_t st) {
    if (p.Tout <= 0 || p.B <= 0) return cudaSuccess;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    int sm_count;
    {
        std::lock_guard<std::mutex> lk(g_devcfg_t_mu);
        DevCfgT& d = g_devcfg_t[dev];
        if (!d.done) {
            const int cap = 225 * 1024;
            e = cudaFuncSetAttribute(conv_tt_kernel<false, ACT_NONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tt_kernel<false, ACT_SNAKE>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tt_kernel<false, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tt_kernel<true, ACT_NONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tt_kernel<true, ACT_SNAKE>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
            if (e != cudaSuccess) return e;
            if (cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || d.sm_count <= 0) d.sm_count = 148;
            d.done = true;
        }
        sm_count = d.sm_count;
    }
    const long long gx = (p.Tout + p.NT - 1) / p.NT, gy = ((p.Cout + tc::kMT - 1) / tc::kMT) / (p.pair ? 2 : 1);
    const long long ntiles = gx * gy * p.B;
    if (ntiles > 0x7fffffffLL) return cudaErrorInvalidValue;
    const unsigned nctas = (unsigned)(ntiles < sm_count ? ntiles : sm_count);
    if (g_tt_probe_on && p.out_act == ACT_NONE) conv_tt_kernel<true, ACT_NONE><<<dim3(nctas), tc::kThreadsT, p.smem_bytes, st>>>(p);
    else if (g_tt_probe_on && p.out_act == ACT_SNAKE) conv_tt_kernel<true, ACT_SNAKE><<<dim3(nctas), tc::kThreadsT, p.smem_bytes, st>>>(p);
    else if (p.out_act == ACT_NONE) conv_tt_kernel<false, ACT_NONE><<<dim3(nctas), tc::kThreadsT, p.smem_bytes, st>>>(p);
    else if (p.out_act == ACT_SNAKE) conv_tt_kernel<false, ACT_SNAKE><<<dim3(nctas), tc::kThreadsT, p.smem_bytes, st>>>(p);
    else conv_tt_kernel<false, -1><<<dim3(nctas), tc::kThreadsT, p.smem_bytes, st>>>(p);
    return cudaGetLastError();
}

}  // namespace fac

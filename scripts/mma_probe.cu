// tcgen05.mma issue-rate probe (kernel-tuning aid, not product code).
//
// Measures cycles per MMA for the operand layouts conv_tc.cu uses (K-major, no swizzle, [k-group][row][16 B]) under the
// variables that decide how the encoder convs should be restructured:
//   kind f16 (K = 16) vs tf32 (K = 8), N, number of TMEM accumulators rotated and how many consecutive MMAs hit one
//   accumulator, one or two issuing warps per CTA, one or two CTAs per SM, A from shared memory (SS) or TMEM (TS),
//   cta_group::1 vs cta_group::2.
// One config per process (a protocol bug traps that process only):
//   mma_probe <kind 0|1> <N> <nacc> <run_len> <nissue> <occ> <ts> <cg> [b_period] [reps]
// Prints one line: cycles per MMA (median / min / max over CTAs), the tensor-pipe floor N/2 (cta_group::1, M = 128) and
// the implied dense TFLOP/s at 1.9 GHz.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

struct Entry { uint32_t d_col, a, b, flags; };   // flags bit 0: accumulate, bit 1: commit to the dummy barrier after

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    long long t0 = clock64();
    while (!mbar_try(bar, parity)) {
        if (clock64() - t0 > 2000000000LL) __trap();
    }
}
__device__ __forceinline__ uint64_t desc_u(uint32_t lo) {
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(0x4008u));
    return d;
}

template <int KIND, bool TS, int CG>
__device__ __forceinline__ void umma(uint32_t d, uint32_t a, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    if constexpr (CG == 1) {
        if constexpr (!TS) {
            if constexpr (KIND == 0)
                asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
            else
                asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
        } else {
            if constexpr (KIND == 0)
                asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                             ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
            else
                asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                             ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
        }
    } else {
        if constexpr (KIND == 0)
            asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
        else
            asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "@q tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
    }
}
template <int CG>
__device__ __forceinline__ void commit(uint64_t* bar) {
    if constexpr (CG == 1)
        asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
                     "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
    else
        asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
                     "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}

struct Hdr { uint64_t done[2]; uint64_t dummy[2]; uint32_t tmem_base; uint32_t pad; };

template <int KIND, bool TS, int CG>
__global__ void __launch_bounds__(128, 1) probe_kernel(const Entry* __restrict__ prog, int P, int reps, int N, int tmem_cols,
                                                       int nissue, int op_bytes, long long* __restrict__ out) {
    extern __shared__ __align__(128) uint8_t smem[];
    Hdr* hd = reinterpret_cast<Hdr*>(smem);
    Entry* sprog = reinterpret_cast<Entry*>(smem + 256);                    // [nissue][P]
    uint8_t* ops = smem + 256 + ((size_t)2 * P * sizeof(Entry) + 127) / 128 * 128;
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t crank = 0;
    if constexpr (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    if (tid == 0) {
        mbar_init(&hd->done[0], 1); mbar_init(&hd->done[1], 1);
        mbar_init(&hd->dummy[0], 1); mbar_init(&hd->dummy[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < nissue * P; i += blockDim.x) sprog[i] = prog[i];
    for (int i = tid; i < op_bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(ops)[i] = 0x3C003C00u + (uint32_t)(i & 7);
    if (warp == 2) {
        if constexpr (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&hd->tmem_base)), "r"((uint32_t)tmem_cols));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&hd->tmem_base)), "r"((uint32_t)tmem_cols));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if constexpr (CG == 2) {
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = hd->tmem_base;
    const uint32_t fmt = KIND == 0 ? 0u : 2u;
    const uint32_t M = CG == 2 ? 256u : 128u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((M >> 4) << 24);
    if (warp < nissue && (CG == 1 || crank == 0)) {
        const uint32_t base16 = __shfl_sync(0xffffffffu, smem_u32(ops), 0) >> 4;
        const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
        const Entry* mp = sprog + warp * P;
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll 4
            for (int i = 0; i < P; ++i) {
                const Entry e = mp[i];
                umma<KIND, TS, CG>(tm + e.d_col, tm + e.a, desc_u(e.a + base16), desc_u(e.b + base16), idesc, e.flags & 1u);
                if (e.flags & 2u) commit<CG>(&hd->dummy[warp]);
            }
        }
        commit<CG>(&hd->done[warp]);
        mbar_wait(&hd->done[warp], 0);
        const long long t1 = clock64();
        if ((tid & 31) == 0) out[blockIdx.x * 2 + warp] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if constexpr (CG == 2) {
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    if (warp == 2) {
        if constexpr (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)tmem_cols));
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)tmem_cols));
    }
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

template <int KIND, bool TS, int CG>
int run(const std::vector<Entry>& prog, int P, int reps, int N, int tmem_cols, int nissue, int occ, int op_bytes, const char* label) {
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = sms * occ / CG * CG;
    Entry* dprog; long long* dout;
    CK(cudaMalloc(&dprog, prog.size() * sizeof(Entry)));
    CK(cudaMemcpy(dprog, prog.data(), prog.size() * sizeof(Entry), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dout, sizeof(long long) * grid * 2));
    CK(cudaMemset(dout, 0, sizeof(long long) * grid * 2));
    const size_t smem = 256 + ((size_t)2 * P * sizeof(Entry) + 127) / 128 * 128 + op_bytes;
    auto kern = probe_kernel<KIND, TS, CG>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    float best_ms = 1e30f;
    std::vector<long long> cyc(grid * 2);
    for (int it = 0; it < 3; ++it) {
        CK(cudaEventRecord(a));
        CK(cudaLaunchKernelEx(&cfg, kern, (const Entry*)dprog, P, reps, N, tmem_cols, nissue, op_bytes, dout));
        CK(cudaEventRecord(b));
        CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        best_ms = std::min(best_ms, ms);
    }
    CK(cudaMemcpy(cyc.data(), dout, sizeof(long long) * grid * 2, cudaMemcpyDeviceToHost));
    std::vector<double> per;
    for (int c = 0; c < grid; ++c)
        for (int w = 0; w < nissue; ++w)
            if (cyc[c * 2 + w] > 0) per.push_back((double)cyc[c * 2 + w] / ((double)P * reps));
    if (per.empty()) { printf("%s: no samples\n", label); return 3; }
    std::sort(per.begin(), per.end());
    const double med = per[per.size() / 2];
    const double Kel = KIND == 0 ? 16 : 8;
    const double M = CG == 2 ? 256 : 128;
    // MMAs retired per SM per cycle-of-one-stream: nissue streams per CTA, occ CTAs per SM (CG = 2: one stream per SM pair)
    const double flop_per_mma = 2.0 * M * N * Kel;
    const double streams = (double)nissue * occ;                 // per SM (cta_group::1) or per SM pair (cta_group::2)
    const double units = CG == 2 ? sms / 2.0 : (double)sms;
    const double tflops = flop_per_mma * streams / med * 1.9e9 * units / 1e12;
    const double floor_cyc = N / 2.0;                            // tensor-pipe cycles one MMA occupies (M = 128 per SM, K = 32 B)
    printf("%-46s cyc/MMA med %7.1f min %7.1f max %7.1f | floor %5.1f | pipe util %5.1f %% | ~%7.1f TF/s dense-eq @1.9GHz | %.3f ms\n",
           label, med, per.front(), per.back(), floor_cyc, 100.0 * floor_cyc * streams / med, tflops, best_ms);
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 9) { printf("usage: mma_probe kind N nacc run_len nissue occ ts cg [b_period] [reps]\n"); return 1; }
    const int kind = atoi(argv[1]), N = atoi(argv[2]), nacc = atoi(argv[3]), run_len = atoi(argv[4]), nissue = atoi(argv[5]);
    const int occ = atoi(argv[6]), ts = atoi(argv[7]), cg = atoi(argv[8]);
    const int b_period = argc > 9 ? atoi(argv[9]) : run_len;
    const int reps = argc > 10 ? atoi(argv[10]) : 16;
    const int P = 336;
    const int KG = kind == 0 ? 2 : 4;                    // 16-byte k-groups per 16-channel chunk
    const int nrows_b = cg == 2 ? N / 2 : N;             // B rows held by one CTA
    // operand area: A = 2 buffers x (hi|lo) x KG x Rpad x 16 B ; B = nb tiles x KG x N x 16 B
    const int Rpad = 128 * nacc + 58;
    const int a_buf16 = KG * Rpad;                       // one hi (or lo) buffer in 16-byte units
    const int nb = occ == 2 ? 4 : 8;
    const int b_tile16 = KG * nrows_b;
    const int a_total16 = 4 * a_buf16;
    int op_bytes = (a_total16 + nb * b_tile16) * 16;
    const int cap = occ == 2 ? 100 * 1024 : 200 * 1024;
    if (op_bytes > cap) { printf("operand area %d B exceeds %d\n", op_bytes, cap); return 1; }
    const int tmem_cols = occ == 2 ? 256 : 512;
    const int acc_cols_avail = ts ? tmem_cols / 2 : tmem_cols;
    if (nacc * N * nissue > acc_cols_avail) { printf("accumulators do not fit TMEM\n"); return 1; }
    std::vector<Entry> prog((size_t)2 * P);
    for (int w = 0; w < nissue; ++w)
        for (int i = 0; i < P; ++i) {
            Entry e;
            const int acc = (i / run_len) % nacc;
            const int tap = (i / (run_len * nacc)) % 7;
            const int pass = i % 3;
            const int abuf = ((i / 84) & 1) * 2 + (pass == 2 ? 1 : 0);
            const int ksub = kind == 0 ? 0 : (i & 1) * 2;                        // tf32: second K = 8 half of the chunk
            e.d_col = (uint32_t)((w * nacc + acc) * N);
            if (ts) e.a = (uint32_t)(acc_cols_avail + ((i * 8) % (tmem_cols - acc_cols_avail)));
            else e.a = (uint32_t)(abuf * a_buf16 + ksub * Rpad + acc * 128 + tap * 9) + ((uint32_t)Rpad << 16);
            const int bt = (i / b_period) % nb;
            e.b = (uint32_t)(a_total16 + bt * b_tile16 + ksub * nrows_b) + ((uint32_t)nrows_b << 16);
            e.flags = (i >= nacc * run_len ? 1u : 0u) | (((i + 1) % b_period == 0) ? 2u : 0u);
            prog[(size_t)w * P + i] = e;
        }
    char label[128];
    snprintf(label, sizeof label, "%s N=%d nacc=%d run=%d issuers=%d occ=%d %s cg=%d bper=%d", kind == 0 ? "f16 " : "tf32", N, nacc,
             run_len, nissue, occ, ts ? "TS" : "SS", cg, b_period);
#define DISPATCH(K, T, C) return run<K, T, C>(prog, P, reps, N, tmem_cols, nissue, occ, op_bytes, label)
    if (cg == 1) {
        if (kind == 0) { if (ts) DISPATCH(0, true, 1); else DISPATCH(0, false, 1); }
        else { if (ts) DISPATCH(1, true, 1); else DISPATCH(1, false, 1); }
    } else {
        if (ts) { printf("TS with cta_group::2 not probed\n"); return 1; }
        if (kind == 0) DISPATCH(0, false, 2); else DISPATCH(1, false, 2);
    }
    return 0;
}

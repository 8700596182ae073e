// Bulk-TMA weight-stream probe (kernel-tuning aid, not product code).
//
// Every CTA (one per SM) streams the SAME `region` bytes of global memory (L2-resident after the first pass) into a 4-slot
// shared-memory ring with cp.async.bulk, `slot` bytes per copy, `iters` copies, and reports bytes per clock per SM.  This
// is the access pattern of conv_tc_kernel's weight producer (every 128-row tile re-streams the layer's weights).
//   tma_probe <cluster 1|2|4> <slot_bytes> <region_bytes> <iters> [ctas]
// cluster > 1: each CTA of a cluster issues 1/cluster of every slot with .multicast::cluster to all CTAs of the cluster
// (the CUTLASS / DeepGEMM weight-sharing pattern), so L2 is read once per cluster instead of once per CTA.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    long long t0 = clock64();
    while (!mbar_try(bar, parity)) {
        if (clock64() - t0 > 2000000000LL) __trap();
    }
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}

constexpr int kSlots = 4;

__global__ void __launch_bounds__(128, 1) tma_probe_kernel(const uint8_t* __restrict__ src, uint32_t slot, uint32_t region, int iters,
                                                           int csz, long long* __restrict__ cycles) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);           // [kSlots]
    uint8_t* ring = smem + 128;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kSlots; ++i) mbar_init(&full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (csz > 1) cluster_sync_all();
    const uint32_t rank = csz > 1 ? cluster_ctarank() : 0;
    const uint32_t part = slot / (uint32_t)csz;
    const uint16_t mask = (uint16_t)((1u << csz) - 1u);
    long long t0 = 0;
    if (threadIdx.x == 0) {
        t0 = clock64();
        uint32_t off = 0;
        // keep kSlots - 1 copies in flight
        for (int it = 0; it < iters + kSlots - 1; ++it) {
            if (it < iters) {
                const int s = it % kSlots;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(slot) : "memory");
                const uint8_t* g = src + off + rank * part;
                uint8_t* d = ring + (size_t)s * slot + rank * part;
                if (csz > 1) {
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                                 ::"r"(smem_u32(d)), "l"(g), "r"(part), "r"(smem_u32(&full[s])), "h"(mask) : "memory");
                } else {
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(smem_u32(d)), "l"(g), "r"(slot), "r"(smem_u32(&full[s])) : "memory");
                }
                off += slot;
                if (off + slot > region) off = 0;
            }
            const int w = it - (kSlots - 1);
            if (w >= 0) mbar_wait(&full[w % kSlots], (w / kSlots) & 1);
        }
        cycles[blockIdx.x] = clock64() - t0;
    }
    __syncthreads();
    if (csz > 1) cluster_sync_all();
}

int main(int argc, char** argv) {
    if (argc < 5) { fprintf(stderr, "usage: tma_probe <cluster> <slot_bytes> <region_bytes> <iters> [ctas]\n"); return 2; }
    const int csz = atoi(argv[1]);
    const uint32_t slot = (uint32_t)atoi(argv[2]), region = (uint32_t)atoi(argv[3]);
    const int iters = atoi(argv[4]);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int ctas = argc > 5 ? atoi(argv[5]) : sms;
    ctas -= ctas % csz;
    uint8_t* src = nullptr;
    long long* cyc = nullptr;
    cudaMalloc(&src, region + slot);
    cudaMemset(src, 1, region + slot);
    cudaMalloc(&cyc, sizeof(long long) * ctas);
    const size_t smem = 128 + (size_t)kSlots * slot;
    cudaFuncSetAttribute(tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = csz; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    for (int rep = 0; rep < 3; ++rep) {
        cudaError_t e = cudaLaunchKernelEx(&cfg, tma_probe_kernel, (const uint8_t*)src, slot, region, iters, csz, cyc);
        if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return 1; }
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
    }
    std::vector<long long> h(ctas);
    cudaMemcpy(h.data(), cyc, sizeof(long long) * ctas, cudaMemcpyDeviceToHost);
    std::sort(h.begin(), h.end());
    const double bytes = (double)slot * iters;
    printf("cluster %d  slot %u B  region %u B  ctas %d: bytes/clk/SM median %.1f  (slowest CTA %.1f, fastest %.1f); chip %.0f B/clk\n", csz, slot,
           region, ctas, bytes / h[ctas / 2], bytes / h[ctas - 1], bytes / h[0], bytes * ctas / h[ctas - 1]);
    return 0;
}

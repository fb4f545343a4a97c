// Microbenchmark 3: throughput / latency of tcgen05.commit -> mbarrier arrive with NO MMAs in flight: (a) cta_group::1 on a
// local barrier, (b) cta_group::2 on the leader's barrier, (c) cta_group::2 multicast to both CTAs of the pair.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t ph) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }

template <int MODE>
__global__ void __launch_bounds__(128, 1) k(int n, int ring, unsigned long long *out) {
    __shared__ uint64_t bar[16], done[16];
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = MODE == 0 ? 0 : ctarank();
    if (threadIdx.x == 0) { for (int i = 0; i < 16; i++) { mbar_init(&bar[i], 1); mbar_init(&done[i], 1); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 1) {
        if (MODE == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&slot)) : "memory");
                         asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
        else { asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&slot)) : "memory");
               asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (MODE != 0) cluster_sync();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    unsigned long long t0 = clock64();
    if (warp == 0 && lane == 0 && rank == 0) {                   // committer (leader)
        for (int i = 0; i < n; i++) {
            uint64_t *b = &bar[i % ring];
            if (i >= ring) mbar_wait(&done[i % ring], (uint32_t)((i / ring - 1) & 1));      // the waiter has consumed the previous use
            if (MODE == 0) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(b)) : "memory");
            if (MODE == 1) asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(b)) : "memory");
            if (MODE == 2) asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(b)), "h"((uint16_t)3) : "memory");
        }
    }
    if (warp == 2 && lane == 0 && (MODE == 2 || rank == 0)) {     // waiter (both CTAs in multicast mode)
        for (int i = 0; i < n; i++) {
            mbar_wait(&bar[i % ring], (uint32_t)((i / ring) & 1));
            if (rank == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&done[i % ring])) : "memory");
        }
        out[blockIdx.x] = clock64() - t0;
    }
    __syncthreads();
    if (MODE != 0) cluster_sync();
    if (warp == 1) {
        if (MODE == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(slot) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 32;" ::"r"(slot) : "memory");
    }
}

template <int MODE> void run(const char *name, int grid) {
    unsigned long long *d; cudaMalloc(&d, 8 * 296); cudaMemset(d, 0, 8 * 296);
    for (int ring : {1, 4}) {
        const int n = 4000;
        cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.stream = 0;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = MODE == 0 ? 1 : 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        for (int rep = 0; rep < 2; rep++) cudaLaunchKernelEx(&cfg, k<MODE>, n, ring, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: error %s\n", name, cudaGetErrorString(e)); return; }
        unsigned long long h[296]; cudaMemcpy(h, d, 8 * grid, cudaMemcpyDeviceToHost);
        unsigned long long mx = 0; for (int i = 0; i < grid; i++) mx = h[i] > mx ? h[i] : mx;
        printf("%-40s ring %d: %.0f cycles per commit->arrive (%s)\n", name, ring, (double)mx / n, ring == 1 ? "latency, one in flight" : "throughput, 4 in flight");
    }
}

int main() {
    run<0>("cta_group::1 local", 148);
    run<1>("cta_group::2 leader barrier", 148);
    run<2>("cta_group::2 multicast to both CTAs", 148);
    return 0;
}

// Microbenchmark: how many bytes per cycle can ONE SM ingest through TMA (cp.async.bulk.tensor, SWIZZLE_128B boxes of
// rows x 128 bytes) from an L2-resident / HBM-resident buffer, as a function of ring depth, box rows, how many SMs
// pull at once and whether they pull the same or different addresses.  Build: nvcc -arch=sm_100a -O3 -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t ph) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap *m, uint64_t *bar, void *dst, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

struct Args { int stages, box_rows, n_boxes, rows_total, cols_total, same, per_stage; };

__global__ void __launch_bounds__(128, 1) ingest(const __grid_constant__ CUtensorMap map, Args a, unsigned long long *cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
    __shared__ uint64_t full[16], empty[16];
    const int stage_bytes = a.box_rows * 128 * a.per_stage;
    if (threadIdx.x == 0) {
        for (int i = 0; i < a.stages; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map) : "memory");
    }
    __syncthreads();
    const int n_row_boxes = a.rows_total / a.box_rows, n_col_boxes = a.cols_total / 64;
    unsigned long long t0 = 0;
    if (threadIdx.x == 0) {                    // producer
        t0 = clock64();
        int s = 0; uint32_t ph = 0;
        int idx = a.same ? 0 : blockIdx.x * 977;
        for (int i = 0; i < a.n_boxes; i++) {
            mbar_wait(&empty[s], ph ^ 1u);
            mbar_expect(&full[s], stage_bytes);
            for (int j = 0; j < a.per_stage; j++) {
                const int b = idx % (n_row_boxes * n_col_boxes);
                idx++;
                tma_load_2d(&map, &full[s], base + s * stage_bytes + j * a.box_rows * 128, (b % n_col_boxes) * 64, (b / n_col_boxes) * a.box_rows);
            }
            if (++s == a.stages) { s = 0; ph ^= 1u; }
        }
    } else if (threadIdx.x == 32) {            // consumer
        int s = 0; uint32_t ph = 0;
        for (int i = 0; i < a.n_boxes; i++) {
            mbar_wait(&full[s], ph);
            mbar_arrive(&empty[s]);
            if (++s == a.stages) { s = 0; ph ^= 1u; }
        }
        cycles[blockIdx.x] = 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void *fnp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
    EncodeFn enc = (EncodeFn)fnp;
    unsigned long long *d_cyc; cudaMalloc(&d_cyc, 148 * 8);
    cudaFuncSetAttribute(ingest, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    printf("# buffer  cols  box_rows per_stage stages grid same | bytes/cycle/SM (min..max over SMs), aggregate TB/s at 1.9 GHz\n");
    for (int big = 0; big < 2; big++) {
        const int cols = big ? 1024 : 256;                     // row pitch 2048 B or 512 B (bf16)
        const int64_t rows = big ? (1 << 18) : 2048;             // 512 MB (HBM) or 1 MB (L2-resident)
        void *buf; cudaMalloc(&buf, rows * cols * 2); cudaMemset(buf, 0, rows * cols * 2);
        for (int box_rows : {64, 128}) {
            CUtensorMap map;
            cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows}; cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
            cuuint32_t box[2] = {64, (cuuint32_t)box_rows}; cuuint32_t es[2] = {1, 1};
            enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            for (int per_stage : {1, 2})
                for (int stages : {2, 4, 8})
                    for (int grid : {1, 148})
                        for (int same : {1, 0}) {
                            if (grid == 1 && same == 0) continue;
                            if (stages * per_stage * box_rows * 128 > 190 * 1024) continue;
                            Args a = {stages, box_rows, 2000, (int)rows, cols, same, per_stage};
                            for (int rep = 0; rep < 2; rep++) ingest<<<grid, 128, 196 * 1024>>>(map, a, d_cyc);
                            cudaError_t e = cudaDeviceSynchronize();
                            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                            std::vector<unsigned long long> h(grid);
                            cudaMemcpy(h.data(), d_cyc, grid * 8, cudaMemcpyDeviceToHost);
                            unsigned long long mn = ~0ull, mx = 0;
                            for (auto c : h) { mn = c < mn ? c : mn; mx = c > mx ? c : mx; }
                            const double bytes = 2000.0 * per_stage * box_rows * 128;
                            printf("%s pitch %4d B  box %3d rows x%d  stages %d  grid %3d  %s | %6.1f .. %6.1f B/cyc/SM   %5.2f TB/s\n", big ? "512MB" : "  1MB",
                                   cols * 2, box_rows, per_stage, stages, grid, same ? "same " : "spread", bytes / mx, bytes / mn, grid * bytes / mx * 1.9e9 / 1e12);
                        }
        }
        cudaFree(buf);
    }
    return 0;
}

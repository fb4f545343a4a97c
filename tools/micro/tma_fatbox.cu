// Microbenchmark 2: one TMA instruction for SEVERAL 64-column k-blocks of an operand tile: 3-D view (64 cols, rows, k-blocks) of a
// row-major (rows x K) bf16 matrix with strides (row pitch, 128 B) — the k-block stride is SMALLER than the row stride — and a
// box (64, R, NKB).  Checks (a) that cuTensorMapEncodeTiled accepts it, (b) that shared memory receives [kb][row][128 B]
// with the SWIZZLE_128B pattern of R-row tiles, i.e. NKB consecutive UMMA operand k-blocks, (c) the ingest rate.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t ph) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap *m, uint64_t *bar, void *dst, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

__global__ void __launch_bounds__(128, 1) dump(const __grid_constant__ CUtensorMap map, int R, int NKB, int row0, int kb0, uint4 *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect(&bar, R * NKB * 128);
        tma_load_3d(&map, &bar, base, 0, row0, kb0);
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < R * NKB * 8; i += blockDim.x) out[i] = reinterpret_cast<uint4 *>(base)[i];
}

struct Args { int stages, R, NKB, n_boxes, rows_total, kb_total; };
__global__ void __launch_bounds__(128, 1) ingest(const __grid_constant__ CUtensorMap map, Args a, unsigned long long *cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
    __shared__ uint64_t full[8], empty[8];
    const int stage_bytes = a.R * a.NKB * 128;
    if (threadIdx.x == 0) {
        for (int i = 0; i < a.stages; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned long long t0 = 0;
    if (threadIdx.x == 0) {
        t0 = clock64();
        int s = 0; uint32_t ph = 0;
        int idx = blockIdx.x * 977;
        const int nrb = a.rows_total / a.R, nkb = a.kb_total / a.NKB;
        for (int i = 0; i < a.n_boxes; i++) {
            mbar_wait(&empty[s], ph ^ 1u);
            mbar_expect(&full[s], stage_bytes);
            const int b = idx++ % (nrb * nkb);
            tma_load_3d(&map, &full[s], base + s * stage_bytes, 0, (b / nkb) * a.R, (b % nkb) * a.NKB);
            if (++s == a.stages) { s = 0; ph ^= 1u; }
        }
    } else if (threadIdx.x == 32) {
        int s = 0; uint32_t ph = 0;
        for (int i = 0; i < a.n_boxes; i++) {
            mbar_wait(&full[s], ph);
            mbar_arrive(&empty[s]);
            if (++s == a.stages) { s = 0; ph ^= 1u; }
        }
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
    const int rows = 2048, K = 256;
    std::vector<__nv_bfloat16> h((size_t)rows * K);
    for (int r = 0; r < rows; r++) for (int k = 0; k < K; k++) h[(size_t)r * K + k] = __float2bfloat16((float)((r * 7 + k) % 251));
    __nv_bfloat16 *d; cudaMalloc(&d, h.size() * 2); cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(dump, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(ingest, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int R : {64, 128}) for (int NKB : {2, 4}) {
        CUtensorMap map;
        cuuint64_t dims[3] = {64, (cuuint64_t)rows, (cuuint64_t)(K / 64)};
        cuuint64_t strides[2] = {(cuuint64_t)K * 2, 128};
        cuuint32_t box[3] = {64, (cuuint32_t)R, (cuuint32_t)NKB}; cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("R %d NKB %d: encode -> %d\n", R, NKB, (int)r);
        if (r != CUDA_SUCCESS) continue;
        uint4 *out; cudaMalloc(&out, R * NKB * 128);
        const int row0 = 264, kb0 = (NKB == 2) ? 2 : 0;
        dump<<<1, 128, 196 * 1024>>>(map, R, NKB, row0, kb0, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("  dump error %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<__nv_bfloat16> s((size_t)R * NKB * 64);
        cudaMemcpy(s.data(), out, s.size() * 2, cudaMemcpyDeviceToHost);
        long bad = 0;
        for (int kb = 0; kb < NKB; kb++) for (int rr = 0; rr < R; rr++) for (int c = 0; c < 8; c++) for (int i = 0; i < 8; i++) {
            const size_t off = (size_t)kb * R * 64 + (size_t)rr * 64 + (size_t)((c ^ (rr & 7)) * 8) + i;
            const float want = (float)(((row0 + rr) * 7 + (kb0 + kb) * 64 + c * 8 + i) % 251);
            if (__bfloat162float(s[off]) != want) bad++;
        }
        printf("  layout [kb][row][128B swizzled]: %ld mismatches of %d\n", bad, R * NKB * 64);
        unsigned long long *d_cyc; cudaMalloc(&d_cyc, 148 * 8);
        for (int stages : {2, 4}) {
            if (stages * R * NKB * 128 > 190 * 1024) continue;
            Args a = {stages, R, NKB, 2000, rows, K / 64};
            for (int rep = 0; rep < 2; rep++) ingest<<<148, 128, 196 * 1024>>>(map, a, d_cyc);
            cudaDeviceSynchronize();
            std::vector<unsigned long long> hc(148);
            cudaMemcpy(hc.data(), d_cyc, 148 * 8, cudaMemcpyDeviceToHost);
            unsigned long long mx = 0; for (auto c : hc) mx = c > mx ? c : mx;
            const double bytes = 2000.0 * R * NKB * 128;
            printf("  stages %d grid 148: %.1f B/cyc/SM (%.0f cycles per %d KB instruction), %.2f TB/s aggregate at 1.9 GHz\n", stages, bytes / mx,
                   (double)mx / 2000, R * NKB * 128 / 1024, 148 * bytes / mx * 1.9e9 / 1e12);
        }
    }
    return 0;
}

// tcgen05 / TMEM / TMA bf16 GEMM + conv1d(k=3) for sm_100a.
//
//   out[seq, t, n] = epi( sum_{tap, k} A[seq, t + (tap - taps/2) * dil, k] * W[n, tap, k] )
//
// One CTA computes a 128 (time steps) x BN (output channels, <= 256) tile:
//   warp 0      : TMA producer — per k-iteration one A box (64 ch x 128 rows, 128B-swizzled) and one
//                 W box (64 ch x BN rows).  A k=3 convolution is an implicit GEMM: the three taps are
//                 three time-shifted TMA loads of the SAME activation tensor accumulating into the same
//                 TMEM tile; rows outside the sequence are zero-filled by TMA (= conv zero padding).
//   warp 1      : allocates TMEM, issues tcgen05.mma (kind::f16, bf16 x bf16 -> fp32, M=128, N=BN,
//                 K=16 per instruction, operands straight from the swizzled shared-memory tiles),
//                 tcgen05.commit releases each smem stage and finally signals the accumulator.
//   warps 2..5  : epilogue — tcgen05.ld (32 lanes x 32 columns per warp and instruction: thread =
//                 output row), shared epilogue (bias / act / LayerScale / residual / mask), 16-byte
//                 row-contiguous stores of fp32 and/or bf16.
// 4-stage mbarrier ring between TMA and MMA.  Shared memory: 4 x (16 KB + BN*128 B) <= 192 KB.
#include <cuda.h>

#include "gemm_common.cuh"

namespace decaf {

constexpr int TBM = 128, TBK = 64, TSTAGES = 4, TC_THREADS = 192;
constexpr int MAX_GROUP = 3;

struct TcMaps { CUtensorMap a[MAX_GROUP]; CUtensorMap w[MAX_GROUP]; };

// ---------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap *map, uint64_t *bar, void *dst, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// K-major, 128B-swizzled operand tile (rows of 64 bf16 = 128 B, 8-row groups 1024 B apart)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // start address, 16-byte units
    d |= (uint64_t)1 << 16;                            // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset: 8 rows x 128 B
    d |= (uint64_t)1 << 46;                            // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[32]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

// ---------------------------------------------------------------------------------- kernel
template <int TMEM_COLS>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ TcMaps maps, GemmArgs p, int BN, int flat, int tiles_per_seq, int kb_per_tap) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve-up (base re-aligned to 1024 B: SWIZZLE_128B atoms)
    uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int a_bytes = TBM * TBK * 2;                 // 16 KB
    const int b_bytes = BN * TBK * 2;                  // multiple of 2 KB
    uint8_t *smem_a = base;
    uint8_t *smem_b = base + TSTAGES * a_bytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_b + TSTAGES * b_bytes);
    uint64_t *empty = full + TSTAGES;
    uint64_t *tmem_full = empty + TSTAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.z;
    const int n0 = blockIdx.y * BN;
    const int tile = blockIdx.x;
    int seq_c, t0;                                      // TMA coordinates of the tile's first row
    if (flat) { seq_c = 0; t0 = tile * TBM; }
    else { seq_c = tile / tiles_per_seq; t0 = (tile % tiles_per_seq) * TBM; }
    const int n_iters = p.taps * kb_per_tap;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&maps.a[g]);
        prefetch_tmap(&maps.w[g]);
        for (int s = 0; s < TSTAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------------------------ TMA producer
            const uint32_t tx_bytes = (uint32_t)(a_bytes + b_bytes);
            for (int it = 0; it < n_iters; it++) {
                const int s = it % TSTAGES;
                const uint32_t ph = (uint32_t)(it / TSTAGES) & 1u;
                mbar_wait(&empty[s], ph ^ 1u);
                const int tap = it / kb_per_tap, kb = it % kb_per_tap;
                const int shift = (tap - p.taps / 2) * p.dil;
                mbar_expect_tx(&full[s], tx_bytes);
                tma_load_3d(&maps.a[g], &full[s], smem_a + s * a_bytes, kb * TBK, t0 + shift, seq_c);
                tma_load_3d(&maps.w[g], &full[s], smem_b + s * b_bytes, kb * TBK, tap, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ------------------------------------------------ MMA issuer
            // instruction descriptor: D fp32, A/B bf16, both K-major, N = BN, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
            for (int it = 0; it < n_iters; it++) {
                const int s = it % TSTAGES;
                const uint32_t ph = (uint32_t)(it / TSTAGES) & 1u;
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint64_t adesc = umma_desc_sw128(smem_u32(smem_a + s * a_bytes));
                const uint64_t bdesc = umma_desc_sw128(smem_u32(smem_b + s * b_bytes));
#pragma unroll
                for (int k = 0; k < TBK / 16; k++)      // +32 B (= 2 x 16 B units) per K = 16 slice inside the swizzle row
                    umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (it > 0 || k > 0) ? 1u : 0u);
                umma_commit(&empty[s]);                 // frees this smem stage once the MMAs above retire
            }
            umma_commit(tmem_full);                     // accumulator complete
        }
    } else {
        // ---------------------------------------------------- epilogue: thread = output row
        const int quarter = warp & 3;                   // TMEM lane quarter this warp may access
        const int row_in_tile = quarter * 32 + lane;
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        int seq, t;
        bool row_ok;
        if (flat) {
            const int64_t r = (int64_t)tile * TBM + row_in_tile;
            row_ok = r < (int64_t)p.n_seq * p.rows_per_seq;
            seq = (int)(r / p.rows_per_seq); t = (int)(r % p.rows_per_seq);
        } else {
            seq = seq_c; t = t0 + row_in_tile;
            row_ok = t < p.rows_per_seq;
        }
        const float *bias = p.bias ? p.bias + (int64_t)g * p.g_stride_bias : nullptr;
        float *of = p.out_f32 ? p.out_f32 + (int64_t)g * p.g_stride_out_f32 : nullptr;
        bf16 *oa = p.out_act ? reinterpret_cast<bf16 *>(p.out_act) + (int64_t)g * p.g_stride_out_act : nullptr;
        float rm = 1.f;
        if (row_ok && p.rowmask) rm = (float)p.rowmask[(int64_t)seq * p.m_seq_stride + t];
        const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16);
        for (int c = 0; c < BN; c += 32) {
            float v[32];
            const int width = min(32, BN - c);          // BN is a multiple of 16
            if (width == 32) tmem_ld32(trow + (uint32_t)c, v); else tmem_ld16(trow + (uint32_t)c, v);
            if (!row_ok) continue;
            const int nbase = n0 + c;
#pragma unroll
            for (int i = 0; i < 32; i++) {
                const int n = nbase + i;
                if (i < width && n < p.N) v[i] = gemm_epilogue_value(p, v[i], seq, t, n, rm, bias, g);
            }
            if (of) {
                float *dst = of + ((int64_t)seq * p.o_seq_stride + t) * p.ldo + nbase;
                if (nbase + width <= p.N && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        if (i < width) *reinterpret_cast<float4 *>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                } else {
                    for (int i = 0; i < width; i++) if (nbase + i < p.N) dst[i] = v[i];
                }
            }
            if (oa) {
                bf16 *dst = oa + ((int64_t)seq * p.o2_seq_stride + t) * p.ldo2 + nbase;
                if (nbase + width <= p.N && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
                    for (int i = 0; i < 32; i += 8) {
                        if (i < width) {
                            uint4 pk;
                            __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&pk);
#pragma unroll
                            for (int j = 0; j < 4; j++) h[j] = __floats2bfloat162_rn(v[i + 2 * j], v[i + 2 * j + 1]);
                            *reinterpret_cast<uint4 *>(dst + i) = pk;
                        }
                    }
                } else {
                    for (int i = 0; i < width; i++) if (nbase + i < p.N) dst[i] = __float2bfloat16_rn(v[i]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

static int pick_bn(int N) {
    const int n_tiles = (N + 255) / 256;
    int bn = (N + n_tiles - 1) / n_tiles;
    bn = (bn + 15) / 16 * 16;
    return bn;
}

const char *gemm_tc_why_not(const GemmArgs &a, int dtype) {
    static int sm100 = -1;
    if (sm100 < 0) sm100 = decaf_device_is_sm100();
    if (!sm100) return "device is not sm_100";
    if (dtype != DECAF_BF16) return "activation dtype is not bf16";
    if (a.K % 8 != 0) return "K is not a multiple of 8 (16-byte TMA pitch)";
    if (a.lda % 8 != 0) return "lda is not a multiple of 8";
    if ((reinterpret_cast<uintptr_t>(a.A) & 15) || (reinterpret_cast<uintptr_t>(a.W) & 15)) return "operands not 16-byte aligned";
    if ((a.a_seq_stride * a.lda) % 8 != 0) return "sequence pitch not 16-byte aligned";
    if ((int64_t)a.n_seq * a.rows_per_seq < 64) return "too few rows for a 128-row tensor-core tile";
    if (get_encode() == nullptr) return "cuTensorMapEncodeTiled not available";
    return nullptr;
}

static int encode_3d(CUtensorMap *m, const void *ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1_bytes,
                     uint64_t s2_bytes, uint32_t b0, uint32_t b1, uint32_t b2) {
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {s1_bytes, s2_bytes};
    cuuint32_t box[3] = {b0, b1, b2};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = get_encode()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(ptr), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): dims %llu,%llu,%llu strides %llu,%llu box %u,%u,%u", (int)r,
                  (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, (unsigned long long)s1_bytes,
                  (unsigned long long)s2_bytes, b0, b1, b2);
        return 1;
    }
    return 0;
}

int gemm_tc_launch(const GemmArgs &a, int n_group, cudaStream_t st) {
    DECAF_CHECK(n_group <= MAX_GROUP, "decaf_gemm(tcgen05): at most %d groups", MAX_GROUP);
    const int BN = pick_bn(a.N);
    const int n_tiles = cdiv(a.N, BN);
    const int flat = (a.taps == 1 && a.a_seq_stride == a.rows_per_seq) ? 1 : 0;
    const int64_t M = (int64_t)a.n_seq * a.rows_per_seq;
    const int tiles_per_seq = cdiv(a.rows_per_seq, TBM);
    const int m_tiles = flat ? cdiv(M, TBM) : a.n_seq * tiles_per_seq;
    const int kb_per_tap = cdiv(a.K, TBK);
    TcMaps maps;
    for (int g = 0; g < n_group; g++) {
        const bf16 *A = reinterpret_cast<const bf16 *>(a.A) + (int64_t)g * a.g_stride_a;
        const bf16 *W = reinterpret_cast<const bf16 *>(a.W) + (int64_t)g * a.g_stride_w;
        if (flat) {
            if (encode_3d(&maps.a[g], A, a.K, M, 1, a.lda * 2, (uint64_t)M * a.lda * 2, TBK, TBM, 1)) return 1;
        } else {
            if (encode_3d(&maps.a[g], A, a.K, a.rows_per_seq, a.n_seq, a.lda * 2, (uint64_t)a.a_seq_stride * a.lda * 2,
                          TBK, TBM, 1)) return 1;
        }
        if (encode_3d(&maps.w[g], W, a.K, a.taps, a.N, (uint64_t)a.K * 2, (uint64_t)a.taps * a.K * 2, TBK, 1, BN)) return 1;
    }
    const size_t smem = 1024 + (size_t)TSTAGES * (TBM * TBK * 2 + BN * TBK * 2) + 256;
    dim3 grid(m_tiles, n_tiles, n_group);
#define TC_LAUNCH(COLS)                                                                                         \
    do {                                                                                                        \
        DECAF_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<COLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        gemm_tc_kernel<COLS><<<grid, TC_THREADS, smem, st>>>(maps, a, BN, flat, tiles_per_seq, kb_per_tap);     \
    } while (0)
    if (BN <= 32) TC_LAUNCH(32);
    else if (BN <= 64) TC_LAUNCH(64);
    else if (BN <= 128) TC_LAUNCH(128);
    else TC_LAUNCH(256);
#undef TC_LAUNCH
    DECAF_LAUNCH_CHECK();
    return 0;
}

}  // namespace decaf

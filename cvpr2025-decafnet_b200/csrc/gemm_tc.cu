// tcgen05 / TMEM / TMA bf16 GEMM + conv1d(k=3) for sm_100a — persistent, warp-specialised.
//
//   out[seq, t, n] = epi( sum_{tap, k} A[seq, t + (tap - taps/2) * dil, k] * W[n, tap, k] )
//
// One CTA per SM loops over 128 (time steps) x BN (output channels) tiles:
//   warp 0      : TMA producer — per k-iteration one A box (64 ch x 128 rows, 128B-swizzled) and the W
//                 boxes (64 ch x BN rows).  A k=3 convolution is an implicit GEMM: the three taps are three
//                 time-shifted TMA loads of the SAME activation tensor accumulating into the same TMEM
//                 tile; rows outside the sequence are zero-filled by TMA (= conv zero padding).
//   warp 1      : allocates TMEM (512 columns = two accumulator stages of <= 256 columns, or one of <= 512),
//                 issues tcgen05.mma (kind::f16, bf16 x bf16 -> fp32, M = 128, N <= 256 per instruction,
//                 K = 16), tcgen05.commit releases each smem stage and signals the accumulator stage.
//   warps 2..9  : epilogue, overlapped with the next tile's main loop through the second TMEM stage.
//                 tcgen05.ld gives "thread = output row"; every 32-column chunk is transposed through a
//                 swizzled (bank-conflict-free) shared-memory slab so that bias / activation / LayerScale /
//                 residual / PE / mask are applied, and fp32 / bf16 results stored, with fully coalesced
//                 16-byte global accesses (8 lanes cover 32 consecutive channels of one row).
//                 Optional fused channel LayerNorm (two-pass, like libs/modeling/blocks.py:125-131): the
//                 whole output row lives in one TMEM lane, so mean / variance are thread-local sums over
//                 the accumulator columns (re-read from TMEM per pass), exchanged once between the two
//                 warps that share a lane quarter.
// mbarrier rings: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue).
#include <cuda.h>

#include "gemm_common.cuh"

namespace decaf {

constexpr int TBM = 128, TBK = 64;
constexpr int EPI_WARPS = 8;
constexpr int TC_THREADS = 64 + 32 * EPI_WARPS;       // 320
constexpr int MAX_GROUP = 3;
constexpr int MAX_STAGES = 8;
constexpr int A_BYTES = TBM * TBK * 2;                // 16 KB
constexpr int STAGING_BYTES = EPI_WARPS * 4096;       // one 32 x 32 fp32 slab per epilogue warp
constexpr int LNX_BYTES = 2 * 2 * TBM * 4;            // [pass][column half][row]
constexpr int BIAS_BYTES = 512 * 4;
constexpr int SMEM_LIMIT = 232448;                    // 227 KB

struct TcMaps { CUtensorMap a[MAX_GROUP]; CUtensorMap w[MAX_GROUP]; };

struct TcSched {
    int BN;             // CTA tile width (output channels)
    int n_mma;          // MMAs per k-step (BN / n_mma columns each, <= 256)
    int acc_stages;     // TMEM accumulator stages (2 when 2 * BN <= 512)
    int acc_stride;     // TMEM columns between stages
    int stages;         // smem pipeline depth
    int flat, tiles_per_seq, kb_per_tap;
    int m_tiles, n_tiles, n_group, total_tiles;
    int w_res;          // 1: the CTA's (group, n tile) weights stay resident in smem, the ring carries A only
    unsigned long long *trace;   // debug: per-role clock64 stamps of CTA 0 (decaf_debug_gemm_trace), else NULL
};

// ---------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
__device__ __forceinline__ void pair_barrier(int id) {      // the two epilogue warps of one TMEM lane quarter
    asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory");
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
// 32 lanes x 32 consecutive fp32 columns: register i of lane l = accumulator[row l of the quarter][col + i]
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

// erf with |error| <= 1.5e-7 (Abramowitz-Stegun 7.1.26) on the SFU: the bf16 path rounds the GELU
// output to 8 mantissa bits, so this is indistinguishable from erff() there and ~3x cheaper.
__device__ __forceinline__ float gelu_fast(float x) {
    const float z = fabsf(x) * 0.70710678118654752440f;
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));   // MUFU.RCP, ~1 ulp
    float poly = fmaf(1.061405429f, t, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    const float e = 1.0f - poly * t * __expf(-z * z);   // erf(|x| / sqrt 2)
    return 0.5f * x * (1.0f + copysignf(e, x));
}

constexpr int TRACE_SLOTS = 2048;                     // per role
__device__ __forceinline__ void trace_put(unsigned long long *tr, int role, int &n) {
    if (tr != nullptr && blockIdx.x == 0 && n < TRACE_SLOTS) tr[role * TRACE_SLOTS + n++] = clock64();
}

__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }

// ---------------------------------------------------------------------------------- kernel
// EPI >= 0 fixes the epilogue at compile time (bit 0 LN, bits 1-2 act, bit 3 colscale+resid, bit 4 fp32
// out, bit 5 bf16 out, bit 6 PE); EPI < 0 is the generic variant that reads the flags from GemmArgs.
constexpr int epi_code(bool ln, int act, bool res, bool f32, bool b16, bool pe) {
    return (ln ? 1 : 0) | (act << 1) | (res ? 8 : 0) | (f32 ? 16 : 0) | (b16 ? 32 : 0) | (pe ? 64 : 0);
}

// Tile id -> (m tile, n tile, group).  combo = (group, n tile) is the fastest index so that (a) CTAs that
// run at the same time share the A tile through L2 and (b) with gridDim.x a multiple of `combos` every CTA
// keeps one combo for its whole life — the weight-resident mode relies on that.
struct TileIdx { int mt, nt, g; };
__device__ __forceinline__ TileIdx decode_tile(const TcSched &sc, int tile) {
    const int combos = sc.n_tiles * sc.n_group;
    const int combo = tile % combos;
    TileIdx t;
    t.mt = tile / combos; t.nt = combo % sc.n_tiles; t.g = combo / sc.n_tiles;
    return t;
}

template <int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ TcMaps maps, const __grid_constant__ GemmArgs p, const __grid_constant__ TcSched sc) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment (SWIZZLE_128B atoms) without laundering the pointer through an integer, so
    // the compiler keeps the shared address space (LDS/STS instead of generic accesses)
    uint8_t *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int b_bytes = sc.BN * TBK * 2;
    const int n_iters = p.taps * sc.kb_per_tap;
    uint8_t *smem_a = base;
    uint8_t *smem_b = smem_a + sc.stages * A_BYTES;     // W ring (stages blocks) or the resident W (n_iters blocks)
    float4 *staging = reinterpret_cast<float4 *>(smem_b + (sc.w_res ? n_iters : sc.stages) * b_bytes);
    float *lnx = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(staging) + STAGING_BYTES);
    float *bias_s = lnx + LNX_BYTES / 4;
    uint64_t *full = reinterpret_cast<uint64_t *>(bias_s + BIAS_BYTES / 4);
    uint64_t *empty = full + MAX_STAGES;
    uint64_t *tmem_full = empty + MAX_STAGES;
    uint64_t *tmem_empty = tmem_full + 2;
    uint64_t *w_full = tmem_empty + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(w_full + 1);

    constexpr bool G = EPI < 0;
    const bool f_ln = G ? (p.ln != 0) : ((EPI & 1) != 0);
    const int f_act = G ? p.act : ((EPI >> 1) & 3);
    const bool f_res = G ? (p.resid != nullptr) : (((EPI >> 3) & 1) != 0);
    const bool f_cs = G ? (p.colscale != nullptr) : (((EPI >> 3) & 1) != 0);
    const bool f_f32 = G ? (p.out_f32 != nullptr) : (((EPI >> 4) & 1) != 0);
    const bool f_b16 = G ? (p.out_act != nullptr) : (((EPI >> 5) & 1) != 0);
    const bool f_pe = G ? (p.pe != nullptr) : (((EPI >> 6) & 1) != 0);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bn_mma = sc.BN / sc.n_mma;
    if (sc.trace != nullptr && threadIdx.x == 0 && blockIdx.x < 256) {   // debug: per-CTA start time (ns)
        unsigned long long gt;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
        sc.trace[3 * TRACE_SLOTS + blockIdx.x] = gt;
    }

    if (warp == 0 && lane == 0) {
        for (int g = 0; g < sc.n_group; g++) { prefetch_tmap(&maps.a[g]); prefetch_tmap(&maps.w[g]); }
        for (int s = 0; s < sc.stages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; s++) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], EPI_WARPS); }
        mbar_init(w_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (f_ln) {                                       // LN mode: a single n-tile, one group -> bias is tile independent
        for (int i = threadIdx.x; i < 512; i += TC_THREADS) bias_s[i] = (p.bias && i < p.N) ? p.bias[i] : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------------------------ TMA producer
            if (sc.w_res && (int)blockIdx.x < sc.total_tiles) {
                // weight-resident mode: this CTA's (group, n tile) never changes -> load its W once
                const TileIdx t = decode_tile(sc, blockIdx.x);
                mbar_expect_tx(w_full, (uint32_t)(n_iters * b_bytes));
                for (int it = 0; it < n_iters; it++)
                    tma_load_3d(&maps.w[t.g], w_full, smem_b + it * b_bytes, (it % sc.kb_per_tap) * TBK, it / sc.kb_per_tap,
                                t.nt * sc.BN);
            }
            const uint32_t tx_bytes = (uint32_t)(sc.w_res ? A_BYTES : A_BYTES + b_bytes);
            // single-thread role: no divisions in the loop (a dependent 32-bit division costs ~150 cycles)
            int s = 0, trn = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < sc.total_tiles; tile += gridDim.x) {
                const TileIdx t = decode_tile(sc, tile);
                int seq_c, t0;
                if (sc.flat) { seq_c = 0; t0 = t.mt * TBM; }
                else { seq_c = t.mt / sc.tiles_per_seq; t0 = (t.mt % sc.tiles_per_seq) * TBM; }
                const int n0 = t.nt * sc.BN;
                for (int tap = 0; tap < p.taps; tap++) {
                    const int shift = (tap - p.taps / 2) * p.dil;
                    for (int kb = 0; kb < sc.kb_per_tap; kb++) {
                        mbar_wait(&empty[s], ph ^ 1u);
                        trace_put(sc.trace, 0, trn);
                        mbar_expect_tx(&full[s], tx_bytes);
                        tma_load_3d(&maps.a[t.g], &full[s], smem_a + s * A_BYTES, kb * TBK, t0 + shift, seq_c);
                        if (!sc.w_res) {
                            for (int j = 0; j < sc.n_mma; j++)
                                tma_load_3d(&maps.w[t.g], &full[s], smem_b + s * b_bytes + j * bn_mma * TBK * 2, kb * TBK, tap,
                                            n0 + j * bn_mma);
                        }
                        if (++s == sc.stages) { s = 0; ph ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ------------------------------------------------ MMA issuer
            // instruction descriptor: D fp32, A/B bf16, both K-major, N = bn_mma, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn_mma >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
            int s = 0, as = 0, trn = 0;
            uint32_t ph = 0, aph = 0;
            if (sc.w_res && (int)blockIdx.x < sc.total_tiles) mbar_wait(w_full, 0);
            for (int tile = blockIdx.x; tile < sc.total_tiles; tile += gridDim.x) {
                mbar_wait(&tmem_empty[as], aph ^ 1u);
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(as * sc.acc_stride);
                for (int it = 0; it < n_iters; it++) {
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    trace_put(sc.trace, 1, trn);
                    const uint64_t adesc = umma_desc_sw128(smem_u32(smem_a + s * A_BYTES));
                    const uint8_t *wblk = smem_b + (sc.w_res ? it : s) * b_bytes;
                    for (int j = 0; j < sc.n_mma; j++) {
                        const uint64_t bdesc = umma_desc_sw128(smem_u32(wblk + j * bn_mma * TBK * 2));
#pragma unroll
                        for (int k = 0; k < TBK / 16; k++)   // +32 B (= 2 x 16 B units) per K = 16 slice inside the swizzle row
                            umma_bf16(tacc + (uint32_t)(j * bn_mma), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                                      (it > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty[s]);             // frees this smem stage once the MMAs above retire
                    if (++s == sc.stages) { s = 0; ph ^= 1u; }
                }
                umma_commit(&tmem_full[as]);            // accumulator stage complete
                if (++as == sc.acc_stages) { as = 0; aph ^= 1u; }
            }
        }
    } else {
        // ---------------------------------------------------- epilogue warps
        const int e = warp - 2;
        const int q = warp & 3;                         // TMEM lane quarter this warp may access
        const int h = e >> 2;                           // column half
        float4 *slab = staging + e * 256;               // 32 rows x 8 float4, chunk index XOR (row & 7)
        const int nch = (sc.BN + 31) / 32;
        const int c_begin = h ? (nch + 1) / 2 : 0;
        const int c_end = h ? nch : (nch + 1) / 2;
        const int cj = lane & 7, rsub = lane >> 3;
        const float *lx0 = lnx + q * 32 + lane;         // [pass][half][row] exchange slots of this thread's row
        float *lxw = lnx + h * TBM + q * 32 + lane;
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f), one4 = make_float4(1.f, 1.f, 1.f, 1.f);
        int as = 0, trn = 0;
        uint32_t aph = 0;
        for (int tile = blockIdx.x; tile < sc.total_tiles; tile += gridDim.x) {
            const TileIdx ti = decode_tile(sc, tile);
            const int mt = ti.mt, g = ti.g;
            const int n0 = ti.nt * sc.BN;

            // rows this lane handles in the coalesced phase: quarter row 4 i + rsub, i = 0..7.
            // Row indices (seq * seq_stride + t) per tensor are 32-bit; the pitch multiply is done at use.
            int ri_f[8], ri_a[8], ri_r[8], r_t[8];
            float r_m[8];
            {
                uint32_t seq, t;
                if (sc.flat) {
                    const uint32_t r = (uint32_t)mt * TBM + q * 32 + rsub;
                    seq = r / (uint32_t)p.rows_per_seq; t = r % (uint32_t)p.rows_per_seq;
                } else {
                    seq = mt / sc.tiles_per_seq; t = (mt % sc.tiles_per_seq) * TBM + q * 32 + rsub;
                }
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const bool ok = sc.flat ? (seq < (uint32_t)p.n_seq) : (t < (uint32_t)p.rows_per_seq);
                    r_t[i] = ok ? (int)t : -1;
                    ri_f[i] = (int)(seq * (uint32_t)p.o_seq_stride + t);
                    ri_a[i] = (int)(seq * (uint32_t)p.o2_seq_stride + t);
                    ri_r[i] = (int)(seq * (uint32_t)p.r_seq_stride + t);
                    r_m[i] = (ok && p.rowmask) ? (float)p.rowmask[(int64_t)seq * p.m_seq_stride + t] : 1.f;
                    t += 4;
                    if (sc.flat) {                      // rows_per_seq may be < 4: carry into the sequence index
                        while (t >= (uint32_t)p.rows_per_seq) { t -= p.rows_per_seq; seq++; }
                    }
                }
            }
            const float *bias = p.bias ? p.bias + (int64_t)g * p.g_stride_bias : nullptr;
            float *of = f_f32 ? p.out_f32 + (int64_t)g * p.g_stride_out_f32 : nullptr;
            bf16 *oa = f_b16 ? reinterpret_cast<bf16 *>(p.out_act) + (int64_t)g * p.g_stride_out_act : nullptr;

            // per-column parameters and residual / PE rows of one chunk, fetched one chunk ahead of their use
            // (x = x * m4 + a4 is the bias add or the LN affine)
            float4 a4 = zero4, m4 = one4, cs4 = one4, r4[8], e4[8];
            auto fetch = [&](int c, float4 &fa, float4 &fm, float4 &fcs, float4 (&fr)[8], float4 (&fe)[8]) {
                const int n = n0 + c * 32 + 4 * cj;
                const bool ok = c < c_end && c * 32 + 4 * cj < sc.BN && n < p.N;
                fa = zero4; fm = one4; fcs = one4;
                if (ok) {
                    if (f_ln) {
                        if (p.ln_w) { fm = ld4(p.ln_w + n); fa = ld4(p.ln_b + n); }
                    } else if (bias) {
                        fa = ld4(bias + n);
                    }
                    if (f_cs) fcs = ld4(p.colscale + n);
                }
                if (f_res) {
#pragma unroll
                    for (int i = 0; i < 8; i++)
                        fr[i] = (ok && r_t[i] >= 0) ? ld4(p.resid + (int64_t)ri_r[i] * p.ldr + n) : zero4;
                }
                if (f_pe) {
#pragma unroll
                    for (int i = 0; i < 8; i++)
                        fe[i] = (ok && r_t[i] >= 0) ? ld4(p.pe + (int64_t)r_t[i] * p.N + n) : zero4;
                }
            };
            fetch(c_begin, a4, m4, cs4, r4, e4);

            if (e == 0 && lane == 0) trace_put(sc.trace, 2, trn);      // tile setup done, waiting for the accumulator
            mbar_wait(&tmem_full[as], aph);
            tc_fence_after();
            if (e == 0 && lane == 0) trace_put(sc.trace, 2, trn);      // accumulator ready
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * sc.acc_stride);

            float rstd = 1.f, nmr = 0.f;                // LN: y = (v + bias) * rstd + nmr,  nmr = -mean * rstd
            if (f_ln) {
                // two-pass statistics over this thread's row (columns of both halves via the pair exchange)
                float s = 0.f;
                for (int c = c_begin; c < c_end; c++) {
                    float v[32];
                    tmem_ld32(trow + (uint32_t)(c * 32), v);
                    const float *bs = bias_s + c * 32;
                    if (c * 32 + 32 <= p.N) {
#pragma unroll
                        for (int i = 0; i < 32; i++) s += v[i] + bs[i];
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; i++) s += (c * 32 + i < p.N) ? v[i] + bs[i] : 0.f;
                    }
                }
                lxw[0] = s;
                pair_barrier(1 + q);
                const float mean = (lx0[0] + lx0[TBM]) / (float)p.N;
                float ss = 0.f;
                for (int c = c_begin; c < c_end; c++) {
                    float v[32];
                    tmem_ld32(trow + (uint32_t)(c * 32), v);
                    const float *bs = bias_s + c * 32;
                    if (c * 32 + 32 <= p.N) {
#pragma unroll
                        for (int i = 0; i < 32; i++) { const float d = v[i] + bs[i] - mean; ss = fmaf(d, d, ss); }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; i++) {
                            const float d = v[i] + bs[i] - mean;
                            ss += (c * 32 + i < p.N) ? d * d : 0.f;
                        }
                    }
                }
                lxw[2 * TBM] = ss;
                pair_barrier(1 + q);
                const float var = (lx0[2 * TBM] + lx0[3 * TBM]) / (float)p.N;
                rstd = 1.0f / sqrtf(var + p.ln_eps);
                nmr = -mean * rstd;
            }

            if (c_begin >= c_end) {                    // narrow tiles: this column half owns no chunk
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[as]);
            }
            for (int c = c_begin; c < c_end; c++) {
                {
                    float v[32];
                    tmem_ld32(trow + (uint32_t)(c * 32), v);
                    if (c == c_end - 1) {               // last TMEM read of this tile: hand the stage back to the MMA warp
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tmem_empty[as]);
                    }
                    if (e == 0 && lane == 0) trace_put(sc.trace, 2, trn);      // chunk: TMEM read done
                    if (f_ln) {
                        const float *bs = bias_s + c * 32;
#pragma unroll
                        for (int i = 0; i < 32; i++) v[i] = fmaf(v[i] + bs[i], rstd, nmr);
                    }
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        slab[lane * 8 + (j ^ (lane & 7))] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
                __syncwarp();
                if (e == 0 && lane == 0) trace_put(sc.trace, 2, trn);          // chunk: transposed
                // next chunk's parameters / residual rows are requested before this chunk is consumed
                float4 a4n, m4n, cs4n, r4n[8], e4n[8];
                fetch(c + 1, a4n, m4n, cs4n, r4n, e4n);
                const int n = n0 + c * 32 + 4 * cj;    // first of this lane's 4 channels
                if (c * 32 + 4 * cj < sc.BN && n < p.N) {
                    const float mm[4] = {m4.x, m4.y, m4.z, m4.w};
                    const float aa[4] = {a4.x, a4.y, a4.z, a4.w};
                    const float cs[4] = {cs4.x, cs4.y, cs4.z, cs4.w};
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int rr = 4 * i + rsub;
                        const float4 a = slab[rr * 8 + (cj ^ (rr & 7))];
                        float x[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                        for (int k = 0; k < 4; k++) x[k] = f_ln ? fmaf(x[k], mm[k], aa[k]) : x[k] + aa[k];
                        if (f_act == DECAF_ACT_RELU) {
#pragma unroll
                            for (int k = 0; k < 4; k++) x[k] = fmaxf(x[k], 0.f);
                        } else if (f_act == DECAF_ACT_GELU) {
#pragma unroll
                            for (int k = 0; k < 4; k++) x[k] = gelu_fast(x[k]);
                        }
                        if (f_res) {
                            const float rv[4] = {r4[i].x, r4[i].y, r4[i].z, r4[i].w};
#pragma unroll
                            for (int k = 0; k < 4; k++) x[k] = fmaf(x[k], cs[k], rv[k]);
                        } else if (f_cs) {
#pragma unroll
                            for (int k = 0; k < 4; k++) x[k] *= cs[k];
                        }
                        if (f_pe) { x[0] += e4[i].x; x[1] += e4[i].y; x[2] += e4[i].z; x[3] += e4[i].w; }
                        const float rm = r_m[i];
#pragma unroll
                        for (int k = 0; k < 4; k++) x[k] *= rm;
                        if (r_t[i] >= 0) {
                            if (f_f32)
                                *reinterpret_cast<float4 *>(of + (int64_t)ri_f[i] * p.ldo + n) = make_float4(x[0], x[1], x[2], x[3]);
                            if (f_b16) {
                                uint2 pk;
                                __nv_bfloat162 *hp = reinterpret_cast<__nv_bfloat162 *>(&pk);
                                hp[0] = __floats2bfloat162_rn(x[0], x[1]);
                                hp[1] = __floats2bfloat162_rn(x[2], x[3]);
                                *reinterpret_cast<uint2 *>(oa + (int64_t)ri_a[i] * p.ldo2 + n) = pk;
                            }
                        }
                    }
                }
                a4 = a4n; m4 = m4n; cs4 = cs4n;
                if (f_res) {
#pragma unroll
                    for (int i = 0; i < 8; i++) r4[i] = r4n[i];
                }
                if (f_pe) {
#pragma unroll
                    for (int i = 0; i < 8; i++) e4[i] = e4n[i];
                }
                __syncwarp();
            }
            if (e == 0 && lane == 0) trace_put(sc.trace, 2, trn);      // tile epilogue done
            if (++as == sc.acc_stages) { as = 0; aph ^= 1u; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (sc.trace != nullptr && threadIdx.x == 0 && blockIdx.x < 256) {   // debug: per-CTA end time (ns)
        unsigned long long gt;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
        sc.trace[3 * TRACE_SLOTS + 256 + blockIdx.x] = gt;
    }
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
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

static int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}

constexpr int FIXED_SMEM = 1024 + STAGING_BYTES + LNX_BYTES + BIAS_BYTES + (2 * MAX_STAGES + 6) * 8 + 16;
constexpr int W_RES_MAX = SMEM_LIMIT - FIXED_SMEM - 3 * A_BYTES;      // leave >= 3 A stages

// CTA tile width.  LN mode needs the full row in one CTA: N <= 512 split into n_mma instructions of <= 256
// columns.  Otherwise prefer the widest BN (multiple of 16, <= 256, >= 64) whose (taps x K x BN) weight block
// fits shared memory next to a 3-stage A ring ("weight resident": every CTA keeps its weights for its whole
// life and TMA only streams activations — the TMA unit issues ~1 128-byte row per 2 cycles, so re-loading a
// 256-row weight box per k-block would cost twice the activation traffic); else the widest BN <= 256.
static void pick_tile(int N, int K, int taps, int ln, int &BN, int &n_mma, int &w_res) {
    w_res = 0;
    if (ln) {
        n_mma = N <= 256 ? 1 : 2;
        BN = (N + 16 * n_mma - 1) / (16 * n_mma) * (16 * n_mma);
        return;
    }
    n_mma = 1;
    const int kpad = (K + TBK - 1) / TBK * TBK;
    const int n_min = (N + 255) / 256;
    for (int n_tiles = n_min; n_tiles <= 2 * n_min; n_tiles++) {   // at most 2x re-reads of A (from L2)
        int bn = (N + n_tiles - 1) / n_tiles;
        bn = (bn + 15) / 16 * 16;
        if ((int64_t)taps * kpad * bn * 2 <= W_RES_MAX) { BN = bn; w_res = 1; return; }
    }
    BN = (N + n_min - 1) / n_min;
    BN = (BN + 15) / 16 * 16;
}

const char *gemm_tc_why_not(const GemmArgs &a, int dtype) {
    static int sm100 = -1;
    if (sm100 < 0) sm100 = decaf_device_is_sm100();
    if (!sm100) return "device is not sm_100";
    if (dtype != DECAF_BF16) return "activation dtype is not bf16";
    if (a.K % 8 != 0) return "K is not a multiple of 8 (16-byte TMA pitch)";
    if (a.lda % 8 != 0) return "lda is not a multiple of 8";
    if (a.N % 8 != 0) return "N is not a multiple of 8";
    if ((reinterpret_cast<uintptr_t>(a.A) & 15) || (reinterpret_cast<uintptr_t>(a.W) & 15)) return "operands not 16-byte aligned";
    if ((a.a_seq_stride * a.lda) % 8 != 0) return "sequence pitch not 16-byte aligned";
    if ((int64_t)a.n_seq * a.rows_per_seq < 64) return "too few rows for a 128-row tensor-core tile";
    if (a.out_f32 && ((reinterpret_cast<uintptr_t>(a.out_f32) & 15) || a.ldo % 4 || a.g_stride_out_f32 % 4)) return "out_f32 not 16-byte aligned";
    if (a.out_act && ((reinterpret_cast<uintptr_t>(a.out_act) & 7) || a.ldo2 % 4 || a.g_stride_out_act % 4)) return "out_act not 8-byte aligned";
    if (a.resid && ((reinterpret_cast<uintptr_t>(a.resid) & 15) || a.ldr % 4)) return "resid not 16-byte aligned";
    if (a.bias && ((reinterpret_cast<uintptr_t>(a.bias) & 15) || a.g_stride_bias % 4)) return "bias not 16-byte aligned";
    if (a.colscale && (reinterpret_cast<uintptr_t>(a.colscale) & 15)) return "colscale not 16-byte aligned";
    if (a.pe && (reinterpret_cast<uintptr_t>(a.pe) & 15)) return "pe not 16-byte aligned";
    if (a.ln && a.N > 512) return "fused LayerNorm needs N <= 512";
    if (a.ln && a.ln_w && ((reinterpret_cast<uintptr_t>(a.ln_w) & 15) || (reinterpret_cast<uintptr_t>(a.ln_b) & 15))) return "ln_w/ln_b not 16-byte aligned";
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

static unsigned long long *g_trace = nullptr;

int gemm_tc_launch(const GemmArgs &a, int n_group, cudaStream_t st) {
    DECAF_CHECK(n_group <= MAX_GROUP, "decaf_gemm(tcgen05): at most %d groups", MAX_GROUP);
    DECAF_CHECK(!a.ln || n_group == 1, "decaf_gemm(tcgen05): fused LayerNorm does not support grouped launches");
    TcSched sc;
    pick_tile(a.N, a.K, a.taps, a.ln, sc.BN, sc.n_mma, sc.w_res);
    sc.n_tiles = cdiv(a.N, sc.BN);
    sc.acc_stages = 2 * sc.BN <= 512 ? 2 : 1;
    sc.acc_stride = sc.acc_stages == 2 ? 256 : 0;
    sc.flat = (a.taps == 1 && a.a_seq_stride == a.rows_per_seq) ? 1 : 0;
    const int64_t M = (int64_t)a.n_seq * a.rows_per_seq;
    DECAF_CHECK(M < (1ll << 31) - TBM, "decaf_gemm(tcgen05): too many rows");
    sc.tiles_per_seq = cdiv(a.rows_per_seq, TBM);
    sc.m_tiles = sc.flat ? cdiv(M, TBM) : a.n_seq * sc.tiles_per_seq;
    sc.kb_per_tap = cdiv(a.K, TBK);
    sc.n_group = n_group;
    const int combos = sc.n_tiles * n_group;
    if (sc.w_res && combos > num_sms()) sc.w_res = 0;
    sc.total_tiles = sc.m_tiles * combos;
    sc.trace = g_trace;
    const int b_bytes = sc.BN * TBK * 2;
    size_t smem;
    if (sc.w_res) {
        const int w_bytes = a.taps * sc.kb_per_tap * b_bytes;
        sc.stages = (SMEM_LIMIT - FIXED_SMEM - w_bytes) / A_BYTES;
        if (sc.stages > MAX_STAGES) sc.stages = MAX_STAGES;
        smem = (size_t)FIXED_SMEM + w_bytes + (size_t)sc.stages * A_BYTES;
    } else {
        sc.stages = (SMEM_LIMIT - FIXED_SMEM) / (A_BYTES + b_bytes);
        if (sc.stages > MAX_STAGES) sc.stages = MAX_STAGES;
        smem = (size_t)FIXED_SMEM + (size_t)sc.stages * (A_BYTES + b_bytes);
    }
    DECAF_CHECK(sc.stages >= 2, "decaf_gemm(tcgen05): tile does not fit shared memory (BN %d)", sc.BN);
    const int bn_mma = sc.BN / sc.n_mma;
    TcMaps maps;
    for (int g = 0; g < n_group; g++) {
        const bf16 *A = reinterpret_cast<const bf16 *>(a.A) + (int64_t)g * a.g_stride_a;
        const bf16 *W = reinterpret_cast<const bf16 *>(a.W) + (int64_t)g * a.g_stride_w;
        if (sc.flat) {
            if (encode_3d(&maps.a[g], A, a.K, M, 1, a.lda * 2, (uint64_t)M * a.lda * 2, TBK, TBM, 1)) return 1;
        } else {
            if (encode_3d(&maps.a[g], A, a.K, a.rows_per_seq, a.n_seq, a.lda * 2, (uint64_t)a.a_seq_stride * a.lda * 2,
                          TBK, TBM, 1)) return 1;
        }
        if (encode_3d(&maps.w[g], W, a.K, a.taps, a.N, (uint64_t)a.K * 2, (uint64_t)a.taps * a.K * 2, TBK, 1, bn_mma)) return 1;
    }
    // persistent grid: one CTA per SM; a multiple of `combos` in weight-resident mode so that tile ids
    // blockIdx.x + i * gridDim.x keep the CTA's (group, n tile)
    int grid = sc.total_tiles < num_sms() ? sc.total_tiles : num_sms();
    if (sc.w_res) {
        int per = num_sms() / combos;
        if (per > sc.m_tiles) per = sc.m_tiles;
        grid = per * combos;
    }
    const bool res = a.resid && a.colscale;
    const bool exact = (a.resid != nullptr) == (a.colscale != nullptr);      // variants tie colscale to resid
    const int code = exact ? epi_code(a.ln != 0, a.act, res, a.out_f32 != nullptr, a.out_act != nullptr, a.pe != nullptr) : -2;
#define TC_VARIANT(CODE)                                                                                          \
    if (code == (CODE)) {                                                                                         \
        static bool attr_set = false;                                                                             \
        if (!attr_set) {                                                                                          \
            DECAF_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<CODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT)); \
            attr_set = true;                                                                                      \
        }                                                                                                         \
        gemm_tc_kernel<CODE><<<grid, TC_THREADS, smem, st>>>(maps, a, sc);                                        \
        DECAF_LAUNCH_CHECK();                                                                                     \
        return 0;                                                                                                 \
    }
    TC_VARIANT(epi_code(false, DECAF_ACT_NONE, false, false, true, false))   // q/k/v, embd, AdaLN scale-shift projections
    TC_VARIANT(epi_code(false, DECAF_ACT_GELU, false, false, true, false))   // FFN fc
    TC_VARIANT(epi_code(false, DECAF_ACT_NONE, true, true, false, false))    // attention proj / FFN proj -> residual stream
    TC_VARIANT(epi_code(false, DECAF_ACT_NONE, true, true, true, false))     // FFN proj -> residual stream + FPN level copy
    TC_VARIANT(epi_code(false, DECAF_ACT_NONE, false, true, false, false))   // vid_map
    TC_VARIANT(epi_code(true, DECAF_ACT_RELU, false, false, true, false))    // conv -> LN -> ReLU (heads, embed convs)
    TC_VARIANT(epi_code(true, DECAF_ACT_RELU, false, true, false, true))     // last embed conv: + PE, fp32 residual stream
#undef TC_VARIANT
    {
        static bool attr_set = false;
        if (!attr_set) {
            DECAF_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<-1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
            attr_set = true;
        }
        gemm_tc_kernel<-1><<<grid, TC_THREADS, smem, st>>>(maps, a, sc);
    }
    DECAF_LAUNCH_CHECK();
    return 0;
}

}  // namespace decaf

// Debug hook (not part of the product path): the next tcgen05 GEMM launches write clock64 stamps of
// CTA 0 into buf[3][2048] (+ buf[3 * 2048 + i] / [.. + 256 + i]: start / end %globaltimer of CTA i < 256) (role 0 producer: stage acquired; 1 MMA: stage full; 2 epilogue warp 2:
// setup done / accumulator ready / tile done).  Pass NULL to switch it off.
extern "C" int decaf_debug_gemm_trace(unsigned long long *buf) {
    decaf::g_trace = buf;
    return 0;
}

// Final k=3 conv of a head tower (C -> 1 or 2 channels) for bf16 activations, on mma.sync tiles.
// out[row, o] = sum_{tap, c} x[row + tap - 1, c] * w[o, tap, c] + bias[o]   (ClsHead.cls_head / RegHead.reg_head,
// libs/modeling/head.py:59-60, 101-103) is a (rows x 3C) . (3C x n_out) product: a CTA stages 128 (+2 halo) rows of
// x in shared memory with 16-byte cp.async (every row is read from HBM once; the warp-per-row version re-read each row
// three times through L1 with 2-byte loads when C = 288), a warp multiplies 16 rows by the n_out <= 2 weight columns
// (padded to the 8 columns of m16n8k16).  The fp32 weights are split into bf16 hi + lo parts (two mma per step), so
// the product keeps ~16 mantissa bits of the weights like the fp32-FMA kernel it replaces.
#include <cstdlib>
#include "common.cuh"
#include "mma.cuh"

namespace decaf {


__device__ __forceinline__ int hm_level_of_row(const decaf_levels_t &lv, int r) {
    for (int l = 0; l < lv.n_levels; l++)
        if (r >= lv.off[l] && r < lv.off[l] + lv.len[l]) return l;
    return -1;
}

// Persistent: a CTA converts the weights once and walks row tiles blockIdx.x, + gridDim.x, ... with a two-stage
// shared-memory ring - the cp.async fetch of the next tile is in flight while the warps multiply the current one.
// (One tile per CTA exposed three dependent round trips per CTA - weights, rows, bias - with two CTAs per SM: 19 us
// for 42 MB at the NLQ shape.)
// KK > 0: C == 16 * KK at compile time - the warp's B fragments (hi / lo weight columns of all 3 * KK steps) stay in
// REGISTERS for the CTA's whole life and the inner loop is one ldmatrix + one mma per step with immediate offsets (with the
// fragments re-read from shared memory and run-time C, a 16-row tile cost ~800 instructions for 54 MMAs and 8 warps per
// SM could not hide their latency).  KK == 0: any C, fragments from shared memory.
template <int NOUT, int HM_WARPS, int KK>
__global__ void __launch_bounds__(32 * HM_WARPS)
head_out_mma_kernel(const bf16 *__restrict__ x, int64_t ldx, int rows_total, int C_rt, const float *__restrict__ w,
                    const float *__restrict__ bias, int mode, const float *__restrict__ level_scale, decaf_levels_t lv,
                    float *__restrict__ out) {
    constexpr int HM_ROWS = 16 * HM_WARPS;
    extern __shared__ __align__(16) uint8_t hm_smem[];
    const int C = KK > 0 ? 16 * KK : C_rt;
    const int ldx_s = C + 8, ldw = 3 * C + 8;
    const int tile_elems = (HM_ROWS + 2) * ldx_s;
    bf16 *Xs0 = reinterpret_cast<bf16 *>(hm_smem);                   // 2 x [HM_ROWS + 2][C + 8]
    bf16 *Wc = Xs0 + (size_t)2 * tile_elems;                          // [2 * NOUT][3C + 8]: hi parts of every output, then lo parts
    const int cpr = C / 8;
    const int n_tiles = (rows_total + HM_ROWS - 1) / HM_ROWS;
    auto load_tile = [&](int tile, int buf) {
        if (tile < n_tiles) {
            bf16 *Xs = Xs0 + (size_t)buf * tile_elems;
            const int r0 = tile * HM_ROWS;
            // (row, chunk) of this thread's pieces advance incrementally: one division per tile instead of one per 16-byte
            // piece (the i / cpr form was half of the kernel's instructions: 640 of 1300 per 16-row tile)
            int r = (int)threadIdx.x / cpr, c = (int)threadIdx.x - r * cpr;
            const int dr = (int)blockDim.x / cpr, dc = (int)blockDim.x - dr * cpr;
            for (; r < HM_ROWS + 2; ) {
                const int row = r0 - 1 + r;
                const bool ok = row >= 0 && row < rows_total;
                cp_async16(Xs + r * ldx_s + c * 8, x + (int64_t)(ok ? row : 0) * ldx + c * 8, ok);
                r += dr; c += dc;
                if (c >= cpr) { c -= cpr; r++; }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    load_tile(blockIdx.x, 0);
    for (int i = threadIdx.x; i < NOUT * 3 * C; i += blockDim.x) {
        const int o = i / (3 * C), k = i - o * 3 * C;
        const float v = w[i];                                         // w: (n_out, 3, C) == [o][tap * C + c]
        const bf16 hi = __float2bfloat16_rn(v);
        Wc[o * ldw + k] = hi;
        Wc[(NOUT + o) * ldw + k] = __float2bfloat16_rn(v - __bfloat162float(hi));
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
    const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lcol = (lane >> 4) * 8;
    // B tile of m16n8k16: column g < 2 * NOUT = {hi(o = 0..NOUT-1), lo(o = 0..NOUT-1)}: ONE mma per 16-channel step gives the
    // hi and lo products in neighbouring accumulator columns (two mma + four more LDS per step before)
    const bool has_b = g < 2 * NOUT;
    const bf16 *wb = Wc + (has_b ? g : 0) * ldw + 2 * t4;
    float bias_r[NOUT];
#pragma unroll
    for (int o = 0; o < NOUT; o++) bias_r[o] = bias[o];
    constexpr int NB_REG = KK > 0 ? 3 * KK : 1;
    uint32_t breg[NB_REG][2];
    if constexpr (KK > 0) {
        __syncthreads();                                              // Wc is complete
#pragma unroll
        for (int i = 0; i < 3 * KK; i++) {
            breg[i][0] = has_b ? *reinterpret_cast<const uint32_t *>(wb + i * 16) : 0u;
            breg[i][1] = has_b ? *reinterpret_cast<const uint32_t *>(wb + i * 16 + 8) : 0u;
        }
    }
    int buf = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, buf ^= 1) {
        load_tile(tile + gridDim.x, buf ^ 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();
        const int r0 = tile * HM_ROWS;
        if (r0 + warp * 16 < rows_total) {
            // one accumulator per tap: three independent MMA chains
            float acc[3][4];
#pragma unroll
            for (int tap = 0; tap < 3; tap++)
#pragma unroll
                for (int e = 0; e < 4; e++) acc[tap][e] = 0.f;
            const bf16 *xa = Xs0 + (size_t)buf * tile_elems + (warp * 16 + lrow) * ldx_s + lcol;   // smem row = row - (r0 - 1)
            if constexpr (KK > 0) {
#pragma unroll
                for (int kk = 0; kk < KK; kk++) {
#pragma unroll
                    for (int tap = 0; tap < 3; tap++) {
                        uint32_t af[4];
                        ldmatrix_x4(af, xa + tap * ldx_s + kk * 16);
                        mma_bf16(acc[tap], af[0], af[1], af[2], af[3], breg[tap * KK + kk][0], breg[tap * KK + kk][1]);
                    }
                }
            } else {
#pragma unroll 2
                for (int kk = 0; kk < C / 16; kk++) {
#pragma unroll
                    for (int tap = 0; tap < 3; tap++) {
                        uint32_t af[4];
                        ldmatrix_x4(af, xa + tap * ldx_s + kk * 16);
                        const int k = tap * C + kk * 16;
                        uint32_t b0 = 0, b1 = 0;
                        if (has_b) { b0 = *reinterpret_cast<const uint32_t *>(wb + k); b1 = *reinterpret_cast<const uint32_t *>(wb + k + 8); }
                        mma_bf16(acc[tap], af[0], af[1], af[2], af[3], b0, b1);
                    }
                }
            }
            // accumulator columns 2 * t4, 2 * t4 + 1 of rows g (e = 0, 1) and g + 8 (e = 2, 3)
            float s4[4];
#pragma unroll
            for (int e = 0; e < 4; e++) s4[e] = (acc[0][e] + acc[1][e]) + acc[2][e];
            float v2[2][NOUT];                                        // [row half][output] = hi + lo
            if constexpr (NOUT == 1) {                                // columns {hi0, lo0} both live in lane t4 == 0
                v2[0][0] = s4[0] + s4[1];
                v2[1][0] = s4[2] + s4[3];
            } else {                                                  // columns {hi0, hi1} in t4 == 0, {lo0, lo1} in t4 == 1
#pragma unroll
                for (int e = 0; e < 4; e++) s4[e] += __shfl_down_sync(0xffffffffu, s4[e], 1);
                v2[0][0] = s4[0]; v2[0][1] = s4[1];
                v2[1][0] = s4[2]; v2[1][1] = s4[3];
            }
            if (t4 == 0) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int row = r0 + warp * 16 + g + 8 * h;
                    if (row >= rows_total) continue;
                    const int level = hm_level_of_row(lv, row % lv.Pp);
#pragma unroll
                    for (int o = 0; o < NOUT; o++) {
                        float v = 0.f;
                        if (level >= 0) {
                            v = v2[h][o] + bias_r[o];
                            if (mode == 1) v = fmaxf(level_scale[level] * v, 0.f);
                        }
                        out[(int64_t)row * NOUT + o] = v;
                    }
                }
            }
        }
        __syncthreads();                                              // everyone is done with this buffer before it is refilled
    }
}

static inline size_t hm_smem_bytes(int C, int n_out, int warps) {
    return ((size_t)2 * (16 * warps + 2) * (C + 8) + (size_t)2 * n_out * (3 * C + 8)) * sizeof(bf16);
}
// 128-row tiles (8 warps, one CTA per SM) when two of them fit in shared memory, else 64-row tiles
static inline int hm_warps(int C, int n_out) {
    static int forced = -1;
    if (forced < 0) { const char *e = getenv("DECAF_HEADOUT_WARPS"); forced = e ? atoi(e) : 0; }
    if ((forced == 4 || forced == 8) && hm_smem_bytes(C, n_out, forced) <= 200 * 1024) return forced;
    if (hm_smem_bytes(C, n_out, 8) <= 200 * 1024) return 8;
    if (hm_smem_bytes(C, n_out, 4) <= 200 * 1024) return 4;
    return 0;
}

template <int NOUT, int NW, int KK>
static int hm_launch(const void *x, int64_t ldx, int rows_total, int C, const float *w, const float *bias, int mode,
                     const float *level_scale, const decaf_levels_t *lv, float *out, cudaStream_t st) {
    const size_t smem = hm_smem_bytes(C, NOUT, NW);
    const int n_tiles = cdiv(rows_total, 16 * NW);
    int per_sm = (int)((size_t)220 * 1024 / (smem + 1024));
    per_sm = per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm);
    if (KK > 0) per_sm = 1;                                       // ~150 registers per thread: one CTA per SM
    const int grid = n_tiles < 148 * per_sm ? n_tiles : 148 * per_sm;
    static size_t attr = 0;
    if (smem > 48 * 1024 && smem > attr) {
        DECAF_CUDA(cudaFuncSetAttribute(head_out_mma_kernel<NOUT, NW, KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = smem;
    }
    head_out_mma_kernel<NOUT, NW, KK><<<grid, 32 * NW, smem, st>>>((const bf16 *)x, ldx, rows_total, C, w, bias, mode, level_scale, *lv, out);
    DECAF_LAUNCH_CHECK();
    return 0;
}

template <int NOUT>
static int hm_dispatch(const void *x, int64_t ldx, int rows_total, int C, const float *w, const float *bias, int mode,
                       const float *level_scale, const decaf_levels_t *lv, float *out, cudaStream_t st) {
    const int nw = hm_warps(C, NOUT);
    if (nw == 8) {
        // the widths of the released configurations (embd 256: C = 256 and C + 32 = 288; embd 128: 128 and 160) keep the
        // weight fragments in registers
        if (C == 256) return hm_launch<NOUT, 8, 16>(x, ldx, rows_total, C, w, bias, mode, level_scale, lv, out, st);
        if (C == 288) return hm_launch<NOUT, 8, 18>(x, ldx, rows_total, C, w, bias, mode, level_scale, lv, out, st);
        if (C == 128) return hm_launch<NOUT, 8, 8>(x, ldx, rows_total, C, w, bias, mode, level_scale, lv, out, st);
        if (C == 160) return hm_launch<NOUT, 8, 10>(x, ldx, rows_total, C, w, bias, mode, level_scale, lv, out, st);
        return hm_launch<NOUT, 8, 0>(x, ldx, rows_total, C, w, bias, mode, level_scale, lv, out, st);
    }
    return hm_launch<NOUT, 4, 0>(x, ldx, rows_total, C, w, bias, mode, level_scale, lv, out, st);
}

int head_out_mma_launch(const void *x, int64_t ldx, int rows_total, int C, const float *w, const float *bias, int n_out,
                        int mode, const float *level_scale, const decaf_levels_t *lv, float *out, cudaStream_t st) {
    if (n_out == 1) return hm_dispatch<1>(x, ldx, rows_total, C, w, bias, mode, level_scale, lv, out, st);
    return hm_dispatch<2>(x, ldx, rows_total, C, w, bias, mode, level_scale, lv, out, st);
}

bool head_out_mma_ok(const void *x, int64_t ldx, int C) {
    return C % 16 == 0 && ldx % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && hm_warps(C, 2) > 0;
}

}  // namespace decaf

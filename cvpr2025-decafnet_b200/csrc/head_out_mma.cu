// Final k=3 conv of a head tower (C -> 1 or 2 channels) for bf16 activations, on mma.sync tiles.
// out[row, o] = sum_{tap, c} x[row + tap - 1, c] * w[o, tap, c] + bias[o]   (ClsHead.cls_head / RegHead.reg_head,
// libs/modeling/head.py:59-60, 101-103) is a (rows x 3C) . (3C x n_out) product: a CTA stages 128 (+2 halo) rows of
// x in shared memory with 16-byte cp.async (every row is read from HBM once; the warp-per-row version re-read each row
// three times through L1 with 2-byte loads when C = 288), a warp multiplies 16 rows by the n_out <= 2 weight columns
// (padded to the 8 columns of m16n8k16).  The fp32 weights are split into bf16 hi + lo parts (two mma per step), so
// the product keeps ~16 mantissa bits of the weights like the fp32-FMA kernel it replaces.
#include "common.cuh"
#include "mma.cuh"

namespace decaf {

constexpr int HM_WARPS = 8;
constexpr int HM_ROWS = 16 * HM_WARPS;

__device__ __forceinline__ int hm_level_of_row(const decaf_levels_t &lv, int r) {
    for (int l = 0; l < lv.n_levels; l++)
        if (r >= lv.off[l] && r < lv.off[l] + lv.len[l]) return l;
    return -1;
}

template <int NOUT>
__global__ void __launch_bounds__(32 * HM_WARPS)
head_out_mma_kernel(const bf16 *__restrict__ x, int64_t ldx, int rows_total, int C, const float *__restrict__ w,
                    const float *__restrict__ bias, int mode, const float *__restrict__ level_scale, decaf_levels_t lv,
                    float *__restrict__ out) {
    extern __shared__ __align__(16) uint8_t hm_smem[];
    const int ldx_s = C + 8, ldw = 3 * C + 8;
    bf16 *Xs = reinterpret_cast<bf16 *>(hm_smem);                    // [HM_ROWS + 2][C + 8]
    bf16 *Wh = Xs + (size_t)(HM_ROWS + 2) * ldx_s;                    // [NOUT][3C + 8] high parts
    bf16 *Wl = Wh + (size_t)NOUT * ldw;                               // low parts
    const int r0 = blockIdx.x * HM_ROWS;
    const int cpr = C / 8;
    for (int i = threadIdx.x; i < (HM_ROWS + 2) * cpr; i += blockDim.x) {
        const int r = i / cpr, c = i - r * cpr;
        const int row = r0 - 1 + r;
        const bool ok = row >= 0 && row < rows_total;
        cp_async16(Xs + r * ldx_s + c * 8, x + (int64_t)(ok ? row : 0) * ldx + c * 8, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (int i = threadIdx.x; i < NOUT * 3 * C; i += blockDim.x) {
        const int o = i / (3 * C), k = i - o * 3 * C;
        const float v = w[i];                                         // w: (n_out, 3, C) == [o][tap * C + c]
        const bf16 hi = __float2bfloat16_rn(v);
        Wh[o * ldw + k] = hi;
        Wl[o * ldw + k] = __float2bfloat16_rn(v - __bfloat162float(hi));
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
    const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lcol = (lane >> 4) * 8;
    if (r0 + warp * 16 >= rows_total) return;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const bool has_b = g < NOUT;
    const bf16 *wh = Wh + (has_b ? g : 0) * ldw + 2 * t4, *wl = Wl + (has_b ? g : 0) * ldw + 2 * t4;
    for (int tap = 0; tap < 3; tap++) {
        const bf16 *xa = Xs + (warp * 16 + lrow + tap) * ldx_s + lcol;     // smem row = row - (r0 - 1); tap shift = tap - 1
        for (int kk = 0; kk < C / 16; kk++) {
            uint32_t af[4];
            ldmatrix_x4(af, xa + kk * 16);
            const int k = tap * C + kk * 16;
            uint32_t bh0 = 0, bh1 = 0, bl0 = 0, bl1 = 0;
            if (has_b) {
                bh0 = *reinterpret_cast<const uint32_t *>(wh + k); bh1 = *reinterpret_cast<const uint32_t *>(wh + k + 8);
                bl0 = *reinterpret_cast<const uint32_t *>(wl + k); bl1 = *reinterpret_cast<const uint32_t *>(wl + k + 8);
            }
            mma_bf16(acc, af[0], af[1], af[2], af[3], bh0, bh1);
            mma_bf16(acc, af[0], af[1], af[2], af[3], bl0, bl1);
        }
    }
    if (t4 == 0) {                                                    // columns 0, 1 of rows g and g + 8
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int row = r0 + warp * 16 + g + 8 * h;
            if (row >= rows_total) continue;
            const int level = hm_level_of_row(lv, row % lv.Pp);
#pragma unroll
            for (int o = 0; o < NOUT; o++) {
                float v = 0.f;
                if (level >= 0) {
                    v = acc[2 * h + o] + bias[o];
                    if (mode == 1) v = fmaxf(level_scale[level] * v, 0.f);
                }
                out[(int64_t)row * NOUT + o] = v;
            }
        }
    }
}

int head_out_mma_launch(const void *x, int64_t ldx, int rows_total, int C, const float *w, const float *bias, int n_out,
                        int mode, const float *level_scale, const decaf_levels_t *lv, float *out, cudaStream_t st) {
    const size_t smem = ((size_t)(HM_ROWS + 2) * (C + 8) + (size_t)2 * n_out * (3 * C + 8)) * sizeof(bf16);
    const int grid = cdiv(rows_total, HM_ROWS);
    static size_t attr1 = 0, attr2 = 0;
    if (n_out == 1) {
        if (smem > 48 * 1024 && smem > attr1) {
            DECAF_CUDA(cudaFuncSetAttribute(head_out_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr1 = smem;
        }
        head_out_mma_kernel<1><<<grid, 32 * HM_WARPS, smem, st>>>((const bf16 *)x, ldx, rows_total, C, w, bias, mode, level_scale, *lv, out);
    } else {
        if (smem > 48 * 1024 && smem > attr2) {
            DECAF_CUDA(cudaFuncSetAttribute(head_out_mma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr2 = smem;
        }
        head_out_mma_kernel<2><<<grid, 32 * HM_WARPS, smem, st>>>((const bf16 *)x, ldx, rows_total, C, w, bias, mode, level_scale, *lv, out);
    }
    DECAF_LAUNCH_CHECK();
    return 0;
}

bool head_out_mma_ok(const void *x, int64_t ldx, int C) {
    const size_t smem = ((size_t)(HM_ROWS + 2) * (C + 8) + (size_t)4 * (3 * C + 8)) * sizeof(bf16);
    return C % 16 == 0 && ldx % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && smem <= 200 * 1024;
}

}  // namespace decaf

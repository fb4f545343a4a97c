// Row-wise (one warp per time step) memory-bound kernels: LayerNorm, the pre-attention
// depthwise-conv block, AdaLN, head output convs, pyramid masks, text-encoder glue.
// HBM-roofline kernels: every row is read once and written once, 16-byte vector accesses when
// the channel count allows, 8 rows per CTA so grids are large multiples of the SM count.
#include "common.cuh"

namespace decaf {

constexpr int ROWS_PER_CTA = 8;   // 8 warps

// ------------------------------------------------------------------------------- layernorm
template <int VEC, typename TO>
__global__ void __launch_bounds__(32 * ROWS_PER_CTA) layernorm_kernel(decaf_layernorm_t p) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * ROWS_PER_CTA + warp;
    if (row >= (int64_t)p.n_seq * p.rows_per_seq) return;
    const int seq = (int)(row / p.rows_per_seq), t = (int)(row % p.rows_per_seq);
    float v[VEC];
    load_row<VEC>(p.x + ((int64_t)seq * p.x_seq_stride + t) * p.ldx, lane, v);
    warp_layernorm<VEC>(v, p.C, p.eps);
    if (p.w) {
        float w[VEC], b[VEC];
        load_row<VEC>(p.w, lane, w);
        load_row<VEC>(p.b, lane, b);
#pragma unroll
        for (int i = 0; i < VEC; i++) v[i] = v[i] * w[i] + b[i];
    }
    if (p.relu) {
#pragma unroll
        for (int i = 0; i < VEC; i++) v[i] = fmaxf(v[i], 0.f);
    }
    if (p.pe) {
        float e[VEC];
        load_row<VEC>(p.pe + (int64_t)t * p.C, lane, e);
#pragma unroll
        for (int i = 0; i < VEC; i++) v[i] += e[i];
    }
    if (p.rowmask) {
        const float m = (float)p.rowmask[(int64_t)seq * p.m_seq_stride + t];
#pragma unroll
        for (int i = 0; i < VEC; i++) v[i] *= m;
    }
    if (p.out_f32) store_row<VEC>(p.out_f32 + ((int64_t)seq * p.o_seq_stride + t) * p.ldo, lane, v);
    if (p.out_act)
        store_row<VEC>(reinterpret_cast<TO *>(p.out_act) + ((int64_t)seq * p.o2_seq_stride + t) * p.ldo2, lane, v);
}

// ------------------------------------------------------------------------------- pre-attention
template <int VEC, typename TA>
__global__ void __launch_bounds__(32 * ROWS_PER_CTA) preattn_kernel(decaf_preattn_t p, int T_out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * ROWS_PER_CTA + warp;
    if (row >= (int64_t)p.n_seq * T_out) return;
    const int seq = (int)(row / T_out), t = (int)(row % T_out);
    const int C = p.C;
    const uint8_t *mrow = p.mask_in + (int64_t)seq * p.mi_seq_stride;

    float wp[VEC], bp[VEC];
    load_row<VEC>(p.w_pre, lane, wp);
    load_row<VEC>(p.b_pre, lane, bp);

    float ln[3][VEC];
    float skip[VEC];
    bool any_valid = false;
#pragma unroll
    for (int i = 0; i < VEC; i++) skip[i] = -INFINITY;
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const int tin = p.stride * t + j - 1;
        const bool ok = tin >= 0 && tin < p.T_in && mrow[tin] != 0;
        if (ok) {
            float x[VEC];
            load_row<VEC>(p.x + ((int64_t)seq * p.T_in + tin) * C, lane, x);
            if (p.skip_out) {
#pragma unroll
                for (int i = 0; i < VEC; i++) skip[i] = fmaxf(skip[i], x[i]);
                any_valid = true;
            }
            warp_layernorm<VEC>(x, C, p.eps);
#pragma unroll
            for (int i = 0; i < VEC; i++) ln[j][i] = x[i] * wp[i] + bp[i];
        } else {
#pragma unroll
            for (int i = 0; i < VEC; i++) ln[j][i] = 0.f;
        }
    }
    const uint8_t m_out = mrow[p.stride * t];
    if (p.skip_out) {
        const bool keep = any_valid && m_out != 0;
#pragma unroll
        for (int i = 0; i < VEC; i++) skip[i] = keep ? skip[i] : 0.f;
        store_row<VEC>(p.skip_out + row * C, lane, skip);
    }
    if (p.mask_out && lane == 0) p.mask_out[(int64_t)seq * p.mo_seq_stride + t] = m_out;

    for (int br = 0; br < p.n_branch; br++) {
        float y[VEC];
        const float *wd = p.wd + ((int64_t)br * C + lane * VEC) * 3;
#pragma unroll
        for (int i = 0; i < VEC; i++)
            y[i] = wd[i * 3 + 0] * ln[0][i] + wd[i * 3 + 1] * ln[1][i] + wd[i * 3 + 2] * ln[2][i];
        warp_layernorm<VEC>(y, C, p.eps);
        float w[VEC], b[VEC];
        load_row<VEC>(p.w_br + (int64_t)br * C, lane, w);
        load_row<VEC>(p.b_br + (int64_t)br * C, lane, b);
#pragma unroll
        for (int i = 0; i < VEC; i++) y[i] = y[i] * w[i] + b[i];
        store_row<VEC>(reinterpret_cast<TA *>(p.out_act) + (int64_t)br * p.out_branch_stride + row * C, lane, y);
    }
}

// ------------------------------------------------------------------------------- adaln
template <int VEC, typename TA>
__global__ void __launch_bounds__(32 * ROWS_PER_CTA) adaln_kernel(decaf_adaln_t p) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * ROWS_PER_CTA + warp;
    if (row >= p.rows) return;
    const int C = p.C;
    float q[VEC], sc[VEC], sh[VEC];
    load_row<VEC>(p.q + row * C, lane, q);
    const TA *ss = reinterpret_cast<const TA *>(p.ss) + row * 2 * C;
    load_row<VEC>(ss, lane, sc);
    load_row<VEC>(ss + C, lane, sh);
    warp_layernorm<VEC>(q, C, p.eps);
    const float m = p.rowmask ? (float)p.rowmask[row] : 1.f;
#pragma unroll
    for (int i = 0; i < VEC; i++) q[i] = (q[i] * sc[i] + sh[i]) * m;
    store_row<VEC>(p.out_q + row * C, lane, q);
    warp_layernorm<VEC>(q, C, p.eps);
    float w[VEC], b[VEC];
    load_row<VEC>(p.w_ffn, lane, w);
    load_row<VEC>(p.b_ffn, lane, b);
#pragma unroll
    for (int i = 0; i < VEC; i++) q[i] = q[i] * w[i] + b[i];
    store_row<VEC>(reinterpret_cast<TA *>(p.out_act) + row * C, lane, q);
}

// ------------------------------------------------------------------------------- head output conv
__device__ __forceinline__ int level_of_row(const decaf_levels_t &lv, int r) {
    for (int l = 0; l < lv.n_levels; l++)
        if (r >= lv.off[l] && r < lv.off[l] + lv.len[l]) return l;
    return -1;
}

template <int VEC, typename TA, int NOUT>
__global__ void __launch_bounds__(32 * ROWS_PER_CTA)
head_out_kernel(const TA *__restrict__ x, int64_t ldx, int rows_total, int C, const float *__restrict__ w,
                const float *__restrict__ bias, int mode, const float *__restrict__ level_scale,
                decaf_levels_t lv, float *__restrict__ out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * ROWS_PER_CTA + warp;
    if (row >= rows_total) return;
    const int level = level_of_row(lv, (int)(row % lv.Pp));
    if (level < 0) {
        if (lane < NOUT) out[row * NOUT + lane] = 0.f;
        return;
    }
    float acc[NOUT];
#pragma unroll
    for (int o = 0; o < NOUT; o++) acc[o] = 0.f;
#pragma unroll
    for (int tap = 0; tap < 3; tap++) {
        float xv[VEC];
        load_row<VEC>(x + (row + tap - 1) * ldx, lane, xv);
#pragma unroll
        for (int o = 0; o < NOUT; o++) {
            float wv[VEC];
            load_row<VEC>(w + ((int64_t)o * 3 + tap) * C, lane, wv);
#pragma unroll
            for (int i = 0; i < VEC; i++) acc[o] = fmaf(xv[i], wv[i], acc[o]);
        }
    }
#pragma unroll
    for (int o = 0; o < NOUT; o++) {
        float v = warp_sum(acc[o]) + bias[o];
        if (mode == 1) v = fmaxf(level_scale[level] * v, 0.f);
        if (lane == 0) out[row * NOUT + o] = v;
    }
}

// ------------------------------------------------------------------------------- masks / text glue
__global__ void build_masks_kernel(const uint8_t *__restrict__ mask0, int64_t m0_seq_stride,
                                   uint8_t *__restrict__ hmask, decaf_levels_t lv, int n_query) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n_query * lv.Pp) return;
    const int q = (int)(i / lv.Pp), r = (int)(i % lv.Pp);
    const int l = level_of_row(lv, r);
    uint8_t m = 0;
    if (l >= 0) {
        const int t = r - lv.off[l];
        // mask_l[t] = mask_{l-1}[2t] = ... = mask_0[t << l]
        m = mask0[(int64_t)q * m0_seq_stride + ((int64_t)t << l)];
    }
    hmask[i] = m;
}

__global__ void text_prep_kernel(float *__restrict__ x, int n_query, int L1, int C,
                                 const float *__restrict__ bkgd, const float *__restrict__ pe,
                                 const int32_t *__restrict__ len) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n_query * L1 * C) return;
    const int c = (int)(i % C);
    const int r = (int)((i / C) % L1);
    const int q = (int)(i / ((int64_t)C * L1));
    if (r == 0) {
        x[i] = bkgd[c];
    } else if (pe && (r - 1) < len[q]) {
        x[i] += pe[(int64_t)(r - 1) * C + c];
    }
}

}  // namespace decaf

using namespace decaf;

extern "C" int decaf_layernorm(const decaf_layernorm_t *pp, void *stream) {
    decaf_layernorm_t p = *pp;
    DECAF_CHECK(p.x && (p.out_f32 || p.out_act), "decaf_layernorm: null pointers");
    DECAF_CHECK(p.C % 32 == 0, "decaf_layernorm: C %% 32 != 0 (%d)", p.C);
    DECAF_CHECK((p.w == nullptr) == (p.b == nullptr), "decaf_layernorm: w/b must both be set or both null");
    if (!p.x_seq_stride) p.x_seq_stride = p.rows_per_seq;
    if (!p.m_seq_stride) p.m_seq_stride = p.rows_per_seq;
    if (!p.o_seq_stride) p.o_seq_stride = p.rows_per_seq;
    if (!p.o2_seq_stride) p.o2_seq_stride = p.rows_per_seq;
    const int64_t rows = (int64_t)p.n_seq * p.rows_per_seq;
    if (rows == 0) return 0;
    const int grid = cdiv(rows, ROWS_PER_CTA);
    cudaStream_t st = as_stream(stream);
    if (p.dtype == DECAF_BF16 && p.out_act) {
        DECAF_DISPATCH_VEC(p.C, (layernorm_kernel<VEC, bf16><<<grid, 32 * ROWS_PER_CTA, 0, st>>>(p)));
    } else {
        DECAF_DISPATCH_VEC(p.C, (layernorm_kernel<VEC, float><<<grid, 32 * ROWS_PER_CTA, 0, st>>>(p)));
    }
    DECAF_LAUNCH_CHECK();
    return 0;
}

extern "C" int decaf_preattn(const decaf_preattn_t *pp, void *stream) {
    decaf_preattn_t p = *pp;
    DECAF_CHECK(p.x && p.mask_in && p.out_act, "decaf_preattn: null pointers");
    DECAF_CHECK(p.stride == 1 || p.stride == 2, "decaf_preattn: stride must be 1 or 2");
    DECAF_CHECK(p.T_in % p.stride == 0, "decaf_preattn: T_in %% stride != 0");
    DECAF_CHECK(p.n_branch >= 1 && p.n_branch <= 3, "decaf_preattn: n_branch in 1..3");
    DECAF_CHECK(p.C % 32 == 0, "decaf_preattn: C %% 32 != 0");
    DECAF_CHECK(p.stride == 2 || (!p.skip_out && !p.mask_out), "decaf_preattn: skip/mask outputs are stride-2 only");
    if (!p.mi_seq_stride) p.mi_seq_stride = p.T_in;
    const int T_out = p.T_in / p.stride;
    if (!p.mo_seq_stride) p.mo_seq_stride = T_out;
    const int64_t rows = (int64_t)p.n_seq * T_out;
    if (rows == 0) return 0;
    const int grid = cdiv(rows, ROWS_PER_CTA);
    cudaStream_t st = as_stream(stream);
    if (p.dtype == DECAF_BF16) {
        DECAF_DISPATCH_VEC(p.C, (preattn_kernel<VEC, bf16><<<grid, 32 * ROWS_PER_CTA, 0, st>>>(p, T_out)));
    } else {
        DECAF_DISPATCH_VEC(p.C, (preattn_kernel<VEC, float><<<grid, 32 * ROWS_PER_CTA, 0, st>>>(p, T_out)));
    }
    DECAF_LAUNCH_CHECK();
    return 0;
}

extern "C" int decaf_adaln(const decaf_adaln_t *pp, void *stream) {
    decaf_adaln_t p = *pp;
    DECAF_CHECK(p.q && p.ss && p.out_q && p.out_act && p.w_ffn && p.b_ffn, "decaf_adaln: null pointers");
    DECAF_CHECK(p.ss_dtype == p.dtype, "decaf_adaln: ss dtype must equal the act dtype");
    DECAF_CHECK(p.C % 32 == 0, "decaf_adaln: C %% 32 != 0");
    if (p.rows == 0) return 0;
    const int grid = cdiv(p.rows, ROWS_PER_CTA);
    cudaStream_t st = as_stream(stream);
    if (p.dtype == DECAF_BF16) {
        DECAF_DISPATCH_VEC(p.C, (adaln_kernel<VEC, bf16><<<grid, 32 * ROWS_PER_CTA, 0, st>>>(p)));
    } else {
        DECAF_DISPATCH_VEC(p.C, (adaln_kernel<VEC, float><<<grid, 32 * ROWS_PER_CTA, 0, st>>>(p)));
    }
    DECAF_LAUNCH_CHECK();
    return 0;
}

extern "C" int decaf_head_out(const void *x, int32_t dtype, int64_t ldx, int32_t rows_total, int32_t C,
                              const float *w, const float *bias, int32_t n_out, int32_t mode,
                              const float *level_scale, const decaf_levels_t *lv, float *out, void *stream) {
    DECAF_CHECK(x && w && bias && out && lv, "decaf_head_out: null pointers");
    DECAF_CHECK(n_out == 1 || n_out == 2, "decaf_head_out: n_out must be 1 or 2");
    DECAF_CHECK(mode == 0 || level_scale, "decaf_head_out: mode 1 needs level_scale");
    DECAF_CHECK(C % 32 == 0, "decaf_head_out: C %% 32 != 0");
    DECAF_CHECK(rows_total % lv->Pp == 0, "decaf_head_out: rows_total %% Pp != 0");
    if (rows_total == 0) return 0;
    const int grid = cdiv(rows_total, ROWS_PER_CTA);
    cudaStream_t st = as_stream(stream);
#define HO_LAUNCH(TA, NO)                                                                                    \
    DECAF_DISPATCH_VEC(C, (head_out_kernel<VEC, TA, NO><<<grid, 32 * ROWS_PER_CTA, 0, st>>>(                 \
                              reinterpret_cast<const TA *>(x), ldx, rows_total, C, w, bias, mode, level_scale, *lv, out)))
    if (dtype == DECAF_BF16) {
        if (n_out == 1) { HO_LAUNCH(bf16, 1); } else { HO_LAUNCH(bf16, 2); }
    } else {
        if (n_out == 1) { HO_LAUNCH(float, 1); } else { HO_LAUNCH(float, 2); }
    }
#undef HO_LAUNCH
    DECAF_LAUNCH_CHECK();
    return 0;
}

extern "C" int decaf_build_masks(const uint8_t *mask0, int64_t m0_seq_stride, uint8_t *hmask,
                                 const decaf_levels_t *lv, int32_t n_query, void *stream) {
    DECAF_CHECK(mask0 && hmask && lv, "decaf_build_masks: null pointers");
    const int64_t n = (int64_t)n_query * lv->Pp;
    if (n == 0) return 0;
    build_masks_kernel<<<cdiv(n, 256), 256, 0, as_stream(stream)>>>(mask0, m0_seq_stride, hmask, *lv, n_query);
    DECAF_LAUNCH_CHECK();
    return 0;
}

extern "C" int decaf_text_prep(float *x, int32_t n_query, int32_t L1, int32_t C, const float *bkgd,
                               const float *pe, const int32_t *len, void *stream) {
    DECAF_CHECK(x && bkgd && len, "decaf_text_prep: null pointers");
    const int64_t n = (int64_t)n_query * L1 * C;
    if (n == 0) return 0;
    text_prep_kernel<<<cdiv(n, 256), 256, 0, as_stream(stream)>>>(x, n_query, L1, C, bkgd, pe, len);
    DECAF_LAUNCH_CHECK();
    return 0;
}

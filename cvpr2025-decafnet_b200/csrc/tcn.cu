// Iterative refinement stage between the first and second heads (reference:
// libs/modeling/model.py:449-467 and libs/modeling/tcn.py): nearest-expand the per-level
// logits to level-0 resolution, a dilated residual TCN over R = 32 channels, then a masked
// max-pool pyramid written next to the FPN features in the concatenated head-input buffer.
// 0.5 % of the FLOPs: one thread per time step, weights broadcast from shared memory.
#include "common.cuh"

namespace decaf {

constexpr int TCN_R = 32;
constexpr int TCN_THREADS = 128;

__global__ void __launch_bounds__(TCN_THREADS)
tcn_in_kernel(const float *__restrict__ logits1, const uint8_t *__restrict__ hmask, decaf_levels_t lv,
              const float *__restrict__ w_in, const float *__restrict__ b_in, float *__restrict__ r0, int n_query) {
    __shared__ float w[TCN_R * DECAF_MAX_LEVELS];
    __shared__ float b[TCN_R];
    const int L = lv.n_levels, T0 = lv.len[0];
    for (int i = threadIdx.x; i < TCN_R * L; i += blockDim.x) w[i] = w_in[i];
    if (threadIdx.x < TCN_R) b[threadIdx.x] = b_in[threadIdx.x];
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n_query * T0) return;
    const int q = (int)(i / T0), t = (int)(i % T0);
    const float *lg = logits1 + (int64_t)q * lv.Pp;
    const float m0 = (float)hmask[(int64_t)q * lv.Pp + lv.off[0] + t];
    float s[DECAF_MAX_LEVELS];
#pragma unroll
    for (int l = 0; l < DECAF_MAX_LEVELS; l++) {
        if (l < L) {
            // F.interpolate(nearest) from len[l] to T0 = len[l] << l: src = t >> l; level 0 is
            // used raw, the others are multiplied by the level-0 mask (model.py:449-454)
            const float v = lg[lv.off[l] + min(t >> l, lv.len[l] - 1)];
            s[l] = l == 0 ? v : v * m0;
        } else {
            s[l] = 0.f;
        }
    }
    float4 *out = reinterpret_cast<float4 *>(r0 + i * TCN_R);
#pragma unroll
    for (int c4 = 0; c4 < TCN_R / 4; c4++) {
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int c = c4 * 4 + j;
            float acc = b[c];
            for (int l = 0; l < L; l++) acc = fmaf(w[c * L + l], s[l], acc);
            o[j] = acc;
        }
        out[c4] = make_float4(o[0], o[1], o[2], o[3]);
    }
}

__global__ void __launch_bounds__(TCN_THREADS)
tcn_layer_kernel(const float *__restrict__ r_in, float *__restrict__ r_out, const uint8_t *__restrict__ mask0,
                 int64_t m_seq_stride, const float *__restrict__ wd, const float *__restrict__ bd,
                 const float *__restrict__ w1, const float *__restrict__ b1, const float *__restrict__ ln_w,
                 const float *__restrict__ ln_b, float eps, int dil, int n_query, int T) {
    __shared__ __align__(16) float s_wd[3][TCN_R][TCN_R];   // [k][cin][cout]
    __shared__ __align__(16) float s_w1[TCN_R][TCN_R];      // [cin][cout]
    __shared__ float s_bd[TCN_R], s_b1[TCN_R], s_lw[TCN_R], s_lb[TCN_R];
    for (int i = threadIdx.x; i < TCN_R * TCN_R * 3; i += blockDim.x) {
        const int k = i % 3, cin = (i / 3) % TCN_R, cout = i / (3 * TCN_R);
        s_wd[k][cin][cout] = wd[i];                          // wd: (cout, cin, k)
    }
    for (int i = threadIdx.x; i < TCN_R * TCN_R; i += blockDim.x) {
        const int cin = i % TCN_R, cout = i / TCN_R;
        s_w1[cin][cout] = w1[i];                             // w1: (cout, cin)
    }
    if (threadIdx.x < TCN_R) {
        s_bd[threadIdx.x] = bd[threadIdx.x]; s_b1[threadIdx.x] = b1[threadIdx.x];
        s_lw[threadIdx.x] = ln_w[threadIdx.x]; s_lb[threadIdx.x] = ln_b[threadIdx.x];
    }
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n_query * T) return;
    const int q = (int)(i / T), t = (int)(i % T);
    const float *base = r_in + (int64_t)q * T * TCN_R;

    float h[TCN_R];
#pragma unroll
    for (int c = 0; c < TCN_R; c++) h[c] = s_bd[c];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int ts = t + (k - 1) * dil;
        if (ts < 0 || ts >= T) continue;                     // zero padding
        const float4 *xr = reinterpret_cast<const float4 *>(base + (int64_t)ts * TCN_R);
#pragma unroll
        for (int c4 = 0; c4 < TCN_R / 4; c4++) {
            const float4 xv = xr[c4];
            const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float4 *wrow = reinterpret_cast<const float4 *>(&s_wd[k][c4 * 4 + j][0]);
#pragma unroll
                for (int o4 = 0; o4 < TCN_R / 4; o4++) {
                    const float4 wv = wrow[o4];
                    h[o4 * 4 + 0] = fmaf(xs[j], wv.x, h[o4 * 4 + 0]);
                    h[o4 * 4 + 1] = fmaf(xs[j], wv.y, h[o4 * 4 + 1]);
                    h[o4 * 4 + 2] = fmaf(xs[j], wv.z, h[o4 * 4 + 2]);
                    h[o4 * 4 + 3] = fmaf(xs[j], wv.w, h[o4 * 4 + 3]);
                }
            }
        }
    }
    float y[TCN_R];
    {
        const float4 *xr = reinterpret_cast<const float4 *>(base + (int64_t)t * TCN_R);
#pragma unroll
        for (int c4 = 0; c4 < TCN_R / 4; c4++) {
            const float4 xv = xr[c4];
            y[c4 * 4 + 0] = xv.x + s_b1[c4 * 4 + 0]; y[c4 * 4 + 1] = xv.y + s_b1[c4 * 4 + 1];
            y[c4 * 4 + 2] = xv.z + s_b1[c4 * 4 + 2]; y[c4 * 4 + 3] = xv.w + s_b1[c4 * 4 + 3];
        }
    }
#pragma unroll
    for (int c = 0; c < TCN_R; c++) {
        const float hv = fmaxf(h[c], 0.f);
        const float4 *wrow = reinterpret_cast<const float4 *>(&s_w1[c][0]);
#pragma unroll
        for (int o4 = 0; o4 < TCN_R / 4; o4++) {
            const float4 wv = wrow[o4];
            y[o4 * 4 + 0] = fmaf(hv, wv.x, y[o4 * 4 + 0]);
            y[o4 * 4 + 1] = fmaf(hv, wv.y, y[o4 * 4 + 1]);
            y[o4 * 4 + 2] = fmaf(hv, wv.z, y[o4 * 4 + 2]);
            y[o4 * 4 + 3] = fmaf(hv, wv.w, y[o4 * 4 + 3]);
        }
    }
    const float m = (float)mask0[(int64_t)q * m_seq_stride + t];
    float mean = 0.f;
#pragma unroll
    for (int c = 0; c < TCN_R; c++) { y[c] *= m; mean += y[c]; }
    mean *= (1.0f / TCN_R);
    float var = 0.f;
#pragma unroll
    for (int c = 0; c < TCN_R; c++) { y[c] -= mean; var = fmaf(y[c], y[c], var); }
    var *= (1.0f / TCN_R);
    const float rs = 1.0f / sqrtf(var + eps);
    float4 *out = reinterpret_cast<float4 *>(r_out + i * TCN_R);
#pragma unroll
    for (int c4 = 0; c4 < TCN_R / 4; c4++) {
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; j++) o[j] = y[c4 * 4 + j] * rs * s_lw[c4 * 4 + j] + s_lb[c4 * 4 + j];
        out[c4] = make_float4(o[0], o[1], o[2], o[3]);
    }
}

template <typename TA>
__global__ void __launch_bounds__(TCN_THREADS)
tcn_out_kernel(const float *__restrict__ r_in, const uint8_t *__restrict__ mask0, int64_t m_seq_stride,
               const float *__restrict__ w_out, const float *__restrict__ b_out, TA *__restrict__ cat,
               int64_t ldc, int col0, decaf_levels_t lv, int n_query) {
    __shared__ __align__(16) float s_w[TCN_R][TCN_R];       // [cin][cout]
    __shared__ float s_b[TCN_R];
    for (int i = threadIdx.x; i < TCN_R * TCN_R; i += blockDim.x) s_w[i % TCN_R][i / TCN_R] = w_out[i];
    if (threadIdx.x < TCN_R) s_b[threadIdx.x] = b_out[threadIdx.x];
    __syncthreads();
    const int T = lv.len[0];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n_query * T) return;
    const int q = (int)(i / T), t = (int)(i % T);
    float y[TCN_R];
#pragma unroll
    for (int c = 0; c < TCN_R; c++) y[c] = s_b[c];
    const float4 *xr = reinterpret_cast<const float4 *>(r_in + i * TCN_R);
#pragma unroll
    for (int c4 = 0; c4 < TCN_R / 4; c4++) {
        const float4 xv = xr[c4];
        const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float4 *wrow = reinterpret_cast<const float4 *>(&s_w[c4 * 4 + j][0]);
#pragma unroll
            for (int o4 = 0; o4 < TCN_R / 4; o4++) {
                const float4 wv = wrow[o4];
                y[o4 * 4 + 0] = fmaf(xs[j], wv.x, y[o4 * 4 + 0]);
                y[o4 * 4 + 1] = fmaf(xs[j], wv.y, y[o4 * 4 + 1]);
                y[o4 * 4 + 2] = fmaf(xs[j], wv.z, y[o4 * 4 + 2]);
                y[o4 * 4 + 3] = fmaf(xs[j], wv.w, y[o4 * 4 + 3]);
            }
        }
    }
    const float m = (float)mask0[(int64_t)q * m_seq_stride + t];
    TA *dst = cat + ((int64_t)q * lv.Pp + lv.off[0] + t) * ldc + col0;
#pragma unroll
    for (int c = 0; c < TCN_R; c++) dst[c] = from_f32<TA>(y[c] * m);
}

// level `level` (>= 1) columns <- masked max-pool(k=3, s=2, pad=1) of level-1 columns, with the
// level-1 mask (libs/modeling/blocks.py:31-47: max over valid taps, 0 when none is valid).
// The result is stored multiplied by this level's own (nearest-down-sampled) mask: the reference
// keeps the pooled value at positions where only the pooled mask is set, but both of its
// consumers drop it again (MaskedConv1D's x * mask in the heads, and the next pooling level,
// which excludes positions masked at this level) — and the head GEMMs here read this buffer
// without a mask multiply.
template <typename TA>
__global__ void refine_pool_kernel(TA *__restrict__ cat, int64_t ldc, int col0, int R,
                                   const uint8_t *__restrict__ hmask, decaf_levels_t lv, int level, int n_query) {
    const int len = lv.len[level], plen = lv.len[level - 1];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n_query * len * R) return;
    const int c = (int)(i % R);
    const int t = (int)((i / R) % len);
    const int q = (int)(i / ((int64_t)R * len));
    const int64_t prow0 = (int64_t)q * lv.Pp + lv.off[level - 1];
    float best = -INFINITY;
    bool any = false;
#pragma unroll
    for (int j = -1; j <= 1; j++) {
        const int ts = 2 * t + j;
        if (ts < 0 || ts >= plen) continue;
        if (!hmask[prow0 + ts]) continue;
        best = fmaxf(best, to_f32<TA>(cat[(prow0 + ts) * ldc + col0 + c]));
        any = true;
    }
    const int64_t orow = (int64_t)q * lv.Pp + lv.off[level] + t;
    cat[orow * ldc + col0 + c] = from_f32<TA>((any && hmask[orow]) ? best : 0.f);
}

}  // namespace decaf

using namespace decaf;

extern "C" int decaf_tcn_in(const float *logits1, const uint8_t *hmask, const decaf_levels_t *lv,
                            const float *w_in, const float *b_in, int32_t R, float *r0, int32_t n_query,
                            void *stream) {
    DECAF_CHECK(logits1 && hmask && lv && w_in && b_in && r0, "decaf_tcn_in: null pointers");
    DECAF_CHECK(R == TCN_R, "decaf_tcn: refine width must be %d (got %d)", TCN_R, R);
    DECAF_CHECK(lv->n_levels <= DECAF_MAX_LEVELS, "decaf_tcn_in: too many levels");
    const int64_t n = (int64_t)n_query * lv->len[0];
    if (n == 0) return 0;
    tcn_in_kernel<<<cdiv(n, TCN_THREADS), TCN_THREADS, 0, as_stream(stream)>>>(logits1, hmask, *lv, w_in, b_in, r0, n_query);
    DECAF_LAUNCH_CHECK();
    return 0;
}

extern "C" int decaf_tcn_layer(const float *r_in, float *r_out, const uint8_t *mask0, int64_t m_seq_stride,
                               const float *wd, const float *bd, const float *w1, const float *b1,
                               const float *ln_w, const float *ln_b, float eps, int32_t R, int32_t dil,
                               int32_t n_query, int32_t T, void *stream) {
    DECAF_CHECK(r_in && r_out && mask0 && wd && bd && w1 && b1 && ln_w && ln_b, "decaf_tcn_layer: null pointers");
    DECAF_CHECK(r_in != r_out, "decaf_tcn_layer: in-place not supported (dilated taps read neighbours)");
    DECAF_CHECK(R == TCN_R, "decaf_tcn: refine width must be %d (got %d)", TCN_R, R);
    if (!m_seq_stride) m_seq_stride = T;
    const int64_t n = (int64_t)n_query * T;
    if (n == 0) return 0;
    tcn_layer_kernel<<<cdiv(n, TCN_THREADS), TCN_THREADS, 0, as_stream(stream)>>>(
        r_in, r_out, mask0, m_seq_stride, wd, bd, w1, b1, ln_w, ln_b, eps, dil, n_query, T);
    DECAF_LAUNCH_CHECK();
    return 0;
}

extern "C" int decaf_tcn_out(const float *r_in, const uint8_t *mask0, int64_t m_seq_stride, const float *w_out,
                             const float *b_out, int32_t R, void *cat, int32_t dtype, int64_t ldc, int32_t col0,
                             const decaf_levels_t *lv, int32_t n_query, void *stream) {
    DECAF_CHECK(r_in && mask0 && w_out && b_out && cat && lv, "decaf_tcn_out: null pointers");
    DECAF_CHECK(R == TCN_R, "decaf_tcn: refine width must be %d (got %d)", TCN_R, R);
    if (!m_seq_stride) m_seq_stride = lv->len[0];
    const int64_t n = (int64_t)n_query * lv->len[0];
    if (n == 0) return 0;
    cudaStream_t st = as_stream(stream);
    if (dtype == DECAF_BF16)
        tcn_out_kernel<bf16><<<cdiv(n, TCN_THREADS), TCN_THREADS, 0, st>>>(r_in, mask0, m_seq_stride, w_out, b_out,
                                                                          (bf16 *)cat, ldc, col0, *lv, n_query);
    else
        tcn_out_kernel<float><<<cdiv(n, TCN_THREADS), TCN_THREADS, 0, st>>>(r_in, mask0, m_seq_stride, w_out, b_out,
                                                                           (float *)cat, ldc, col0, *lv, n_query);
    DECAF_LAUNCH_CHECK();
    return 0;
}

extern "C" int decaf_refine_pool(void *cat, int32_t dtype, int64_t ldc, int32_t col0, int32_t R,
                                 const uint8_t *hmask, const decaf_levels_t *lv, int32_t level, int32_t n_query,
                                 void *stream) {
    DECAF_CHECK(cat && hmask && lv, "decaf_refine_pool: null pointers");
    DECAF_CHECK(level >= 1 && level < lv->n_levels, "decaf_refine_pool: bad level %d", level);
    const int64_t n = (int64_t)n_query * lv->len[level] * R;
    if (n == 0) return 0;
    cudaStream_t st = as_stream(stream);
    if (dtype == DECAF_BF16)
        refine_pool_kernel<bf16><<<cdiv(n, 256), 256, 0, st>>>((bf16 *)cat, ldc, col0, R, hmask, *lv, level, n_query);
    else
        refine_pool_kernel<float><<<cdiv(n, 256), 256, 0, st>>>((float *)cat, ldc, col0, R, hmask, *lv, level, n_query);
    DECAF_LAUNCH_CHECK();
    return 0;
}

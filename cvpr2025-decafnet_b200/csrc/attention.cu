// Banded local-window self-attention and short-key global (cross) attention.
// One warp per query time step; lane = head * (32 / n_heads) + sub, each lane owns C/32
// contiguous channels of its head, so a warp reads whole 128B-multiple rows (coalesced) and a
// score needs log2(32 / n_heads) shuffles.  Softmax is computed online in fp32.
// The reference materialises (2s x 2s) chunk products and a second "mask matmul"
// (libs/modeling/blocks.py:224-325); here the band is evaluated directly.
#include "common.cuh"

namespace decaf {

constexpr int AROWS = 8;

template <int LPH>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = LPH / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One chunk of KC keys of the online softmax: the KC score reductions, exponentials and row loads are
// independent, so their latencies overlap (the former one-key-at-a-time loop was a single dependent chain).
// kptr(j) / vptr(j) give the row pointers of key j of the chunk, valid(j) whether it exists, bias(j) an
// additive score term (the reference's -1e4 for masked keys of the local branch).
constexpr int KC = 5;

template <int VEC, int LPH, typename TK, typename FK, typename FV, typename FOK, typename FB>
__device__ __forceinline__ void attn_chunk(const float (&qv)[VEC], float (&acc)[VEC], float &m, float &l, int lane,
                                           FK kptr, FV vptr, FOK valid, FB bias) {
    float d[KC];
    bool ok[KC];
#pragma unroll
    for (int j = 0; j < KC; j++) {
        ok[j] = valid(j);
        d[j] = 0.f;
        if (ok[j]) {
            float kv[VEC];
            load_row<VEC>(kptr(j), lane, kv);
#pragma unroll
            for (int i = 0; i < VEC; i++) d[j] = fmaf(qv[i], kv[i], d[j]);
        }
    }
#pragma unroll
    for (int o = LPH / 2; o > 0; o >>= 1) {
#pragma unroll
        for (int j = 0; j < KC; j++) d[j] += __shfl_xor_sync(0xffffffffu, d[j], o);
    }
    float mn = m;
#pragma unroll
    for (int j = 0; j < KC; j++) {
        d[j] = ok[j] ? d[j] + bias(j) : -INFINITY;
        mn = fmaxf(mn, d[j]);
    }
    if (mn == -INFINITY) return;                     // no key in this chunk and none before
    const float corr = __expf(m - mn);               // m = -inf on the first chunk -> 0
    float pj[KC], ps = 0.f;
#pragma unroll
    for (int j = 0; j < KC; j++) { pj[j] = __expf(d[j] - mn); ps += pj[j]; }
    l = l * corr + ps;
#pragma unroll
    for (int i = 0; i < VEC; i++) acc[i] *= corr;
#pragma unroll
    for (int j = 0; j < KC; j++) {
        if (ok[j]) {
            float vv[VEC];
            load_row<VEC>(vptr(j), lane, vv);
#pragma unroll
            for (int i = 0; i < VEC; i++) acc[i] = fmaf(pj[j], vv[i], acc[i]);
        }
    }
    m = mn;
}

template <int VEC, int LPH, typename TA>
__global__ void __launch_bounds__(32 * AROWS)
local_attn_kernel(const TA *__restrict__ q, const TA *__restrict__ k, const TA *__restrict__ v,
                  TA *__restrict__ out, int n_seq, int T, int C, int window, float scale2,
                  const uint8_t *__restrict__ mask, int64_t m_seq_stride) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * AROWS + warp;
    if (row >= (int64_t)n_seq * T) return;
    const int seq = (int)(row / T), t = (int)(row % T);
    const uint8_t *mrow = mask + (int64_t)seq * m_seq_stride;
    float acc[VEC];
#pragma unroll
    for (int i = 0; i < VEC; i++) acc[i] = 0.f;
    if (mrow[t]) {
        float qv[VEC];
        load_row<VEC>(q + row * C, lane, qv);
#pragma unroll
        for (int i = 0; i < VEC; i++) qv[i] *= scale2;
        const int s = window / 2;
        float m = -INFINITY, l = 0.f;
        const int j0 = max(t - s, 0), j1 = min(t + s, T - 1);
        const TA *kb = k + (int64_t)seq * T * C, *vb = v + (int64_t)seq * T * C;
        for (int base = j0; base <= j1; base += KC) {
            attn_chunk<VEC, LPH, TA>(
                qv, acc, m, l, lane,
                [&](int j) { return kb + (int64_t)(base + j) * C; },
                [&](int j) { return vb + (int64_t)(base + j) * C; },
                [&](int j) { return base + j <= j1; },
                [&](int j) { return mrow[base + j] ? 0.f : -1e4f; });
        }
        const float inv = 1.0f / l;
#pragma unroll
        for (int i = 0; i < VEC; i++) acc[i] *= inv;
    }
    store_row<VEC>(out + row * C, lane, acc);
}

template <int VEC, int LPH, typename TQ, typename TO>
__global__ void __launch_bounds__(32 * AROWS)
xattn_kernel(const TQ *__restrict__ q, const float *__restrict__ k, const float *__restrict__ v,
             TO *__restrict__ out, int n_seq, int Tq, int Lk, int C, float scale2,
             const int32_t *__restrict__ kv_len) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * AROWS + warp;
    if (row >= (int64_t)n_seq * Tq) return;
    const int seq = (int)(row / Tq);
    const int n_kv = min(kv_len[seq], Lk);
    float acc[VEC];
#pragma unroll
    for (int i = 0; i < VEC; i++) acc[i] = 0.f;
    float qv[VEC];
    load_row<VEC>(q + row * C, lane, qv);
#pragma unroll
    for (int i = 0; i < VEC; i++) qv[i] *= scale2;
    float m = -INFINITY, l = 0.f;
    const float *kb = k + (int64_t)seq * Lk * C, *vb = v + (int64_t)seq * Lk * C;
    for (int base = 0; base < n_kv; base += KC) {
        attn_chunk<VEC, LPH, float>(
            qv, acc, m, l, lane,
            [&](int j) { return kb + (int64_t)(base + j) * C; },
            [&](int j) { return vb + (int64_t)(base + j) * C; },
            [&](int j) { return base + j < n_kv; },
            [&](int) { return 0.f; });
    }
    const float inv = n_kv > 0 ? 1.0f / l : 0.f;
#pragma unroll
    for (int i = 0; i < VEC; i++) acc[i] *= inv;
    store_row<VEC>(out + row * C, lane, acc);
}

}  // namespace decaf

using namespace decaf;

#define DECAF_DISPATCH_LPH(n_heads, ...)                                              \
    do {                                                                              \
        switch (n_heads) {                                                            \
            case 1:  { constexpr int LPH = 32; __VA_ARGS__; } break;                  \
            case 2:  { constexpr int LPH = 16; __VA_ARGS__; } break;                  \
            case 4:  { constexpr int LPH = 8;  __VA_ARGS__; } break;                  \
            case 8:  { constexpr int LPH = 4;  __VA_ARGS__; } break;                  \
            default:                                                                  \
                decaf::set_error("n_heads must be 1,2,4 or 8 (got %d)", (int)(n_heads)); \
                return 1;                                                             \
        }                                                                             \
    } while (0)

extern "C" int decaf_local_attn(const void *q, const void *k, const void *v, void *out, int32_t dtype,
                                int32_t n_seq, int32_t T, int32_t C, int32_t n_heads, int32_t window,
                                const uint8_t *mask, int64_t m_seq_stride, void *stream) {
    DECAF_CHECK(q && k && v && out && mask, "decaf_local_attn: null pointers");
    DECAF_CHECK(C % 32 == 0 && C % n_heads == 0, "decaf_local_attn: bad C/n_heads");
    DECAF_CHECK(window > 0 && window % 2 == 1, "decaf_local_attn: window must be odd and > 0");
    if (!m_seq_stride) m_seq_stride = T;
    const int64_t rows = (int64_t)n_seq * T;
    if (rows == 0) return 0;
    const int grid = cdiv(rows, AROWS);
    const float scale2 = 1.0f / sqrtf((float)(C / n_heads));
    cudaStream_t st = as_stream(stream);
    if (dtype == DECAF_BF16) {
        DECAF_DISPATCH_LPH(n_heads, DECAF_DISPATCH_VEC_ATTN(C, (local_attn_kernel<VEC, LPH, bf16><<<grid, 32 * AROWS, 0, st>>>(
            (const bf16 *)q, (const bf16 *)k, (const bf16 *)v, (bf16 *)out, n_seq, T, C, window, scale2, mask, m_seq_stride))));
    } else {
        DECAF_DISPATCH_LPH(n_heads, DECAF_DISPATCH_VEC_ATTN(C, (local_attn_kernel<VEC, LPH, float><<<grid, 32 * AROWS, 0, st>>>(
            (const float *)q, (const float *)k, (const float *)v, (float *)out, n_seq, T, C, window, scale2, mask, m_seq_stride))));
    }
    DECAF_LAUNCH_CHECK();
    return 0;
}

extern "C" int decaf_xattn(const void *q, int32_t q_dtype, const float *k, const float *v, void *out,
                           int32_t out_dtype, int32_t n_seq, int32_t Tq, int32_t Lk, int32_t C,
                           int32_t n_heads, const int32_t *kv_len, void *stream) {
    DECAF_CHECK(q && k && v && out && kv_len, "decaf_xattn: null pointers");
    DECAF_CHECK(C % 32 == 0 && C % n_heads == 0, "decaf_xattn: bad C/n_heads");
    DECAF_CHECK(q_dtype == out_dtype, "decaf_xattn: q and out dtypes must match");
    const int64_t rows = (int64_t)n_seq * Tq;
    if (rows == 0) return 0;
    const int grid = cdiv(rows, AROWS);
    const float scale2 = 1.0f / sqrtf((float)(C / n_heads));
    cudaStream_t st = as_stream(stream);
    if (q_dtype == DECAF_BF16) {
        DECAF_DISPATCH_LPH(n_heads, DECAF_DISPATCH_VEC_ATTN(C, (xattn_kernel<VEC, LPH, bf16, bf16><<<grid, 32 * AROWS, 0, st>>>(
            (const bf16 *)q, k, v, (bf16 *)out, n_seq, Tq, Lk, C, scale2, kv_len))));
    } else {
        DECAF_DISPATCH_LPH(n_heads, DECAF_DISPATCH_VEC_ATTN(C, (xattn_kernel<VEC, LPH, float, float><<<grid, 32 * AROWS, 0, st>>>(
            (const float *)q, k, v, (float *)out, n_seq, Tq, Lk, C, scale2, kv_len))));
    }
    DECAF_LAUNCH_CHECK();
    return 0;
}

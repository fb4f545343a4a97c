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
    warp_layernorm<VEC>(v, 32 * VEC, p.eps);               // 32 * VEC == p.C (dispatch)
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
// Each warp walks a strip of PRE_STRIP consecutive output rows with a 3-row sliding window in
// registers, so every input row is loaded and LayerNorm-ed once (plus a 2-row halo per strip);
// the depthwise taps and branch affines are staged once per CTA in shared memory as
// [branch][tap|w|b][C] (the ABI keeps the reference's (C, 1, 3) weight layout); the NB branch
// LayerNorms are reduced together so their shuffle chains overlap.
constexpr int PRE_STRIP = 8;
constexpr int PRE_DEPTH = 4;     // input rows in flight per warp

template <int NB, int VEC>
__device__ __forceinline__ void warp_layernorm_multi(float (&v)[NB][VEC], int C, float eps) {
    float s[NB];
#pragma unroll
    for (int b = 0; b < NB; b++) s[b] = pk_sum<VEC>(v[b]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int b = 0; b < NB; b++) s[b] += __shfl_xor_sync(0xffffffffu, s[b], o);
    }
    float ss[NB];
#pragma unroll
    for (int b = 0; b < NB; b++) ss[b] = pk_center_sq<VEC>(v[b], s[b] / (float)C);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int b = 0; b < NB; b++) ss[b] += __shfl_xor_sync(0xffffffffu, ss[b], o);
    }
#pragma unroll
    for (int b = 0; b < NB; b++) pk_scale<VEC>(v[b], rsqrtf(ss[b] / (float)C + eps));   // MUFU.RSQ, <= 2 ulp
}

template <int VEC, int NB, typename TA, int STRIDE>
__global__ void __launch_bounds__(32 * ROWS_PER_CTA, (VEC <= 8 ? 2 : 1))
preattn_kernel(decaf_preattn_t p, int T_out, int strip_len, int strips_per_seq) {
    extern __shared__ __align__(16) float pre_smem[];      // [NB][5][C]: taps 0..2, w_br, b_br
    constexpr int C = 32 * VEC;                            // == p.C (dispatch): compile-time so that the shared-memory offsets are immediates
    for (int i = threadIdx.x; i < NB * 5 * C; i += blockDim.x) {
        const int c = i % C, k = (i / C) % 5, b = i / (5 * C);
        float v;
        if (k < 3) v = p.wd[((int64_t)b * C + c) * 3 + k];
        else if (k == 3) v = p.w_br[(int64_t)b * C + c];
        else v = p.b_br[(int64_t)b * C + c];
        pre_smem[i] = v;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t strip = (int64_t)blockIdx.x * ROWS_PER_CTA + warp;
    if (strip >= (int64_t)p.n_seq * strips_per_seq) return;
    const int seq = (int)(strip / strips_per_seq);
    const int t_begin = (int)(strip % strips_per_seq) * strip_len;
    const int t_end = min(t_begin + strip_len, T_out);
    const uint8_t *mrow = p.mask_in + (int64_t)seq * p.mi_seq_stride;
    const float *xs = p.x + (int64_t)seq * p.T_in * C;
    constexpr int stride = STRIDE;                         // == p.stride (dispatch)

    // validity of the strip's input rows tin0 .. tin0 + n_in - 1 (<= 2 * PRE_STRIP + 2 <= 32), one bit per row
    const int tin0 = stride * t_begin - 1;
    const int n_in = stride * (t_end - t_begin) + 2;
    unsigned okbits;
    {
        const int r = tin0 + lane;
        okbits = __ballot_sync(0xffffffffu, lane < n_in && r >= 0 && r < p.T_in && mrow[r] != 0);
    }
    float wp[VEC], bp[VEC];
    iload_row<VEC>(p.w_pre, lane, wp);
    iload_row<VEC>(p.b_pre, lane, bp);

    // window: w0 / w1 / w2 = LayerNorm-ed input rows (zero when masked / outside), r* = the raw rows for the
    // stride-2 max-pool skip (-inf when masked / outside).  Input rows are fetched PRE_DEPTH rows ahead with
    // cp.async into a per-warp shared-memory ring: ~1 KB x PRE_DEPTH x 16 warps per SM in flight is what it takes
    // to cover the HBM latency at full bandwidth (registers could not hold that many rows); every lane reads
    // back only the bytes it copied itself, so cp.async.wait_group is the only synchronisation needed.
    float w0[VEC], w1[VEC], w2[VEC];
    float r0[VEC], r1[VEC], r2[VEC];
    const bool want_skip = STRIDE == 2 && p.skip_out != nullptr;
    // (lane <-> channel mapping of iload_row: 16-byte pieces of a warp are contiguous when VEC % 4 == 0)
    float *ring = pre_smem + NB * 5 * C + (warp * PRE_DEPTH) * C;
    constexpr int LOFF = VEC % 4 == 0 ? 4 : VEC;         // floats between the first elements of neighbouring lanes
    auto prefetch = [&](int idx) {                       // idx = row index relative to tin0
        if (idx < n_in && ((okbits >> idx) & 1u)) {
            const float *src = xs + (int64_t)(tin0 + idx) * C + lane * LOFF;
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(ring + (idx % PRE_DEPTH) * C + lane * LOFF);
            constexpr int CH = VEC % 4 == 0 ? 16 : (VEC % 2 == 0 ? 8 : 4);
#pragma unroll
            for (int b = 0; b < VEC * 4; b += CH) {
                if constexpr (CH == 16)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + b * 32), "l"(reinterpret_cast<const char *>(src) + b * 32) : "memory");
                else if constexpr (CH == 8)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + b), "l"(reinterpret_cast<const char *>(src) + b) : "memory");
                else
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + b), "l"(reinterpret_cast<const char *>(src) + b) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");      // one group per row index, empty or not
    };
    auto consume = [&](int idx, float (&w)[VEC], float (&r)[VEC]) {     // waits for row idx, then prefetches idx + PRE_DEPTH
        const bool ok = (okbits >> idx) & 1u;
        asm volatile("cp.async.wait_group %0;" ::"n"(PRE_DEPTH - 1) : "memory");
        float x[VEC];
        if (ok) iload_row<VEC>(ring + (idx % PRE_DEPTH) * C, lane, x);
        prefetch(idx + PRE_DEPTH);
        if (ok) {
            if (want_skip) {
#pragma unroll
                for (int i = 0; i < VEC; i++) r[i] = x[i];
            }
            warp_layernorm<VEC>(x, C, p.eps);
            pk_affine<VEC>(x, wp, bp);
#pragma unroll
            for (int i = 0; i < VEC; i++) w[i] = x[i];
        } else {
#pragma unroll
            for (int i = 0; i < VEC; i++) { w[i] = 0.f; r[i] = -INFINITY; }
        }
    };
#pragma unroll
    for (int i = 0; i < PRE_DEPTH; i++) prefetch(i);
    consume(0, w1, r1);
    consume(1, w2, r2);
    int idx = 1;
    for (int t = t_begin; t < t_end; t++) {
        // advance the window so that (w0, w1, w2) = rows (stride*t - 1, stride*t, stride*t + 1)
        const int adv = (t == t_begin) ? 1 : stride;
        for (int a = 0; a < adv; a++) {
#pragma unroll
            for (int i = 0; i < VEC; i++) { w0[i] = w1[i]; w1[i] = w2[i]; r0[i] = r1[i]; r1[i] = r2[i]; }
            idx++;
            consume(idx, w2, r2);
        }
        const int64_t row = (int64_t)seq * T_out + t;
        const bool m_out = (okbits >> (idx - 1)) & 1u;      // row stride*t is always inside the sequence
        if (want_skip) {
            float sk[VEC];
#pragma unroll
            for (int i = 0; i < VEC; i++) {
                const float mx = fmaxf(fmaxf(r0[i], r1[i]), r2[i]);
                sk[i] = (m_out && mx > -INFINITY) ? mx : 0.f;
            }
            istore_row<VEC>(p.skip_out + row * C, lane, sk);
        }
        if (p.mask_out && lane == 0) p.mask_out[(int64_t)seq * p.mo_seq_stride + t] = m_out ? 1 : 0;
        float y[NB][VEC];
#pragma unroll
        for (int b = 0; b < NB; b++) {
            float k0[VEC], k1[VEC], k2[VEC];
            iload_row<VEC>(pre_smem + (b * 5 + 0) * C, lane, k0);
            iload_row<VEC>(pre_smem + (b * 5 + 1) * C, lane, k1);
            iload_row<VEC>(pre_smem + (b * 5 + 2) * C, lane, k2);
            pk_mul<VEC>(y[b], k0, w0);
            pk_fma<VEC>(y[b], k1, w1);
            pk_fma<VEC>(y[b], k2, w2);
        }
        warp_layernorm_multi<NB, VEC>(y, C, p.eps);
#pragma unroll
        for (int b = 0; b < NB; b++) {
            float gw[VEC], gb[VEC];
            iload_row<VEC>(pre_smem + (b * 5 + 3) * C, lane, gw);
            iload_row<VEC>(pre_smem + (b * 5 + 4) * C, lane, gb);
            pk_affine<VEC>(y[b], gw, gb);
            istore_row<VEC>(reinterpret_cast<TA *>(p.out_act) + (int64_t)b * p.out_branch_stride + row * C, lane, y[b]);
        }
    }
}

// ------------------------------------------------------------------------------- adaln
template <int VEC, typename TA>
__global__ void __launch_bounds__(32 * ROWS_PER_CTA) adaln_kernel(decaf_adaln_t p) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * ROWS_PER_CTA + warp;
    if (row >= p.rows) return;
    constexpr int C = 32 * VEC;                            // == p.C (dispatch)
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
    // this lane's slice of the conv weights stays in registers for all the rows the warp walks (the first version
    // re-read the 3 x C weights from L1 for every row and was L1-bound at 85 % of its throughput)
    float wreg[NOUT][3][VEC];
#pragma unroll
    for (int o = 0; o < NOUT; o++)
#pragma unroll
        for (int tap = 0; tap < 3; tap++) load_row<VEC>(w + ((int64_t)o * 3 + tap) * C, lane, wreg[o][tap]);
    for (int64_t row = (int64_t)blockIdx.x * ROWS_PER_CTA + warp; row < rows_total; row += (int64_t)gridDim.x * ROWS_PER_CTA) {
    const int level = level_of_row(lv, (int)(row % lv.Pp));
    if (level < 0) {
        if (lane < NOUT) out[row * NOUT + lane] = 0.f;
        continue;
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
#pragma unroll
            for (int i = 0; i < VEC; i++) acc[o] = fmaf(xv[i], wreg[o][tap][i], acc[o]);
        }
    }
#pragma unroll
    for (int o = 0; o < NOUT; o++) {
        float v = warp_sum(acc[o]) + bias[o];
        if (mode == 1) v = fmaxf(level_scale[level] * v, 0.f);
        if (lane == 0) out[row * NOUT + o] = v;
    }
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
                                 const float *__restrict__ bkgd, const float *__restrict__ pe, int pe_rows,
                                 const int32_t *__restrict__ len) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n_query * L1 * C) return;
    const int c = (int)(i % C);
    const int r = (int)((i / C) % L1);
    const int q = (int)(i / ((int64_t)C * L1));
    if (r == 0) {
        x[i] = bkgd[c];
    } else if (pe && (r - 1) < len[q]) {
        x[i] += text_pe_value(pe, pe_rows, C, len[q], r - 1, c);
    }
}

// start of the composed text encoder: x = 0, kv_len = len + 1 (background token), tmask[q, r] = r < kv_len[q]
__global__ void text_init_kernel(float *__restrict__ x, int n_query, int L1, int C, const int32_t *__restrict__ len,
                                 int32_t *__restrict__ kv_len, uint8_t *__restrict__ tmask) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (int64_t)n_query * L1 * C) x[i] = 0.f;
    if (i < (int64_t)n_query * L1) {
        const int q = (int)(i / L1), r = (int)(i % L1);
        tmask[i] = r < len[q] + 1 ? 1 : 0;
        if (r == 0) kv_len[q] = len[q] + 1;
    }
}

int head_out_mma_launch(const void *x, int64_t ldx, int rows_total, int C, const float *w, const float *bias, int n_out,
                        int mode, const float *level_scale, const decaf_levels_t *lv, float *out, cudaStream_t st);
bool head_out_mma_ok(const void *x, int64_t ldx, int C);

__global__ void cast_bf16_kernel(const float *__restrict__ in, bf16 *__restrict__ out, int64_t n4) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 v = reinterpret_cast<const float4 *>(in)[i];
    uint2 t;
    __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&t);
    h[0] = __floats2bfloat162_rn(v.x, v.y);
    h[1] = __floats2bfloat162_rn(v.z, v.w);
    reinterpret_cast<uint2 *>(out)[i] = t;
}

// fp32 -> three bf16 K-blocks per row (the FP32 configuration's tensor-core GEMMs): x = hi + lo + O(2^-17 |x|) with
// hi = bf16(x), lo = bf16(x - hi).  order 0 writes [hi | hi | lo], order 1 [hi | lo | hi]: the K-concatenated product
// of an order-0 activation row with an order-1 weight row is hi.hi + hi.lo + lo.hi, i.e. the fp32 product up to the
// lo.lo term (2^-16 relative), accumulated in fp32 by the tensor core.
__global__ void __launch_bounds__(256)
split_bf16x3_kernel(const float *__restrict__ src, int64_t rows, int K4, int64_t ld_src, bf16 *__restrict__ dst, int order) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * K4) return;
    const int64_t r = i / K4;
    const int c = (int)(i - r * K4) * 4;
    const float4 v = *reinterpret_cast<const float4 *>(src + r * ld_src + c);
    uint2 hi, lo;
    __nv_bfloat162 *hh = reinterpret_cast<__nv_bfloat162 *>(&hi), *ll = reinterpret_cast<__nv_bfloat162 *>(&lo);
    hh[0] = __floats2bfloat162_rn(v.x, v.y);
    hh[1] = __floats2bfloat162_rn(v.z, v.w);
    const float2 h0 = __bfloat1622float2(hh[0]), h1 = __bfloat1622float2(hh[1]);
    ll[0] = __floats2bfloat162_rn(v.x - h0.x, v.y - h0.y);
    ll[1] = __floats2bfloat162_rn(v.z - h1.x, v.w - h1.y);
    const int K = 4 * K4;
    bf16 *d = dst + r * 3 * (int64_t)K + c;
    *reinterpret_cast<uint2 *>(d) = hi;
    *reinterpret_cast<uint2 *>(d + K) = order == 0 ? hi : lo;
    *reinterpret_cast<uint2 *>(d + 2 * K) = order == 0 ? lo : hi;
}

}  // namespace decaf

using namespace decaf;

extern "C" int decaf_split_bf16x3(const float *src, int64_t rows, int32_t K, int64_t ld_src, void *dst, int32_t order, void *stream) {
    DECAF_CHECK(src && dst, "decaf_split_bf16x3: null pointers");
    DECAF_CHECK(K > 0 && K % 4 == 0 && ld_src % 4 == 0 && ld_src >= K, "decaf_split_bf16x3: K and ld_src must be multiples of 4, ld_src >= K");
    DECAF_CHECK(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0, "decaf_split_bf16x3: unaligned buffers");
    DECAF_CHECK(order == 0 || order == 1, "decaf_split_bf16x3: order must be 0 or 1");
    const int64_t n = rows * (K / 4);
    if (n <= 0) return 0;
    split_bf16x3_kernel<<<(unsigned)cdiv(n, 256), 256, 0, as_stream(stream)>>>(src, rows, K / 4, ld_src, reinterpret_cast<bf16 *>(dst), order);
    DECAF_LAUNCH_CHECK();
    return 0;
}

extern "C" int decaf_cast_bf16(const float *in, void *out, int64_t n, void *stream) {
    DECAF_CHECK(in && out, "decaf_cast_bf16: null pointers");
    DECAF_CHECK(n % 4 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0,
                "decaf_cast_bf16: n %% 4 != 0 or unaligned buffers");
    if (n == 0) return 0;
    cast_bf16_kernel<<<cdiv(n / 4, 256), 256, 0, as_stream(stream)>>>(in, reinterpret_cast<bf16 *>(out), n / 4);
    DECAF_LAUNCH_CHECK();
    return 0;
}

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
    // strip length: long strips amortise the 2-row halo, short ones keep enough warps in flight on the small levels
    int strip_len = (int)(rows / 4096);
    strip_len = strip_len < 1 ? 1 : (strip_len > PRE_STRIP ? PRE_STRIP : strip_len);
    const int strips_per_seq = cdiv(T_out, strip_len);
    const int grid = cdiv((int64_t)p.n_seq * strips_per_seq, ROWS_PER_CTA);
    const size_t smem = ((size_t)p.n_branch * 5 + ROWS_PER_CTA * PRE_DEPTH) * p.C * sizeof(float);
    DECAF_CHECK(smem <= 200 * 1024, "decaf_preattn: C too large for the staged weights and row ring (%d)", p.C);
    DECAF_CHECK(p.n_branch == 1 || p.n_branch == 3, "decaf_preattn: n_branch must be 1 or 3");
    cudaStream_t st = as_stream(stream);
#define PRE_LAUNCH(NB, TA)                                                                                              \
    DECAF_DISPATCH_VEC(p.C, {                                                                                           \
        static bool attr_set = false;                                                                                   \
        if (!attr_set && smem > 48 * 1024) {                                                                            \
            DECAF_CUDA(cudaFuncSetAttribute(preattn_kernel<VEC, NB, TA, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                            200 * 1024));                                                               \
            DECAF_CUDA(cudaFuncSetAttribute(preattn_kernel<VEC, NB, TA, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                            200 * 1024));                                                               \
            attr_set = true;                                                                                            \
        }                                                                                                               \
        if (p.stride == 2)                                                                                              \
            preattn_kernel<VEC, NB, TA, 2><<<grid, 32 * ROWS_PER_CTA, smem, st>>>(p, T_out, strip_len, strips_per_seq); \
        else                                                                                                            \
            preattn_kernel<VEC, NB, TA, 1><<<grid, 32 * ROWS_PER_CTA, smem, st>>>(p, T_out, strip_len, strips_per_seq); \
    })
    if (p.dtype == DECAF_BF16) {
        if (p.n_branch == 3) { PRE_LAUNCH(3, bf16); } else { PRE_LAUNCH(1, bf16); }
    } else {
        if (p.n_branch == 3) { PRE_LAUNCH(3, float); } else { PRE_LAUNCH(1, float); }
    }
#undef PRE_LAUNCH
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
    cudaStream_t st = as_stream(stream);
    if (dtype == DECAF_BF16 && head_out_mma_ok(x, ldx, C))   // tensor-core path (head_out_mma.cu)
        return head_out_mma_launch(x, ldx, rows_total, C, w, bias, n_out, mode, level_scale, lv, out, st);
    int grid = cdiv(rows_total, ROWS_PER_CTA);
    if (grid > 148 * 16) grid = 148 * 16;                    // grid-stride: each warp keeps its weight slice in registers
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

extern "C" int decaf_text_init(float *x, int32_t n_query, int32_t L1, int32_t C, const int32_t *len, int32_t *kv_len,
                               uint8_t *tmask, void *stream) {
    DECAF_CHECK(x && len && kv_len && tmask, "decaf_text_init: null pointers");
    const int64_t n = (int64_t)n_query * L1 * C;
    if (n == 0) return 0;
    text_init_kernel<<<cdiv(n, 256), 256, 0, as_stream(stream)>>>(x, n_query, L1, C, len, kv_len, tmask);
    DECAF_LAUNCH_CHECK();
    return 0;
}

extern "C" int decaf_text_prep(float *x, int32_t n_query, int32_t L1, int32_t C, const float *bkgd,
                               const float *pe, int32_t pe_rows, const int32_t *len, void *stream) {
    DECAF_CHECK(x && bkgd && len, "decaf_text_prep: null pointers");
    DECAF_CHECK(!pe || pe_rows >= 2, "decaf_text_prep: pe_rows must be >= 2");
    const int64_t n = (int64_t)n_query * L1 * C;
    if (n == 0) return 0;
    text_prep_kernel<<<cdiv(n, 256), 256, 0, as_stream(stream)>>>(x, n_query, L1, C, bkgd, pe, pe_rows, len);
    DECAF_LAUNCH_CHECK();
    return 0;
}

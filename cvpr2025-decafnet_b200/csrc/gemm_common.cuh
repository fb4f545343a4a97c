// Epilogue shared by the SIMT and tcgen05 GEMM kernels.
#pragma once
#include "common.cuh"

namespace decaf {

// Device-side copy of decaf_gemm_t for ONE group (pointers already offset by the group).
struct GemmArgs {
    const void *A; int64_t lda, a_seq_stride;
    int n_seq, rows_per_seq;
    const void *W;
    int N, K, taps, dil;
    const float *bias;
    int act;
    const float *colscale;
    const float *resid; int64_t ldr, r_seq_stride;
    const uint8_t *rowmask; int64_t m_seq_stride;
    float *out_f32; int64_t ldo, o_seq_stride;
    void *out_act; int64_t ldo2, o2_seq_stride;
    int64_t g_stride_a, g_stride_w, g_stride_bias, g_stride_out_f32, g_stride_out_act;
    int ln; const float *ln_w, *ln_b; float ln_eps;
    const float *pe;
};

// v = acc + bias; act; * colscale; + resid; + pe; * rowmask  (see include/decaf_b200.h)
__device__ __forceinline__ float gemm_epilogue_value(const GemmArgs &p, float acc, int seq, int t, int n,
                                                     float rowmask, const float *bias, int group) {
    float v = acc;
    if (bias) v += bias[n];
    if (p.act == DECAF_ACT_RELU) v = fmaxf(v, 0.f);
    else if (p.act == DECAF_ACT_GELU) v = gelu_erf(v);
    if (p.colscale) v *= p.colscale[n];
    if (p.resid) v += p.resid[((int64_t)seq * p.r_seq_stride + t) * p.ldr + n];
    if (p.pe) v += p.pe[(int64_t)t * p.N + n];
    return v * rowmask;
}

}  // namespace decaf

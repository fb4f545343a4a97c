// SIMT fp32-FMA GEMM / conv1d (taps) with the shared epilogue.  This is the arithmetic of the
// FP32 configuration (the reference disables TF32, eval.py:40-41) and the numerical reference
// the tcgen05 kernel is unit-tested against.  A and W may be fp32 or bf16; accumulation is fp32.
#include "gemm_common.cuh"

namespace decaf {

constexpr int SBM = 64, SBN = 64, SBK = 16;

template <typename TA>
__global__ void __launch_bounds__(256) gemm_simt_kernel(GemmArgs p, int tiles_per_seq) {
    __shared__ float As[SBK][SBM + 4];
    __shared__ float Bs[SBK][SBN + 4];

    const int g = blockIdx.z;
    const TA *A = reinterpret_cast<const TA *>(p.A) + (int64_t)g * p.g_stride_a;
    const TA *W = reinterpret_cast<const TA *>(p.W) + (int64_t)g * p.g_stride_w;
    const float *bias = p.bias ? p.bias + (int64_t)g * p.g_stride_bias : nullptr;

    const int seq = blockIdx.x / tiles_per_seq;
    const int m0 = (blockIdx.x % tiles_per_seq) * SBM;
    const int n0 = blockIdx.y * SBN;
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

    const int lr = tid / 4;          // 0..63: tile row (A) / tile col (W)
    const int lk = (tid % 4) * 4;    // 0,4,8,12
    const TA *Aseq = A + (int64_t)seq * p.a_seq_stride * p.lda;

    for (int tap = 0; tap < p.taps; tap++) {
        const int shift = (tap - p.taps / 2) * p.dil;
        const int t_src = m0 + lr + shift;
        const bool row_ok = (m0 + lr) < p.rows_per_seq && t_src >= 0 && t_src < p.rows_per_seq;
        const TA *arow = Aseq + (int64_t)t_src * p.lda;
        const int n_src = n0 + lr;
        const bool col_ok = n_src < p.N;
        const TA *wrow = W + ((int64_t)n_src * p.taps + tap) * p.K;
        // software pipeline: the next k-block is fetched into registers while the current one is multiplied (the text
        // encoder's GEMMs are 14 CTAs with K <= 768: un-pipelined, every k-block exposed a full L2 round trip)
        float ra[4], rb[4];
        auto fetch = [&](int k0) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int k = k0 + lk + i;
                ra[i] = (row_ok && k < p.K) ? to_f32<TA>(arow[k]) : 0.f;
                rb[i] = (col_ok && k < p.K) ? to_f32<TA>(wrow[k]) : 0.f;
            }
        };
        fetch(0);
        for (int k0 = 0; k0 < p.K; k0 += SBK) {
#pragma unroll
            for (int i = 0; i < 4; i++) { As[lk + i][lr] = ra[i]; Bs[lk + i][lr] = rb[i]; }
            __syncthreads();
            if (k0 + SBK < p.K) fetch(k0 + SBK);
#pragma unroll
            for (int k = 0; k < SBK; k++) {
                const float4 a = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
                const float4 b = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
                const float av[4] = {a.x, a.y, a.z, a.w};
                const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            __syncthreads();
        }
    }

    float *of = p.out_f32 ? p.out_f32 + (int64_t)g * p.g_stride_out_f32 : nullptr;
    TA *oa = p.out_act ? reinterpret_cast<TA *>(p.out_act) + (int64_t)g * p.g_stride_out_act : nullptr;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int t = m0 + ty * 4 + i;
        if (t >= p.rows_per_seq) continue;
        const float rm = p.rowmask ? (float)p.rowmask[(int64_t)seq * p.m_seq_stride + t] : 1.f;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int n = n0 + tx * 4 + j;
            if (n >= p.N) continue;
            const float v = gemm_epilogue_value(p, acc[i][j], seq, t, n, rm, bias, g);
            if (of) of[((int64_t)seq * p.o_seq_stride + t) * p.ldo + n] = v;
            if (oa) oa[((int64_t)seq * p.o2_seq_stride + t) * p.ldo2 + n] = from_f32<TA>(v);
        }
    }
}

int gemm_simt_launch(const GemmArgs &a, int dtype, int n_group, cudaStream_t st) {
    const int tiles_per_seq = cdiv(a.rows_per_seq, SBM);
    dim3 grid(a.n_seq * tiles_per_seq, cdiv(a.N, SBN), n_group);
    if (dtype == DECAF_F32)
        gemm_simt_kernel<float><<<grid, 256, 0, st>>>(a, tiles_per_seq);
    else
        gemm_simt_kernel<bf16><<<grid, 256, 0, st>>>(a, tiles_per_seq);
    DECAF_LAUNCH_CHECK();
    return 0;
}

}  // namespace decaf

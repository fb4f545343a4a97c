// Query-aware saliency scorer, exact top-k clip selection and the merge of expert/sidekick
// features into the dense per-query timeline (reference: libs/modeling/model.py:500-554).
// All three are HBM-bound: the (C, T) feature planes are streamed once with T-contiguous
// (coalesced) reads; the merge transposes 32x32 tiles through shared memory so both the
// (C, T) reads and the channels-last (T, C) writes are full 128-byte lines.
#include "common.cuh"

namespace decaf {

constexpr int SAL_WARPS = 8;    // channel slices
constexpr int SAL_QCH = 8;      // queries per CTA pass

// One CTA = 32 * TPT consecutive time steps x SAL_QCH queries; warp w walks channels w, w + 8, ...: a lane loads TPT
// consecutive steps of the channel row (128- / 512-byte coalesced warp loads, several rows in flight) and multiplies
// them with the SAL_QCH normalised text weights of that channel, read as two broadcast LDS.128 from the [channel][query]
// table - with TPT = 4 that is 36 FMAs per 3 memory instructions, so long timelines (MAD: 64 queries x 71k steps,
// 2.9 GFLOP of fp32 FMA over 91 MB) run at the FMA rate instead of the shared-memory rate (8 LDS per 9 FMA before).
// TPT = 1 keeps short videos spread over enough CTAs.
template <int TPT>
__global__ void __launch_bounds__(32 * SAL_WARPS)
saliency_kernel(const float *__restrict__ shallow, const float *__restrict__ text_cls,
                float *__restrict__ correl, int Cs, int T, int n_query, int norm, int tiles) {
    constexpr int TX = 32 * TPT;
    extern __shared__ __align__(16) float smem[];
    float *tn = smem;                                    // [Cs][SAL_QCH] normalised text vectors
    float *red = smem + SAL_QCH * Cs;                    // [SAL_WARPS][SAL_QCH + 1][TX]
    __shared__ float tscale[SAL_QCH];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int q0 = blockIdx.y * SAL_QCH;
    const int nq = min(SAL_QCH, n_query - q0);

    // text norms: warp ty handles query q0 + ty
    if (ty < nq) {
        float ss = 0.f;
        for (int h = tx; h < Cs; h += 32) {
            const float x = text_cls[(int64_t)(q0 + ty) * Cs + h];
            ss += x * x;
        }
        ss = warp_sum(ss);
        if (tx == 0) tscale[ty] = norm ? 1.0f / (sqrtf(ss) + 1e-4f) : 1.0f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SAL_QCH * Cs; i += blockDim.x) {
        const int qq = i / Cs, h = i % Cs;
        tn[h * SAL_QCH + qq] = qq < nq ? text_cls[(int64_t)(q0 + qq) * Cs + h] * tscale[qq] : 0.f;
    }
    __syncthreads();

    // `tiles` consecutive TX-step tiles per CTA: the text-vector prologue above (two dependent global round trips) is paid
    // once for all of them
    for (int tile = 0; tile < tiles; tile++) {
        const int tbase = ((int)blockIdx.x * tiles + tile) * TX;
        if (tbase >= T) break;
        const int t0 = tbase + tx * TPT;
        float ss[TPT], dot[SAL_QCH][TPT];
#pragma unroll
        for (int j = 0; j < TPT; j++) {
            ss[j] = 0.f;
#pragma unroll
            for (int i = 0; i < SAL_QCH; i++) dot[i][j] = 0.f;
        }
        const bool vec_ok = TPT == 4 && (T % 4 == 0) && t0 + 3 < T && (reinterpret_cast<uintptr_t>(shallow) & 15) == 0;
#pragma unroll 4
        for (int h = ty; h < Cs; h += SAL_WARPS) {
            float x[TPT];
            const float *row = shallow + (int64_t)h * T;
            if constexpr (TPT == 4) {
                if (vec_ok) {
                    const float4 v4 = *reinterpret_cast<const float4 *>(row + t0);
                    x[0] = v4.x; x[1] = v4.y; x[2] = v4.z; x[3] = v4.w;
                } else {
#pragma unroll
                    for (int j = 0; j < TPT; j++) x[j] = t0 + j < T ? row[t0 + j] : 0.f;
                }
            } else {
                x[0] = t0 < T ? row[t0] : 0.f;
            }
            const float4 wa = *reinterpret_cast<const float4 *>(tn + h * SAL_QCH), wb = *reinterpret_cast<const float4 *>(tn + h * SAL_QCH + 4);
            const float w[SAL_QCH] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
            for (int j = 0; j < TPT; j++) {
                ss[j] = fmaf(x[j], x[j], ss[j]);
#pragma unroll
                for (int i = 0; i < SAL_QCH; i++) dot[i][j] = fmaf(x[j], w[i], dot[i][j]);
            }
        }
        if (tile > 0) __syncthreads();                     // the previous tile's finalisation has read `red`
        float *r = red + (ty * (SAL_QCH + 1)) * TX;
#pragma unroll
        for (int j = 0; j < TPT; j++) {
            r[tx * TPT + j] = ss[j];
#pragma unroll
            for (int i = 0; i < SAL_QCH; i++) r[(i + 1) * TX + tx * TPT + j] = dot[i][j];
        }
        __syncthreads();
        // warp ty finalises query q0 + ty for the tile's time steps
        if (ty < nq) {
#pragma unroll
            for (int j = 0; j < TPT; j++) {
                const int tt = tx + 32 * j, t = tbase + tt;
                if (t >= T) continue;
                float s2 = 0.f, d = 0.f;
#pragma unroll
                for (int w = 0; w < SAL_WARPS; w++) {
                    s2 += red[(w * (SAL_QCH + 1)) * TX + tt];
                    d += red[(w * (SAL_QCH + 1) + ty + 1) * TX + tt];
                }
                const float inv = norm ? 1.0f / (sqrtf(s2) + 1e-4f) : 1.0f;
                correl[(int64_t)(q0 + ty) * T + t] = d * inv;
            }
        }
    }
}

// one CTA per query.  Block means: the scores of up to `chunk` blocks are staged in shared memory with coalesced loads
// (rows padded to sn + 1 floats: conflict-free), then one thread per block adds its sn values IN SEQUENCE ORDER - the sum of
// avg_pool1d - out of shared memory.  (A thread summing straight from global memory waits on sn dependent loads: 14 us of
// pure latency at the NLQ shape; one warp per block with shuffles is fast there but leaves only 8 chains for the 1191 blocks
// of a MAD timeline: 0.5 ms.)  Ranks: exact stable ascending rank by counting.  Expansion: 4 steps per thread.
__global__ void __launch_bounds__(256)
select_kernel(const float *__restrict__ correl, const uint8_t *__restrict__ vid_mask,
              uint8_t *__restrict__ sel, uint8_t *__restrict__ out_mask, float *__restrict__ pooled_out,
              int max_blocks, int T, int sn, double sratio, int and_mask, int32_t *__restrict__ vid_len_out, int chunk) {
    // [max_blocks, rounded up to 4, + 4 floats of slack: the rank loop reads pooled[] four at a time] | selected flags
    // [max_blocks] (padded to 16 B) | staging
    extern __shared__ float pooled[];
    const int pooled_len = ((max_blocks + 3) & ~3) + 4;
    uint8_t *selected = reinterpret_cast<uint8_t *>(pooled + pooled_len);
    float *stage = pooled + pooled_len + (max_blocks + 15) / 16 * 4;    // [chunk][sn + 1]
    __shared__ int s_len;
    __shared__ int warp_cnt[8];
    const int q = blockIdx.x, tid = threadIdx.x;
    const float *c = correl + (int64_t)q * T;

    int cnt = 0;
    if ((reinterpret_cast<uintptr_t>(vid_mask) & 15) == 0) {
        const int T16 = T / 16;
        for (int i = tid; i < T16; i += blockDim.x) {
            const uint4 m = reinterpret_cast<const uint4 *>(vid_mask)[i];
            const uint32_t w[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
            for (int k = 0; k < 4; k++)
#pragma unroll
                for (int b = 0; b < 4; b++) cnt += ((w[k] >> (8 * b)) & 0xffu) != 0;
        }
        for (int t = T16 * 16 + tid; t < T; t += blockDim.x) cnt += vid_mask[t] != 0;
    } else {
        for (int t = tid; t < T; t += blockDim.x) cnt += vid_mask[t] != 0;
    }
    cnt = warp_sum_i(cnt);
    if ((tid & 31) == 0) warp_cnt[tid >> 5] = cnt;
    __syncthreads();
    if (tid == 0) {
        int s = 0;
        for (int w = 0; w < 8; w++) s += warp_cnt[w];
        s_len = s;
        if (vid_len_out && q == 0) *vid_len_out = s;
    }
    __syncthreads();
    const int len = s_len;
    const int M = (len + sn - 1) / sn;                      // <= max_blocks (host guarantees)

    // block means: sequential fp32 sum over the valid part of each block (ceil_mode avg_pool1d)
    const int pitch = sn + 1;
    for (int j0 = 0; j0 < M; j0 += chunk) {
        const int nb = min(chunk, M - j0);
        const int a0 = j0 * sn, n_el = min(nb * sn, len - a0);
#pragma unroll 4
        for (int i = tid; i < n_el; i += blockDim.x) {
            const int jb = i / sn;
            stage[jb * pitch + (i - jb * sn)] = c[a0 + i];
        }
        __syncthreads();
        if (tid < nb) {
            const int j = j0 + tid;
            const int a = j * sn, b = min(a + sn, len);
            const float *sp = stage + tid * pitch;
            float acc = 0.f;
            for (int t = 0; t < b - a; t++) acc += sp[t];
            const float mean = acc / (float)(b - a);
            pooled[j] = mean;
            if (pooled_out) pooled_out[(int64_t)q * max_blocks + j] = mean;
        }
        __syncthreads();
    }
    const int k = (int)(sratio * (double)M);                // Python: int(ratio * M)
    for (int j = tid; j < M; j += blockDim.x) {
        const float pj = pooled[j];
        int rank = 0;                                       // stable ascending rank
#pragma unroll 4
        for (int i = 0; i < M; i++) {
            const float pi = pooled[i];
            rank += (pi < pj) || (pi == pj && i < j);
        }
        selected[j] = (k == 0) || (rank >= M - k);          // ranked[-0:] selects everything
    }
    __syncthreads();
    const float scale = len > 0 ? (float)M / (float)len : 0.f;
    auto one = [&](int t, uint8_t vmb) -> uchar2 {
        uint8_t s = 0;
        if (t < len) {
            int src = (int)floorf((float)t * scale);        // F.interpolate(mode='nearest') index
            src = min(src, M - 1);
            s = selected[src];
        }
        const uint8_t vm = vmb != 0;
        return make_uchar2(s, and_mask ? (uint8_t)(vm & s) : vm);
    };
    uint8_t *sq = sel + (int64_t)q * T, *oq = out_mask + (int64_t)q * T;
    if (T % 4 == 0 && ((reinterpret_cast<uintptr_t>(vid_mask) | reinterpret_cast<uintptr_t>(sq) | reinterpret_cast<uintptr_t>(oq)) & 3) == 0) {
        for (int i = tid; i < T / 4; i += blockDim.x) {
            const uchar4 vm4 = reinterpret_cast<const uchar4 *>(vid_mask)[i];
            const uchar2 r0 = one(4 * i, vm4.x), r1 = one(4 * i + 1, vm4.y), r2 = one(4 * i + 2, vm4.z), r3 = one(4 * i + 3, vm4.w);
            reinterpret_cast<uchar4 *>(sq)[i] = make_uchar4(r0.x, r1.x, r2.x, r3.x);
            reinterpret_cast<uchar4 *>(oq)[i] = make_uchar4(r0.y, r1.y, r2.y, r3.y);
        }
    } else {
        for (int t = tid; t < T; t += blockDim.x) {
            const uchar2 r = one(t, vid_mask[t]);
            sq[t] = r.x;
            oq[t] = r.y;
        }
    }
}

template <typename TA>
__global__ void __launch_bounds__(256)
merge_kernel(const float *__restrict__ vid, int Ce, const float *__restrict__ shallow, int Cs,
             const float *__restrict__ correl, int scat, const uint8_t *__restrict__ sel,
             const uint8_t *__restrict__ out_mask, TA *__restrict__ x0, int64_t ldx, int T) {
    __shared__ float tile[32][33];
    const int q = blockIdx.y;
    const int t0 = blockIdx.x * 32, c0 = blockIdx.z * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
    // read: tx over t (contiguous in the (C, T) planes)
    const int t = t0 + tx;
    float ms = 0.f, mm = 0.f;
    if (t < T) {
        ms = (float)sel[(int64_t)q * T + t];
        mm = (float)out_mask[(int64_t)q * T + t];
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int c = c0 + ty + i * 8;
        float v = 0.f;
        if (t < T) {
            if (c < Ce) v = vid[(int64_t)c * T + t] * ms;
            else if (c < Ce + Cs) v = shallow[(int64_t)(c - Ce) * T + t];
            else if (scat && c == Ce + Cs) v = correl[(int64_t)q * T + t];
            v *= mm;
        }
        tile[ty + i * 8][tx] = v;
    }
    __syncthreads();
    // write: tx over c (contiguous in channels-last)
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int tt = t0 + ty + i * 8;
        const int c = c0 + tx;
        if (tt < T && c < ldx) x0[((int64_t)q * T + tt) * ldx + c] = from_f32<TA>(tile[tx][ty + i * 8]);
    }
}

// vid_map by linearity.  The reference applies the 1x1 conv to cat[vid * sel_q, shallow (, correl_q)] once per query
// (libs/modeling/model.py:543-555); the conv is linear and sel_q is a 0/1 row mask, so
//   vid_map_q[t] = mask_q[t] * ( sel_q[t] * E[t] + S[t] + bias (+ correl_q[t] * w_c) ),  E = W_e vid,  S = W_s shallow
// with E and S computed ONCE per video (T rows instead of n_query * T rows of K = 2 Cin).  One warp per (query, step).
template <int VEC>
__global__ void __launch_bounds__(256)
map_combine_kernel(const float *__restrict__ E, const float *__restrict__ S, const float *__restrict__ bias,
                   const float *__restrict__ correl, const float *__restrict__ wc, const uint8_t *__restrict__ sel,
                   const uint8_t *__restrict__ mask, float *__restrict__ X, int T, int C, int n_query, int q_per_warp) {
    // One warp per (time step, group of q_per_warp queries): the E / S rows of the step are loaded once and combined
    // for every query of the group; lane j fetches the mask / selection / correlation scalars of query j and the loop
    // broadcasts them.  (One warp per (query, step) re-read both rows per query and waited on a dependent
    // mask -> row load chain: 16 us for 38 MB of output at the NLQ shape.)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t item = (int64_t)blockIdx.x * 8 + warp;
    const int groups = (n_query + q_per_warp - 1) / q_per_warp;
    if (item >= (int64_t)T * groups) return;
    const int t = (int)(item / groups), q0 = (int)(item % groups) * q_per_warp;
    const int nq = min(q_per_warp, n_query - q0);
    float base[VEC], e[VEC], w[VEC];
    load_row<VEC>(bias, lane, base);
    if (S) {
        float s[VEC];
        load_row<VEC>(S + (int64_t)t * C, lane, s);
#pragma unroll
        for (int i = 0; i < VEC; i++) base[i] += s[i];
    }
#pragma unroll
    for (int i = 0; i < VEC; i++) { e[i] = 0.f; w[i] = 0.f; }
    if (E) load_row<VEC>(E + (int64_t)t * C, lane, e);
    if (correl) load_row<VEC>(wc, lane, w);
    for (int qb = 0; qb < nq; qb += 32) {
        const int ql = q0 + qb + lane;
        const bool in = qb + lane < nq;
        const int64_t r = (int64_t)ql * T + t;
        const unsigned m_bits = __ballot_sync(0xffffffffu, in && mask[r] != 0);
        const unsigned s_bits = __ballot_sync(0xffffffffu, in && E != nullptr && sel[r] != 0);
        const float cq_l = (in && correl) ? correl[r] : 0.f;
        const int cnt = min(32, nq - qb);
        for (int j = 0; j < cnt; j++) {
            const float cq = __shfl_sync(0xffffffffu, cq_l, j);
            float v[VEC];
#pragma unroll
            for (int i = 0; i < VEC; i++) v[i] = 0.f;
            if ((m_bits >> j) & 1u) {
                // same order as the reference expression: bias + S, + E when selected, + correl * w_c
#pragma unroll
                for (int i = 0; i < VEC; i++) v[i] = base[i];
                if ((s_bits >> j) & 1u) {
#pragma unroll
                    for (int i = 0; i < VEC; i++) v[i] += e[i];
                }
                if (correl) {
#pragma unroll
                    for (int i = 0; i < VEC; i++) v[i] = fmaf(cq, w[i], v[i]);
                }
            }
            store_row<VEC>(X + ((int64_t)(q0 + qb + j) * T + t) * C, lane, v);
        }
    }
}

// Compact expert-feature ingest: only the clips some query selected need expert features.
// dense[c, index[k]] = compact[c, k]; everything else stays zero (the merge multiplies unselected steps by 0 anyway).
__global__ void scatter_clips_kernel(const float *__restrict__ compact, int64_t ld, const int32_t *__restrict__ index, int Ce, int K,
                                     float *__restrict__ dense, int T) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)Ce * K) return;
    const int c = (int)(i / K), k = (int)(i % K);
    const int t = index[k];
    if (t >= 0 && t < T) dense[(int64_t)c * T + t] = compact[(int64_t)c * ld + k];
}

}  // namespace decaf

using namespace decaf;

extern "C" int decaf_scatter_clips(const float *compact, int64_t ld, const int32_t *index, int32_t Ce, int32_t K, float *dense,
                                   int32_t T, void *stream) {
    DECAF_CHECK(dense && (K == 0 || (compact && index)), "decaf_scatter_clips: null pointers");
    DECAF_CHECK(ld >= K, "decaf_scatter_clips: ld < K");
    cudaStream_t st = as_stream(stream);
    DECAF_CUDA(cudaMemsetAsync(dense, 0, (size_t)Ce * T * sizeof(float), st));
    const int64_t n = (int64_t)Ce * K;
    if (n == 0) return 0;
    scatter_clips_kernel<<<cdiv(n, 256), 256, 0, st>>>(compact, ld, index, Ce, K, dense, T);
    DECAF_LAUNCH_CHECK();
    return 0;
}

extern "C" int decaf_map_combine(const float *E, const float *S, const float *bias, const float *correl, const float *wc,
                                 const uint8_t *sel, const uint8_t *mask, float *X, int32_t T, int32_t C, int32_t n_query,
                                 void *stream) {
    DECAF_CHECK((E || S) && bias && sel && mask && X, "decaf_map_combine: null pointers");
    DECAF_CHECK((correl == nullptr) == (wc == nullptr), "decaf_map_combine: correl and wc go together");
    DECAF_CHECK(C % 32 == 0, "decaf_map_combine: C %% 32 != 0");
    const int64_t rows = (int64_t)n_query * T;
    if (rows == 0) return 0;
    // queries per warp: as many as keep >= ~8 warps per SM-resident slot (16 at the NLQ shape, all 64 of a MAD video)
    int qpw = n_query;
    while (qpw > 4 && (int64_t)T * cdiv(n_query, qpw) < 148 * 64) qpw = (qpw + 1) / 2;
    const int64_t items = (int64_t)T * cdiv(n_query, qpw);
    const int grid = cdiv(items, 8);
    cudaStream_t st = as_stream(stream);
    DECAF_DISPATCH_VEC(C, (map_combine_kernel<VEC><<<grid, 256, 0, st>>>(E, S, bias, correl, wc, sel, mask, X, T, C, n_query, qpw)));
    DECAF_LAUNCH_CHECK();
    return 0;
}

extern "C" int decaf_saliency(const float *shallow, const float *text_cls, float *correl, int32_t Cs,
                              int32_t T, int32_t n_query, int32_t norm, void *stream) {
    DECAF_CHECK(shallow && text_cls && correl, "decaf_saliency: null pointers");
    if (T == 0 || n_query == 0) return 0;
    DECAF_CHECK(Cs <= 4096, "decaf_saliency: Cs too large (%d)", Cs);
    cudaStream_t st = as_stream(stream);
    const int qc = cdiv(n_query, SAL_QCH);
    // four steps per thread once that still leaves >= 2 CTAs per SM
    if ((int64_t)cdiv(T, 128) * qc >= 2 * 148) {
        const size_t smem = sizeof(float) * (SAL_QCH * Cs + SAL_WARPS * (SAL_QCH + 1) * 128);
        static bool attr = false;
        if (!attr) { DECAF_CUDA(cudaFuncSetAttribute(saliency_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr = true; }
        // one tile per CTA: 2 / 4 / 8 tiles per CTA (prologue amortised) measured 235 us against 188 us at the MAD shape
        const int tiles = 1;
        saliency_kernel<4><<<dim3(cdiv(T, 128 * tiles), qc), 32 * SAL_WARPS, smem, st>>>(shallow, text_cls, correl, Cs, T, n_query, norm, tiles);
    } else {
        const size_t smem = sizeof(float) * (SAL_QCH * Cs + SAL_WARPS * (SAL_QCH + 1) * 32);
        static bool attr = false;
        if (!attr) { DECAF_CUDA(cudaFuncSetAttribute(saliency_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr = true; }
        saliency_kernel<1><<<dim3(cdiv(T, 32), qc), 32 * SAL_WARPS, smem, st>>>(shallow, text_cls, correl, Cs, T, n_query, norm, 1);
    }
    DECAF_LAUNCH_CHECK();
    return 0;
}

extern "C" int decaf_select(const float *correl, const uint8_t *vid_mask, uint8_t *sel, uint8_t *out_mask,
                            float *pooled, int32_t max_blocks, int32_t T, int32_t n_query, int32_t sn,
                            double sratio, int32_t and_mask, int32_t *vid_len_out, void *stream) {
    DECAF_CHECK(correl && vid_mask && sel && out_mask, "decaf_select: null pointers");
    DECAF_CHECK(sn > 0, "decaf_select: sn must be > 0");
    DECAF_CHECK(max_blocks >= (T + sn - 1) / sn, "decaf_select: max_blocks %d < ceil(T/sn)", max_blocks);
    if (T == 0 || n_query == 0) return 0;
    // staging: up to 256 blocks of sn + 1 floats per pass, at most ~64 KB (a single block larger than that: one block per pass)
    int chunk = 16384 / (sn + 1);
    chunk = chunk < 1 ? 1 : (chunk > 256 ? 256 : chunk);
    if (chunk > max_blocks) chunk = max_blocks;
    const size_t smem = ((size_t)((max_blocks + 3) & ~3) + 4 + (size_t)(max_blocks + 15) / 16 * 4 + (size_t)chunk * (sn + 1)) * sizeof(float) + 16;
    DECAF_CHECK(smem <= 200 * 1024, "decaf_select: too many blocks / too long blocks (%d x %d)", max_blocks, sn);
    static size_t attr = 0;
    if (smem > 48 * 1024 && smem > attr) {
        DECAF_CUDA(cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr = smem;
    }
    select_kernel<<<n_query, 256, smem, as_stream(stream)>>>(correl, vid_mask, sel, out_mask, pooled, max_blocks, T,
                                                             sn, sratio, and_mask, vid_len_out, chunk);
    DECAF_LAUNCH_CHECK();
    return 0;
}

extern "C" int decaf_merge(const float *vid, int32_t Ce, const float *shallow, int32_t Cs, const float *correl,
                           int32_t scat, const uint8_t *sel, const uint8_t *out_mask, void *x0, int32_t dtype,
                           int64_t ldx, int32_t T, int32_t n_query, void *stream) {
    DECAF_CHECK(sel && out_mask && x0, "decaf_merge: null pointers");
    DECAF_CHECK((Ce == 0 || vid) && (Cs == 0 || shallow) && (!scat || correl), "decaf_merge: missing input plane");
    DECAF_CHECK(ldx >= Ce + Cs + (scat ? 1 : 0), "decaf_merge: ldx too small");
    if (T == 0 || n_query == 0) return 0;
    dim3 grid(cdiv(T, 32), n_query, cdiv(ldx, 32));
    cudaStream_t st = as_stream(stream);
    if (dtype == DECAF_BF16)
        merge_kernel<bf16><<<grid, 256, 0, st>>>(vid, Ce, shallow, Cs, correl, scat, sel, out_mask, (bf16 *)x0, ldx, T);
    else
        merge_kernel<float><<<grid, 256, 0, st>>>(vid, Ce, shallow, Cs, correl, scat, sel, out_mask, (float *)x0, ldx, T);
    DECAF_LAUNCH_CHECK();
    return 0;
}

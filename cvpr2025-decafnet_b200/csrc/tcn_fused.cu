// The whole refinement TCN in ONE launch (bf16 configuration), and the masked max-pool pyramid in one more.
//
// decaf_tcn_fused: expand (tcn_in) -> n_layers x DilatedResidualLayer -> conv_out for a tile of 256 level-0 steps of
// one query per CTA, with a recompute halo of 2^n_layers - 1 steps on both sides (the receptive field of the
// dilated stack), so no CTA ever needs a neighbour's data and any timeline length works (also the 70k-step MAD
// shape).  The 32-channel state lives in shared memory for the whole stack: fp32 for the residual / LayerNorm path
// (updated in place: an element is only ever read and written by the thread that owns it), a bf16 ping-pong copy as
// the tensor-core operand the dilated taps of OTHER warps read.  Per 16-step tile a warp issues 24 + 8 mma.sync
// (m16n8k16, bf16 x bf16 -> fp32): H = relu([x(t-d) | x(t) | x(t+d)] . Wd + bd), y = x + H . W1 + b1, then the
// mask multiply and the 32-channel LayerNorm on the accumulator fragments (4 lanes share a row).
// The layer's B fragments and per-channel vectors sit in registers for all the tiles of a layer.
// Replaces (bf16 configuration) decaf_tcn_in + n x decaf_tcn_layer + decaf_tcn_out: 10 launches and ~280 us of
// shared-memory-broadcast-bound fp32 FMAs at the NLQ shape (one thread per step, 1 LDS.128 per 4 FMAs).
// The fp32 configuration keeps the fp32-FMA kernels of tcn.cu.
//
// decaf_refine_pyramid: all L - 1 masked max-pool levels (libs/modeling/blocks.py:31-47) in one launch; a CTA owns
// 2^L level-0 steps plus a left halo of 2^(L-1) and walks the levels in shared memory.
#include "common.cuh"
#include "mma.cuh"

namespace decaf {

constexpr int TF_R = 32;
constexpr int TF_TL = 256;            // level-0 steps a CTA produces
constexpr int TF_THREADS = 256;
constexpr int TF_LDB = 40;            // bf16 row pitch (80 B: 16-byte aligned, conflict-free ldmatrix / fragment stores)

struct TcnFusedArgs {
    const float *logits1;
    const uint8_t *hmask;
    decaf_levels_t lv;
    const float *w_in, *b_in;
    const bf16 *wblob;                // per layer: Wd^T [32 cout][96 = tap * 32 + cin], W1^T [32 cout][32 cin]
    const float *vblob;               // per layer: bd[32], b1[32], ln_w[32], ln_b[32]
    const bf16 *w_out;                // [32 cout][32 cin]
    const float *b_out;
    bf16 *cat;
    int64_t ldc;
    int col0, n_layers, halo;
    float eps;
    // layer range [layer_lo, n_layers) of this launch: the stack may be split in two launches (see decaf_tcn_fused) that
    // hand the fp32 state of every step over through `state` (n_query, T, 32); state_in == nullptr: expand from the
    // logits (first launch), state_out == nullptr: conv_out into cat (last launch)
    int layer_lo;
    const float *state_in;
    float *state_out;
};

__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}

__global__ void __launch_bounds__(TF_THREADS, 1) tcn_fused_kernel(const __grid_constant__ TcnFusedArgs p) {
    extern __shared__ __align__(16) uint8_t tf_smem[];
    const int NR = TF_TL + 2 * p.halo;
    float *Xf = reinterpret_cast<float *>(tf_smem);                  // [NR][32] fp32 state
    bf16 *cur = reinterpret_cast<bf16 *>(tf_smem + (size_t)NR * TF_R * 4);   // [NR][TF_LDB] bf16 copy (ping)
    bf16 *nxt = cur + (size_t)NR * TF_LDB;                           // (pong)
    uint8_t *ms = reinterpret_cast<uint8_t *>(nxt + (size_t)NR * TF_LDB);    // [NR]: bit 0 mask, bit 1 inside [0, T)
    float *win_s = reinterpret_cast<float *>(ms + ((NR + 15) & ~15));        // [32 * L] + [32]
    const int L = p.lv.n_levels, T = p.lv.len[0];
    const int q = blockIdx.y, u0 = blockIdx.x * TF_TL, a = u0 - p.halo;
    const int64_t qrow = (int64_t)q * p.lv.Pp;
    for (int i = threadIdx.x; i < TF_R * L; i += TF_THREADS) win_s[(i % L) * TF_R + i / L] = p.w_in[i];   // [l][c]: conflict-free float4 reads
    if (threadIdx.x < TF_R) win_s[TF_R * L + threadIdx.x] = p.b_in[threadIdx.x];
    for (int r = threadIdx.x; r < NR; r += TF_THREADS) {
        const int t = a + r;
        const bool in = t >= 0 && t < T;
        ms[r] = in ? (uint8_t)(2 | (p.hmask[qrow + p.lv.off[0] + t] ? 1 : 0)) : (uint8_t)0;
    }
    __syncthreads();
    if (p.state_in != nullptr) {
        // ---- second launch of a split stack: the fp32 state of the region's steps (zero outside the sequence)
        const float *sin = p.state_in + (int64_t)q * T * TF_R;
        for (int idx = threadIdx.x; idx < NR * 8; idx += TF_THREADS) {
            const int r = idx >> 3, c4 = idx & 7, t = a + r;
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ms[r] & 2) o = *reinterpret_cast<const float4 *>(sin + (int64_t)t * TF_R + c4 * 4);
            *reinterpret_cast<float4 *>(Xf + r * TF_R + ((c4 * 4) ^ ((r & 3) << 3))) = o;
            uint2 pk;
            pk.x = pack_bf16(o.x, o.y); pk.y = pack_bf16(o.z, o.w);
            *reinterpret_cast<uint2 *>(cur + r * TF_LDB + c4 * 4) = pk;
        }
    } else
    // ---- expand: x0[t, c] = b_in[c] + sum_l w_in[c, l] * (l == 0 ? lg_0[t] : lg_l[t >> l] * m0[t])  (tcn_in_kernel)
    // the logits of every level under this region are staged first (coalesced loads, into the still unused pong
    // buffer): fetching them per element from global memory made this prologue half of the kernel's time
    {
        const float *lg = p.logits1 + qrow;
        float *stage = reinterpret_cast<float *>(nxt);
        const int tb = max(a, 0), te = min(a + NR, T);             // steps of the region inside the sequence
        int off = 0;
        for (int l = 0; l < L; l++) {
            const int s_l = tb >> l, cnt = te > tb ? ((te - 1) >> l) - s_l + 1 : 0;
            for (int i = threadIdx.x; i < cnt; i += TF_THREADS) stage[off + i] = lg[p.lv.off[l] + min(s_l + i, p.lv.len[l] - 1)];
            off += (NR >> l) + 2;
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < NR * 8; idx += TF_THREADS) {
            const int r = idx >> 3, c4 = idx & 7, t = a + r;
            float o[4] = {0.f, 0.f, 0.f, 0.f};
            if (ms[r] & 2) {
                const float m0 = (float)(ms[r] & 1);
#pragma unroll
                for (int j = 0; j < 4; j++) o[j] = win_s[TF_R * L + c4 * 4 + j];
                int ofs = 0;
                for (int l = 0; l < L; l++) {
                    const float v = stage[ofs + (t >> l) - (tb >> l)];
                    const float s = l == 0 ? v : v * m0;
                    const float4 w4 = *reinterpret_cast<const float4 *>(win_s + l * TF_R + c4 * 4);
                    o[0] = fmaf(w4.x, s, o[0]); o[1] = fmaf(w4.y, s, o[1]); o[2] = fmaf(w4.z, s, o[2]); o[3] = fmaf(w4.w, s, o[3]);
                    ofs += (NR >> l) + 2;
                }
            }
            *reinterpret_cast<float4 *>(Xf + r * TF_R + ((c4 * 4) ^ ((r & 3) << 3))) = make_float4(o[0], o[1], o[2], o[3]);
            uint2 pk;
            pk.x = pack_bf16(o[0], o[1]); pk.y = pack_bf16(o[2], o[3]);
            *reinterpret_cast<uint2 *>(cur + r * TF_LDB + c4 * 4) = pk;
        }
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
    const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lcol = (lane >> 4) * 8;      // ldmatrix.x4 address roles
    int rsum = (1 << p.n_layers) - (1 << p.layer_lo);
    for (int i = p.layer_lo; i < p.n_layers; i++) {
        const int d = 1 << i;
        rsum -= d;                                       // receptive radius of the layers still to come
        const int tile_lo = (p.halo - rsum) >> 4, tile_hi = (p.halo + TF_TL + rsum + 15) >> 4;
        const bf16 *wd = p.wblob + (size_t)i * (TF_R * 96 + TF_R * TF_R), *w1 = wd + TF_R * 96;
        const float *vb = p.vblob + (size_t)i * 4 * TF_R;
        uint32_t wdf[6][4][2], w1f[2][4][2];
#pragma unroll
        for (int kk = 0; kk < 6; kk++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const bf16 *wp = wd + (j * 8 + g) * 96 + kk * 16 + 2 * t4;
                wdf[kk][j][0] = *reinterpret_cast<const uint32_t *>(wp);
                wdf[kk][j][1] = *reinterpret_cast<const uint32_t *>(wp + 8);
            }
#pragma unroll
        for (int kk = 0; kk < 2; kk++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const bf16 *wp = w1 + (j * 8 + g) * TF_R + kk * 16 + 2 * t4;
                w1f[kk][j][0] = *reinterpret_cast<const uint32_t *>(wp);
                w1f[kk][j][1] = *reinterpret_cast<const uint32_t *>(wp + 8);
            }
        float bdv[4][2], b1v[4][2], lw[4][2], lb[4][2];
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int c = j * 8 + 2 * t4 + e;
                bdv[j][e] = vb[c]; b1v[j][e] = vb[TF_R + c]; lw[j][e] = vb[2 * TF_R + c]; lb[j][e] = vb[3 * TF_R + c];
            }
        // two independent 16-step tiles per iteration: with 8 warps per SM the mma / shuffle / shared-memory latency
        // chains of a single tile leave the schedulers idle most of the time
        for (int tile = tile_lo + 2 * warp; tile < tile_hi; tile += 2 * (TF_THREADS / 32)) {
            const int nt = min(2, tile_hi - tile);
            float h[2][4][4];
#pragma unroll
            for (int u = 0; u < 2; u++)
#pragma unroll
                for (int j = 0; j < 4; j++) { h[u][j][0] = h[u][j][2] = bdv[j][0]; h[u][j][1] = h[u][j][3] = bdv[j][1]; }
#pragma unroll
            for (int tap = 0; tap < 3; tap++) {
#pragma unroll
                for (int half = 0; half < 2; half++) {
                    uint32_t af[2][4];
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        // rows outside the region are never needed by a row that matters (halo = receptive field): clamp
                        const int rr = min(max((tile + u) * 16 + lrow + (tap - 1) * d, 0), NR - 1);
                        ldmatrix_x4(af[u], cur + rr * TF_LDB + half * 16 + lcol);
                    }
#pragma unroll
                    for (int u = 0; u < 2; u++)
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            mma_bf16(h[u][j], af[u][0], af[u][1], af[u][2], af[u][3], wdf[tap * 2 + half][j][0], wdf[tap * 2 + half][j][1]);
                }
            }
            float y[2][4][4];
#pragma unroll
            for (int u = 0; u < 2; u++)
#pragma unroll
                for (int j = 0; j < 4; j++) { y[u][j][0] = y[u][j][2] = b1v[j][0]; y[u][j][1] = y[u][j][3] = b1v[j][1]; }
#pragma unroll
            for (int kk = 0; kk < 2; kk++) {
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    const uint32_t a0 = pack_bf16(fmaxf(h[u][2 * kk][0], 0.f), fmaxf(h[u][2 * kk][1], 0.f));
                    const uint32_t a1 = pack_bf16(fmaxf(h[u][2 * kk][2], 0.f), fmaxf(h[u][2 * kk][3], 0.f));
                    const uint32_t a2 = pack_bf16(fmaxf(h[u][2 * kk + 1][0], 0.f), fmaxf(h[u][2 * kk + 1][1], 0.f));
                    const uint32_t a3 = pack_bf16(fmaxf(h[u][2 * kk + 1][2], 0.f), fmaxf(h[u][2 * kk + 1][3], 0.f));
#pragma unroll
                    for (int j = 0; j < 4; j++) mma_bf16(y[u][j], a0, a1, a2, a3, w1f[kk][j][0], w1f[kk][j][1]);
                }
            }
            // residual, mask, LayerNorm(32) (two-pass, biased variance: nn.LayerNorm), rows ra = r0 + g and rb = ra + 8.
            // Xf columns are XOR-swizzled by (row & 3) << 3 so the 8 rows of a fragment hit different banks.
#pragma unroll
            for (int u = 0; u < 2; u++) {
                if (u < nt) {
                    const int ra = (tile + u) * 16 + g, rb = ra + 8;
                    const int sw = (ra & 3) << 3;                    // rb = ra + 8: same swizzle
                    const float ma = (float)(ms[ra] & 1), mb = (float)(ms[rb] & 1);
                    const bool ina = ms[ra] & 2, inb = ms[rb] & 2;
                    float *xa = Xf + ra * TF_R + 2 * t4, *xb = Xf + rb * TF_R + 2 * t4;
                    float sa = 0.f, sb = 0.f;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const float2 va = *reinterpret_cast<const float2 *>(xa + ((j * 8) ^ sw));
                        const float2 vb2 = *reinterpret_cast<const float2 *>(xb + ((j * 8) ^ sw));
                        y[u][j][0] = (y[u][j][0] + va.x) * ma; y[u][j][1] = (y[u][j][1] + va.y) * ma;
                        y[u][j][2] = (y[u][j][2] + vb2.x) * mb; y[u][j][3] = (y[u][j][3] + vb2.y) * mb;
                        sa += y[u][j][0] + y[u][j][1];
                        sb += y[u][j][2] + y[u][j][3];
                    }
                    const float mean_a = quad_sum(sa) * (1.0f / TF_R), mean_b = quad_sum(sb) * (1.0f / TF_R);
                    float qa = 0.f, qb = 0.f;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        y[u][j][0] -= mean_a; y[u][j][1] -= mean_a; y[u][j][2] -= mean_b; y[u][j][3] -= mean_b;
                        qa = fmaf(y[u][j][0], y[u][j][0], fmaf(y[u][j][1], y[u][j][1], qa));
                        qb = fmaf(y[u][j][2], y[u][j][2], fmaf(y[u][j][3], y[u][j][3], qb));
                    }
                    const float rsa = ina ? 1.0f / sqrtf(quad_sum(qa) * (1.0f / TF_R) + p.eps) : 0.f;
                    const float rsb = inb ? 1.0f / sqrtf(quad_sum(qb) * (1.0f / TF_R) + p.eps) : 0.f;
                    bf16 *na = nxt + ra * TF_LDB + 2 * t4, *nb = nxt + rb * TF_LDB + 2 * t4;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        // steps outside [0, T) stay exactly zero: they are the zero padding of the next layer's dilated taps
                        const float oa0 = ina ? fmaf(y[u][j][0] * rsa, lw[j][0], lb[j][0]) : 0.f;
                        const float oa1 = ina ? fmaf(y[u][j][1] * rsa, lw[j][1], lb[j][1]) : 0.f;
                        const float ob0 = inb ? fmaf(y[u][j][2] * rsb, lw[j][0], lb[j][0]) : 0.f;
                        const float ob1 = inb ? fmaf(y[u][j][3] * rsb, lw[j][1], lb[j][1]) : 0.f;
                        *reinterpret_cast<float2 *>(xa + ((j * 8) ^ sw)) = make_float2(oa0, oa1);
                        *reinterpret_cast<float2 *>(xb + ((j * 8) ^ sw)) = make_float2(ob0, ob1);
                        *reinterpret_cast<uint32_t *>(na + j * 8) = pack_bf16(oa0, oa1);
                        *reinterpret_cast<uint32_t *>(nb + j * 8) = pack_bf16(ob0, ob1);
                    }
                }
            }
        }
        __syncthreads();
        bf16 *tmp = cur; cur = nxt; nxt = tmp;
    }
    if (p.state_out != nullptr) {
        // ---- first launch of a split stack: the fp32 state of the CTA's own steps goes to the hand-over buffer
        float *sout = p.state_out + (int64_t)q * T * TF_R;
        for (int idx = threadIdx.x; idx < TF_TL * 8; idx += TF_THREADS) {
            const int r = p.halo + (idx >> 3), c4 = idx & 7, t = a + r;
            if (t < T)
                *reinterpret_cast<float4 *>(sout + (int64_t)t * TF_R + c4 * 4) =
                    *reinterpret_cast<const float4 *>(Xf + r * TF_R + ((c4 * 4) ^ ((r & 3) << 3)));
        }
        return;
    }
    // ---- conv_out: y = (W_out x + b_out) * m -> cat[q, off0 + t, col0 : col0 + 32]  (own 256 steps only)
    {
        uint32_t wof[2][4][2];
#pragma unroll
        for (int kk = 0; kk < 2; kk++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const bf16 *wp = p.w_out + (j * 8 + g) * TF_R + kk * 16 + 2 * t4;
                wof[kk][j][0] = *reinterpret_cast<const uint32_t *>(wp);
                wof[kk][j][1] = *reinterpret_cast<const uint32_t *>(wp + 8);
            }
        const int tile_lo = p.halo >> 4, tile_hi = (p.halo + TF_TL) >> 4;
        for (int tile = tile_lo + warp; tile < tile_hi; tile += TF_THREADS / 32) {
            const int r0 = tile * 16;
            float y[4][4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                y[j][0] = y[j][2] = p.b_out[j * 8 + 2 * t4];
                y[j][1] = y[j][3] = p.b_out[j * 8 + 2 * t4 + 1];
            }
#pragma unroll
            for (int half = 0; half < 2; half++) {
                uint32_t af[4];
                ldmatrix_x4(af, cur + (r0 + lrow) * TF_LDB + half * 16 + lcol);
#pragma unroll
                for (int j = 0; j < 4; j++) mma_bf16(y[j], af[0], af[1], af[2], af[3], wof[half][j][0], wof[half][j][1]);
            }
            const int ra = r0 + g, rb = ra + 8;
            const float ma = (float)(ms[ra] & 1), mb = (float)(ms[rb] & 1);
            bf16 *da = p.cat + (qrow + p.lv.off[0] + a + ra) * p.ldc + p.col0 + 2 * t4;
            bf16 *db = p.cat + (qrow + p.lv.off[0] + a + rb) * p.ldc + p.col0 + 2 * t4;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (ms[ra] & 2) *reinterpret_cast<uint32_t *>(da + j * 8) = pack_bf16(y[j][0] * ma, y[j][1] * ma);
                if (ms[rb] & 2) *reinterpret_cast<uint32_t *>(db + j * 8) = pack_bf16(y[j][2] * mb, y[j][3] * mb);
            }
        }
    }
}

// ------------------------------------------------------------------------------- pooling pyramid
template <typename TA> struct Vec16;       // 16 bytes of TA
template <> struct Vec16<float> { static constexpr int N = 4; };
template <> struct Vec16<bf16> { static constexpr int N = 8; };

template <typename TA>
__device__ __forceinline__ void store4(TA *dst, const float4 &v);
template <> __device__ __forceinline__ void store4<float>(float *dst, const float4 &v) { *reinterpret_cast<float4 *>(dst) = v; }
template <> __device__ __forceinline__ void store4<bf16>(bf16 *dst, const float4 &v) {
    uint2 pk;
    pk.x = pack_bf16(v.x, v.y); pk.y = pack_bf16(v.z, v.w);
    *reinterpret_cast<uint2 *>(dst) = pk;
}

template <typename TA>
__global__ void __launch_bounds__(256)
refine_pyramid_kernel(TA *__restrict__ cat, int64_t ldc, int col0, const uint8_t *__restrict__ hmask, decaf_levels_t lv,
                      int halo, int tile) {
    extern __shared__ __align__(16) uint8_t rp_smem[];
    const int n0 = tile + halo;
    float *b0 = reinterpret_cast<float *>(rp_smem);          // [n0][32]
    float *b1 = b0 + (size_t)n0 * TF_R;                      // [n0 / 2][32]
    uint8_t *m0s = reinterpret_cast<uint8_t *>(b1 + (size_t)(n0 / 2) * TF_R);
    const int L = lv.n_levels;
    const int q = blockIdx.y, a = blockIdx.x * tile - halo;
    const int64_t qrow = (int64_t)q * lv.Pp;
    constexpr int VN = Vec16<TA>::N, CPR = TF_R / VN;        // 16-byte chunks per row
#pragma unroll 4
    for (int idx = threadIdx.x; idx < n0 * CPR; idx += blockDim.x) {
        const int r = idx / CPR, c = (idx % CPR) * VN, t = a + r;
        const bool in = t >= 0 && t < lv.len[0];
        float v[VN];
        if (in) {
            const uint4 raw = *reinterpret_cast<const uint4 *>(cat + (qrow + lv.off[0] + t) * ldc + col0 + c);
            const TA *e = reinterpret_cast<const TA *>(&raw);
#pragma unroll
            for (int i = 0; i < VN; i++) v[i] = to_f32<TA>(e[i]);
        } else {
#pragma unroll
            for (int i = 0; i < VN; i++) v[i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < VN; i += 4) *reinterpret_cast<float4 *>(b0 + r * TF_R + c + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    }
    for (int r = threadIdx.x; r < n0; r += blockDim.x) {
        const int t = a + r;
        m0s[r] = (t >= 0 && t < lv.len[0]) ? hmask[qrow + lv.off[0] + t] : 0;
    }
    __syncthreads();
    float *src = b0, *dst = b1;
    for (int l = 1; l < L; l++) {
        const int nl = n0 >> l, np = n0 >> (l - 1);
        const int al = a >> l, ap = a >> (l - 1);            // exact: a is a multiple of halo = 2^(L-1)
        const int own0 = halo >> l;
#pragma unroll 2
        for (int idx = threadIdx.x; idx < nl * 8; idx += blockDim.x) {
            const int i = idx >> 3, c = (idx & 7) * 4;
            float4 best = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
            bool any = false;
#pragma unroll
            for (int j = -1; j <= 1; j++) {
                const int ci = 2 * i + j;
                if (ci < 0 || ci >= np) continue;
                const int tc = ap + ci;
                if (tc < 0 || tc >= lv.len[l - 1]) continue;
                if (!m0s[ci << (l - 1)]) continue;           // mask_{l-1}[tc] = mask_0[tc << (l-1)]
                const float4 x = *reinterpret_cast<const float4 *>(src + ci * TF_R + c);
                best.x = fmaxf(best.x, x.x); best.y = fmaxf(best.y, x.y); best.z = fmaxf(best.z, x.z); best.w = fmaxf(best.w, x.w);
                any = true;
            }
            const int tg = al + i;
            const bool inl = tg >= 0 && tg < lv.len[l];
            if (!(any && inl && m0s[i << l])) best = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4 *>(dst + i * TF_R + c) = best;
            if (i >= own0 && inl) store4<TA>(cat + (qrow + lv.off[l] + tg) * ldc + col0 + c, best);
        }
        __syncthreads();
        float *tmp = src; src = dst; dst = tmp;
    }
}

}  // namespace decaf

using namespace decaf;

extern "C" int decaf_tcn_fused_supported(int32_t n_layers, int32_t n_levels) {
    return n_layers >= 1 && n_layers <= 8 && n_levels >= 1 && n_levels <= DECAF_MAX_LEVELS;
}

static int tcn_fused_launch(TcnFusedArgs a, int layer_lo, int layer_hi, const float *state_in, float *state_out, int n_query,
                            cudaStream_t st) {
    a.layer_lo = layer_lo; a.n_layers = layer_hi; a.state_in = state_in; a.state_out = state_out;
    a.halo = (((1 << layer_hi) - (1 << layer_lo)) + 15) / 16 * 16;        // receptive field of the launch's layers
    const int NR = TF_TL + 2 * a.halo;
    const size_t smem = (size_t)NR * (TF_R * 4 + 2 * TF_LDB * 2) + ((NR + 15) & ~15) + (size_t)(TF_R * a.lv.n_levels + TF_R) * 4;
    static size_t attr = 0;
    if (smem > attr) {
        DECAF_CUDA(cudaFuncSetAttribute(tcn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = smem;
    }
    dim3 grid(cdiv(a.lv.len[0], TF_TL), n_query);
    tcn_fused_kernel<<<grid, TF_THREADS, smem, st>>>(a);
    DECAF_LAUNCH_CHECK();
    return 0;
}

extern "C" int decaf_tcn_fused(const float *logits1, const uint8_t *hmask, const decaf_levels_t *lv, const float *w_in,
                               const float *b_in, const void *wblob, const float *vblob, int32_t n_layers,
                               const void *w_out, const float *b_out, int32_t R, float eps, void *cat, int64_t ldc,
                               int32_t col0, int32_t n_query, float *scratch, void *stream) {
    DECAF_CHECK(logits1 && hmask && lv && w_in && b_in && wblob && vblob && w_out && b_out && cat, "decaf_tcn_fused: null pointers");
    DECAF_CHECK(R == TF_R, "decaf_tcn_fused: refine width must be %d (got %d)", TF_R, R);
    DECAF_CHECK(decaf_tcn_fused_supported(n_layers, lv->n_levels), "decaf_tcn_fused: unsupported depth %d", n_layers);
    DECAF_CHECK(ldc % 2 == 0 && col0 % 2 == 0 && (reinterpret_cast<uintptr_t>(cat) & 3) == 0, "decaf_tcn_fused: cat must allow 4-byte stores");
    DECAF_CHECK(!scratch || (reinterpret_cast<uintptr_t>(scratch) & 15) == 0, "decaf_tcn_fused: scratch must be 16-byte aligned");
    if (n_query == 0 || lv->len[0] == 0) return 0;
    TcnFusedArgs a;
    a.logits1 = logits1; a.hmask = hmask; a.lv = *lv; a.w_in = w_in; a.b_in = b_in;
    a.wblob = reinterpret_cast<const bf16 *>(wblob); a.vblob = vblob;
    a.w_out = reinterpret_cast<const bf16 *>(w_out); a.b_out = b_out;
    a.cat = reinterpret_cast<bf16 *>(cat); a.ldc = ldc; a.col0 = col0; a.eps = eps;
    cudaStream_t st = as_stream(stream);
    // With a hand-over buffer (n_query x T x 32 fp32) a deep stack runs as TWO launches: a tile's recompute halo is the
    // receptive field of the launch's own layers only - 31 steps for dilations 1..16 and 224 for 32..128 instead of 255 for
    // all eight - which takes the row-layers computed per 256 output steps from 5124 to 2884 and the first launch's CTAs
    // down to 92 KB of shared memory (other lanes' kernels fit beside them).  Per-step arithmetic is unchanged (bit-identical
    // output).  Measured at the NLQ shape: 35.8 + 30.0 us against 67.9 us for the single launch - the kernel is bound by
    // per-layer latency (two __syncthreads and a handful of dependent tile iterations per layer), not by the halo rows;
    // staging the launch's layer weights in shared memory changed nothing (36.1 + 30.6 us).
    if (scratch != nullptr && n_layers >= 6) {
        const int split = n_layers - 3;
        if (tcn_fused_launch(a, 0, split, nullptr, scratch, n_query, st)) return 1;
        return tcn_fused_launch(a, split, n_layers, scratch, nullptr, n_query, st);
    }
    return tcn_fused_launch(a, 0, n_layers, nullptr, nullptr, n_query, st);
}

extern "C" int decaf_refine_pyramid_supported(int32_t n_levels) { return n_levels >= 2 && n_levels <= 9; }

extern "C" int decaf_refine_pyramid(void *cat, int32_t dtype, int64_t ldc, int32_t col0, int32_t R, const uint8_t *hmask,
                                    const decaf_levels_t *lv, int32_t n_query, void *stream) {
    DECAF_CHECK(cat && hmask && lv, "decaf_refine_pyramid: null pointers");
    DECAF_CHECK(R == TF_R, "decaf_refine_pyramid: refine width must be %d (got %d)", TF_R, R);
    DECAF_CHECK(dtype == DECAF_BF16 ? (ldc % 8 == 0 && col0 % 8 == 0) : (ldc % 4 == 0 && col0 % 4 == 0), "decaf_refine_pyramid: ldc / col0 must keep 16-byte alignment");
    DECAF_CHECK((reinterpret_cast<uintptr_t>(cat) & 15) == 0, "decaf_refine_pyramid: cat must be 16-byte aligned");
    DECAF_CHECK(decaf_refine_pyramid_supported(lv->n_levels), "decaf_refine_pyramid: unsupported level count %d", lv->n_levels);
    for (int l = 1; l < lv->n_levels; l++)
        DECAF_CHECK(lv->len[l] * 2 == lv->len[l - 1], "decaf_refine_pyramid: level lengths must halve exactly");
    if (n_query == 0 || lv->len[0] == 0) return 0;
    const int halo = 1 << (lv->n_levels - 1);
    const int tile = 2 * halo < 128 ? 128 / halo * halo : 2 * halo;
    const int n0 = tile + halo;
    const size_t smem = (size_t)(n0 + n0 / 2) * TF_R * 4 + n0;
    cudaStream_t st = as_stream(stream);
    dim3 grid(cdiv(lv->len[0], tile), n_query);
    static size_t attr_b = 0, attr_f = 0;
    if (dtype == DECAF_BF16) {
        if (smem > 48 * 1024 && smem > attr_b) {
            DECAF_CUDA(cudaFuncSetAttribute(refine_pyramid_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_b = smem;
        }
        refine_pyramid_kernel<bf16><<<grid, 256, smem, st>>>((bf16 *)cat, ldc, col0, hmask, *lv, halo, tile);
    } else {
        if (smem > 48 * 1024 && smem > attr_f) {
            DECAF_CUDA(cudaFuncSetAttribute(refine_pyramid_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_f = smem;
        }
        refine_pyramid_kernel<float><<<grid, 256, smem, st>>>((float *)cat, ldc, col0, hmask, *lv, halo, tile);
    }
    DECAF_LAUNCH_CHECK();
    return 0;
}

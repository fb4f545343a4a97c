// Banded local-window self-attention and short-key global (cross) attention.
// One warp per query time step; lane = head * (32 / n_heads) + sub, each lane owns C/32
// contiguous channels of its head, so a warp reads whole 128B-multiple rows (coalesced) and a
// score needs log2(32 / n_heads) shuffles.  Softmax is computed online in fp32.
// The reference materialises (2s x 2s) chunk products and a second "mask matmul"
// (libs/modeling/blocks.py:224-325); here the band is evaluated directly.
#include "common.cuh"
#include "mma.cuh"

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


// ---------------------------------------------------------------------------------- tensor-core variants (bf16)
constexpr int XM_WARPS = 8;                  // 16 query rows per warp -> 128 query rows per CTA (4 warps when C is wide)

// Cross attention over <= 64 text keys (libs/modeling/blocks.py:374-389, global branch with a -inf key mask).
// One CTA = 128 query rows of one sequence, all heads.  The Q tile is fetched with 16-byte cp.async (whole 512-byte
// rows, every byte of the CTA's input in flight at once) while the threads convert the sequence's K (fp32 -> bf16,
// [key][C + 8]) and V (transposed, [C][LKP + 8]) into shared memory (padded rows: conflict-free fragment reads);
// S = Q.K^T, softmax and O = P.V run on mma.sync tiles with ldmatrix A fragments; a warp writes O over the Q columns
// of the head it has just consumed and streams its 16 finished rows out with 16-byte coalesced stores.  (The first
// tensor-core version loaded Q fragments with 4-byte global loads head by head: one CTA of 8 warps per SM, every
// head a dependent load -> 37 us at the NLQ shape against 6 us of HBM time.)
// PACKED: k points at the per-sequence shared-memory image written once by xattn_pack_kv_kernel (decaf_xattn_pack_kv)
// and the CTA copies it with 16-byte cp.async; otherwise every CTA converts the fp32 K/V rows itself (~10 us of
// dependent L2 loads per CTA: measured 16 us for a 16-CTA launch, 23 of the 37 us of the NLQ video launch).
template <int HD, int NKT, bool PACKED, int NW>
__global__ void __launch_bounds__(32 * NW)
xattn_mma_kernel(const bf16 *__restrict__ q, const float *__restrict__ k, const float *__restrict__ v, bf16 *__restrict__ out,
                 int Tq, int Lk, int C, int n_heads, float scale2, const int32_t *__restrict__ kv_len) {
    constexpr int LKP = 8 * NKT;
    extern __shared__ __align__(16) uint8_t xsm[];
    const int ldk = C + 8, ldv = LKP + 8, ldq = C + 8;
    bf16 *Ks = reinterpret_cast<bf16 *>(xsm);                 // [LKP][ldk]
    bf16 *Vt = Ks + LKP * ldk;                                // [C][ldv]
    bf16 *Qs = Vt + C * ldv;                                  // [16 * NW][ldq]
    const int seq = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int r0 = blockIdx.x * (16 * NW) + warp * 16;  // first query row of this warp
    const int cpr = C / 8;                                    // 16-byte chunks per row
    {
        // this warp's 16 Q rows (rows past the end of the sequence are zero-filled)
        bf16 *qw = Qs + warp * 16 * ldq;
        int r = lane / cpr, c = lane - r * cpr;               // (row, chunk) advance incrementally: no division per piece
        const int dr = 32 / cpr, dc = 32 - dr * cpr;
        while (r < 16) {
            const bool ok = r0 + r < Tq;
            cp_async16(qw + r * ldq + c * 8, q + ((int64_t)seq * Tq + (ok ? r0 + r : 0)) * C + c * 8, ok);
            r += dr; c += dc;
            if (c >= cpr) { c -= cpr; r++; }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    const int n_kv = min(kv_len[seq], Lk);
    if constexpr (PACKED) {
        const int n16 = (LKP * ldk + C * ldv) / 8;            // 16-byte pieces of the packed image
        const bf16 *src = reinterpret_cast<const bf16 *>(k) + (int64_t)seq * n16 * 8;
        for (int i = threadIdx.x; i < n16; i += blockDim.x) cp_async16(Ks + i * 8, src + i * 8, true);
        asm volatile("cp.async.commit_group;" ::: "memory");
    } else {
        const float *kb = k + (int64_t)seq * Lk * C, *vb = v + (int64_t)seq * Lk * C;
        const int c4n = C / 4;
#pragma unroll 4
        for (int i = threadIdx.x; i < LKP * c4n; i += blockDim.x) {
            const int key = i / c4n, ch = (i - key * c4n) * 4;
            float4 kk = make_float4(0.f, 0.f, 0.f, 0.f);
            if (key < n_kv) kk = *reinterpret_cast<const float4 *>(kb + (int64_t)key * C + ch);
            uint2 pk;
            pk.x = pack_bf16(kk.x, kk.y); pk.y = pack_bf16(kk.z, kk.w);
            *reinterpret_cast<uint2 *>(Ks + key * ldk + ch) = pk;
        }
#pragma unroll 4
        for (int i = threadIdx.x; i < (LKP / 2) * C; i += blockDim.x) {     // two keys per thread: one 4-byte store
            const int kp = i / C, ch = i - kp * C;
            const float v0 = 2 * kp < n_kv ? vb[(int64_t)(2 * kp) * C + ch] : 0.f;
            const float v1 = 2 * kp + 1 < n_kv ? vb[(int64_t)(2 * kp + 1) * C + ch] : 0.f;
            *reinterpret_cast<uint32_t *>(Vt + ch * ldv + 2 * kp) = pack_bf16(v0, v1);
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (r0 >= Tq) return;
    const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lcol = (lane >> 4) * 8;
    bf16 *qw = Qs + warp * 16 * ldq;
    const float sl2 = scale2 * 1.4426950408889634f;            // scores in log2 units -> exp2
    for (int h = 0; h < n_heads; h++) {
        const int c0 = h * HD;
        float s[NKT][4];
#pragma unroll
        for (int j = 0; j < NKT; j++) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
        uint32_t aq[HD / 16][4];
#pragma unroll
        for (int kk = 0; kk < HD / 16; kk++) ldmatrix_x4(aq[kk], qw + lrow * ldq + c0 + kk * 16 + lcol);
#pragma unroll
        for (int kk = 0; kk < HD / 16; kk++) {
#pragma unroll
            for (int j = 0; j < NKT; j++) {
                const bf16 *kp = Ks + (j * 8 + g) * ldk + c0 + kk * 16 + 2 * t;
                mma_bf16(s[j], aq[kk][0], aq[kk][1], aq[kk][2], aq[kk][3], *reinterpret_cast<const uint32_t *>(kp),
                         *reinterpret_cast<const uint32_t *>(kp + 8));
            }
        }
        float ma = -INFINITY, mb = -INFINITY;
#pragma unroll
        for (int j = 0; j < NKT; j++) {
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int col = j * 8 + 2 * t + (e & 1);
                s[j][e] = col < n_kv ? s[j][e] * sl2 : -INFINITY;
            }
            ma = fmaxf(ma, fmaxf(s[j][0], s[j][1]));
            mb = fmaxf(mb, fmaxf(s[j][2], s[j][3]));
        }
        ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 1)); ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 2));
        mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 1)); mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 2));
        if (ma == -INFINITY) ma = 0.f;                          // no key at all: every weight becomes exp2(-inf) = 0
        if (mb == -INFINITY) mb = 0.f;
        float la = 0.f, lb = 0.f;
#pragma unroll
        for (int j = 0; j < NKT; j++) {
            s[j][0] = round_bf16(exp2f(s[j][0] - ma)); s[j][1] = round_bf16(exp2f(s[j][1] - ma));
            s[j][2] = round_bf16(exp2f(s[j][2] - mb)); s[j][3] = round_bf16(exp2f(s[j][3] - mb));
            la += s[j][0] + s[j][1];
            lb += s[j][2] + s[j][3];
        }
        la += __shfl_xor_sync(0xffffffffu, la, 1); la += __shfl_xor_sync(0xffffffffu, la, 2);
        lb += __shfl_xor_sync(0xffffffffu, lb, 1); lb += __shfl_xor_sync(0xffffffffu, lb, 2);
        float o[HD / 8][4];
#pragma unroll
        for (int n = 0; n < HD / 8; n++) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
#pragma unroll
        for (int kk = 0; kk < NKT / 2; kk++) {
            const uint32_t a0 = pack_bf16(s[2 * kk][0], s[2 * kk][1]), a1 = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
            const uint32_t a2 = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]), a3 = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
            for (int n = 0; n < HD / 8; n++) {
                const bf16 *vp = Vt + (c0 + n * 8 + g) * ldv + kk * 16 + 2 * t;
                mma_bf16(o[n], a0, a1, a2, a3, *reinterpret_cast<const uint32_t *>(vp), *reinterpret_cast<const uint32_t *>(vp + 8));
            }
        }
        const float ia = la > 0.f ? 1.0f / la : 0.f, ib = lb > 0.f ? 1.0f / lb : 0.f;
        __syncwarp();                                           // every lane has its Q fragments of this head
        bf16 *oa = qw + g * ldq + c0 + 2 * t, *ob = oa + 8 * ldq;
#pragma unroll
        for (int n = 0; n < HD / 8; n++) {
            *reinterpret_cast<uint32_t *>(oa + n * 8) = pack_bf16(o[n][0] * ia, o[n][1] * ia);
            *reinterpret_cast<uint32_t *>(ob + n * 8) = pack_bf16(o[n][2] * ib, o[n][3] * ib);
        }
    }
    __syncwarp();
    {
        int r = lane / cpr, c = lane - r * cpr;
        const int dr = 32 / cpr, dc = 32 - dr * cpr;
        while (r < 16) {
            if (r0 + r < Tq)
                *reinterpret_cast<uint4 *>(out + ((int64_t)seq * Tq + r0 + r) * C + c * 8) = *reinterpret_cast<const uint4 *>(qw + r * ldq + c * 8);
            r += dr; c += dc;
            if (c >= cpr) { c -= cpr; r++; }
        }
    }
}


// Banded local attention on mma.sync tiles (libs/modeling/blocks.py:357-373 == a band |i - j| <= s with -inf outside
// the sequence, -1e4 added to masked keys, masked query rows zeroed).  One CTA = 64 query steps (4 warps x 16) of ONE
// head of one sequence; the K and V head slices of the 64 + 2s (+ tile padding) steps around it are staged in shared
// memory with cp.async (rows of HD + 8 bf16: conflict-free B fragments for Q.K^T, ldmatrix.trans for P.V).  A warp
// evaluates its 16 queries against the 8 * NKT keys starting s steps before its first query.
// `phase` (0..15) shifts the 16-row warp tiles to start at step -phase: which MMA column a key lands in — and with it the
// fp32 summation order of the softmax denominator and of P.V — depends on the query's position inside its 16-row tile, so a
// time shard whose window starts at global step w0 passes phase = w0 mod 16 and every row is computed in exactly the
// arithmetic of the unsharded run (the difference is ~1 ulp before the bf16 rounding of the output, i.e. a rare flipped
// bf16 bit that then shows up in one row of the residual stream).
constexpr int LM_WARPS = 4;

template <int HD, int NKT>
__global__ void __launch_bounds__(32 * LM_WARPS)
local_attn_mma_kernel(const bf16 *__restrict__ q, const bf16 *__restrict__ k, const bf16 *__restrict__ v, bf16 *__restrict__ out,
                      int T, int C, int s, float scale2, const uint8_t *__restrict__ mask, int64_t m_seq_stride, int phase) {
    constexpr int NK = 8 * NKT;                              // keys a warp looks at (>= 16 + 2s)
    constexpr int ROWS = 16 * (LM_WARPS - 1) + NK;           // staged steps
    constexpr int LD = HD + 8;
    __shared__ __align__(16) bf16 Ks[ROWS * LD];
    __shared__ __align__(16) bf16 Vs[ROWS * LD];
    __shared__ uint8_t km[ROWS];
    const int seq = blockIdx.z, head = blockIdx.y, t0 = blockIdx.x * (16 * LM_WARPS) - phase;
    const int c0 = head * HD;
    const int64_t base = (int64_t)seq * T;
    const uint8_t *mrow = mask + (int64_t)seq * m_seq_stride;
    constexpr int CPR = HD / 8;                              // 16-byte chunks per staged row
    for (int i = threadIdx.x; i < ROWS * CPR; i += blockDim.x) {
        const int r = i / CPR, c = i - r * CPR;
        const int tk = t0 - s + r;
        const bool ok = tk >= 0 && tk < T;
        const int64_t off = (base + (ok ? tk : 0)) * C + c0 + c * 8;
        cp_async16(Ks + r * LD + c * 8, k + off, ok);
        cp_async16(Vs + r * LD + c * 8, v + off, ok);
    }
    for (int r = threadIdx.x; r < ROWS; r += blockDim.x) {
        const int tk = t0 - s + r;
        km[r] = (tk >= 0 && tk < T) ? mrow[tk] : 0;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int w0 = t0 + warp * 16;
    const int ra = w0 + g, rb = w0 + g + 8;
    const bool va = ra >= 0 && ra < T, vb = rb >= 0 && rb < T;
    // Q fragments straight from global memory (each thread: 4-byte pieces of rows ra / rb)
    const bf16 *qa = q + (base + (va ? ra : 0)) * C + c0 + 2 * t;
    const bf16 *qb = q + (base + (vb ? rb : 0)) * C + c0 + 2 * t;
    uint32_t aq[HD / 16][4];
#pragma unroll
    for (int kk = 0; kk < HD / 16; kk++) {
        aq[kk][0] = *reinterpret_cast<const uint32_t *>(qa + kk * 16);
        aq[kk][1] = *reinterpret_cast<const uint32_t *>(qb + kk * 16);
        aq[kk][2] = *reinterpret_cast<const uint32_t *>(qa + kk * 16 + 8);
        aq[kk][3] = *reinterpret_cast<const uint32_t *>(qb + kk * 16 + 8);
    }
    const bool qma = va && mrow[va ? ra : 0], qmb = vb && mrow[vb ? rb : 0];
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (w0 >= T || w0 + 16 <= 0) return;
    const bf16 *Kw = Ks + warp * 16 * LD, *Vw = Vs + warp * 16 * LD;
    const uint8_t *kmw = km + warp * 16;
    float sc[NKT][4];
#pragma unroll
    for (int j = 0; j < NKT; j++) { sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < HD / 16; kk++) {
#pragma unroll
        for (int j = 0; j < NKT; j++) {
            const bf16 *kp = Kw + (j * 8 + g) * LD + kk * 16 + 2 * t;
            mma_bf16(sc[j], aq[kk][0], aq[kk][1], aq[kk][2], aq[kk][3], *reinterpret_cast<const uint32_t *>(kp),
                     *reinterpret_cast<const uint32_t *>(kp + 8));
        }
    }
    // key kl of the warp's window is step w0 - s + kl; query row g (+8) sees kl in [g (+8), g (+8) + 2s]
    constexpr float L2E = 1.4426950408889634f;
    const float sl2 = scale2 * L2E, mbias = -1e4f * L2E;
    float ma = -INFINITY, mb = -INFINITY;
#pragma unroll
    for (int j = 0; j < NKT; j++) {
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int kl = j * 8 + 2 * t + (e & 1);
            const int lo = g + ((e & 2) ? 8 : 0);
            const int tk = w0 - s + kl;
            const bool in = kl >= lo && kl <= lo + 2 * s && tk >= 0 && tk < T;
            sc[j][e] = in ? fmaf(sc[j][e], sl2, kmw[kl] ? 0.f : mbias) : -INFINITY;
        }
        ma = fmaxf(ma, fmaxf(sc[j][0], sc[j][1]));
        mb = fmaxf(mb, fmaxf(sc[j][2], sc[j][3]));
    }
    ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 1)); ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 2));
    mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 1)); mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 2));
    if (ma == -INFINITY) ma = 0.f;                           // rows past the end of the sequence
    if (mb == -INFINITY) mb = 0.f;
    float la = 0.f, lb = 0.f;
#pragma unroll
    for (int j = 0; j < NKT; j++) {
        sc[j][0] = round_bf16(exp2f(sc[j][0] - ma)); sc[j][1] = round_bf16(exp2f(sc[j][1] - ma));
        sc[j][2] = round_bf16(exp2f(sc[j][2] - mb)); sc[j][3] = round_bf16(exp2f(sc[j][3] - mb));
        la += sc[j][0] + sc[j][1];
        lb += sc[j][2] + sc[j][3];
    }
    la += __shfl_xor_sync(0xffffffffu, la, 1); la += __shfl_xor_sync(0xffffffffu, la, 2);
    lb += __shfl_xor_sync(0xffffffffu, lb, 1); lb += __shfl_xor_sync(0xffffffffu, lb, 2);
    float o[HD / 8][4];
#pragma unroll
    for (int n = 0; n < HD / 8; n++) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
    // ldmatrix.trans: lane i supplies the address of row (i & 7) of matrix (i >> 3): matrices 0/1 = keys 0-7 / 8-15 of
    // channel tile n, matrices 2/3 the same keys of channel tile n + 1 -> {b0, b1} of two P.V instructions
    const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lcol = (lane >> 4) * 8;
#pragma unroll
    for (int kk = 0; kk < NKT / 2; kk++) {
        const uint32_t a0 = pack_bf16(sc[2 * kk][0], sc[2 * kk][1]), a1 = pack_bf16(sc[2 * kk][2], sc[2 * kk][3]);
        const uint32_t a2 = pack_bf16(sc[2 * kk + 1][0], sc[2 * kk + 1][1]), a3 = pack_bf16(sc[2 * kk + 1][2], sc[2 * kk + 1][3]);
#pragma unroll
        for (int n = 0; n < HD / 8; n += 2) {
            uint32_t b[4];
            ldmatrix_x4_trans(b, Vw + (kk * 16 + lrow) * LD + n * 8 + lcol);
            mma_bf16(o[n], a0, a1, a2, a3, b[0], b[1]);
            mma_bf16(o[n + 1], a0, a1, a2, a3, b[2], b[3]);
        }
    }
    const float ia = qma ? 1.0f / la : 0.f, ib = qmb ? 1.0f / lb : 0.f;
    bf16 *oa = out + (base + ra) * C + c0 + 2 * t;
    bf16 *ob = out + (base + rb) * C + c0 + 2 * t;
#pragma unroll
    for (int n = 0; n < HD / 8; n++) {
        if (va) *reinterpret_cast<uint32_t *>(oa + n * 8) = pack_bf16(o[n][0] * ia, o[n][1] * ia);
        if (vb) *reinterpret_cast<uint32_t *>(ob + n * 8) = pack_bf16(o[n][2] * ib, o[n][3] * ib);
    }
}

template <int HD, int NKT>
static int launch_local_attn_mma(const bf16 *q, const bf16 *k, const bf16 *v, bf16 *out, int n_seq, int T, int C, int n_heads,
                                 int s, float scale2, const uint8_t *mask, int64_t m_seq_stride, int phase, cudaStream_t st) {
    dim3 grid(cdiv(T + phase, 16 * LM_WARPS), n_heads, n_seq);
    local_attn_mma_kernel<HD, NKT><<<grid, 32 * LM_WARPS, 0, st>>>(q, k, v, out, T, C, s, scale2, mask, m_seq_stride, phase);
    DECAF_LAUNCH_CHECK();
    return 0;
}

// The shared-memory image of one sequence's keys/values for xattn_mma_kernel<.., PACKED>: Ks [LKP][C + 8] then
// Vt [C][LKP + 8], bf16, keys >= kv_len zero.  One CTA per (sequence, 8 keys).
__global__ void __launch_bounds__(256)
xattn_pack_kv_kernel(const float *__restrict__ k, const float *__restrict__ v, const int32_t *__restrict__ kv_len,
                     bf16 *__restrict__ packed, int Lk, int C, int LKP) {
    const int seq = blockIdx.x, key0 = blockIdx.y * 8;
    const int ldk = C + 8, ldv = LKP + 8;
    bf16 *Ks = packed + (int64_t)seq * (LKP * ldk + C * ldv), *Vt = Ks + LKP * ldk;
    const int n_kv = min(kv_len[seq], Lk);
    const float *kb = k + (int64_t)seq * Lk * C, *vb = v + (int64_t)seq * Lk * C;
    const int c4n = C / 4;
    for (int i = threadIdx.x; i < 8 * c4n; i += blockDim.x) {
        const int key = key0 + i / c4n, ch = (i % c4n) * 4;
        float4 kk = make_float4(0.f, 0.f, 0.f, 0.f);
        if (key < n_kv) kk = *reinterpret_cast<const float4 *>(kb + (int64_t)key * C + ch);
        uint2 pk;
        pk.x = pack_bf16(kk.x, kk.y); pk.y = pack_bf16(kk.z, kk.w);
        *reinterpret_cast<uint2 *>(Ks + key * ldk + ch) = pk;
    }
    for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
        float x[8];
#pragma unroll
        for (int j = 0; j < 8; j++) x[j] = key0 + j < n_kv ? vb[(int64_t)(key0 + j) * C + ch] : 0.f;
        uint4 pk;
        pk.x = pack_bf16(x[0], x[1]); pk.y = pack_bf16(x[2], x[3]); pk.z = pack_bf16(x[4], x[5]); pk.w = pack_bf16(x[6], x[7]);
        *reinterpret_cast<uint4 *>(Vt + ch * ldv + key0) = pk;
    }
}

static inline size_t xattn_smem_bytes(int nkt, int C, int warps) {
    return ((size_t)8 * nkt * (C + 8) + (size_t)C * (8 * nkt + 8) + (size_t)16 * warps * (C + 8)) * sizeof(bf16);
}
// 8 warps (128 query rows) per CTA, 4 when the K/V image of a wide model leaves no room for a 128-row Q tile
static inline int xattn_warps(int nkt, int C) {
    if (xattn_smem_bytes(nkt, C, XM_WARPS) <= 200 * 1024) return XM_WARPS;
    if (xattn_smem_bytes(nkt, C, 4) <= 200 * 1024) return 4;
    return 0;
}

template <int HD, int NKT, bool PACKED, int NW>
static int launch_xattn_mma_w(const bf16 *q, const float *k, const float *v, bf16 *out, int n_seq, int Tq, int Lk, int C,
                              int n_heads, float scale2, const int32_t *kv_len, cudaStream_t st) {
    const size_t smem = xattn_smem_bytes(NKT, C, NW);
    static size_t attr = 0;
    if (smem > 48 * 1024 && smem > attr) {
        DECAF_CUDA(cudaFuncSetAttribute(xattn_mma_kernel<HD, NKT, PACKED, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = smem;
    }
    dim3 grid(cdiv(Tq, 16 * NW), n_seq);
    xattn_mma_kernel<HD, NKT, PACKED, NW><<<grid, 32 * NW, smem, st>>>(q, k, v, out, Tq, Lk, C, n_heads, scale2, kv_len);
    DECAF_LAUNCH_CHECK();
    return 0;
}

template <int HD, int NKT, bool PACKED>
static int launch_xattn_mma(const bf16 *q, const float *k, const float *v, bf16 *out, int n_seq, int Tq, int Lk, int C,
                            int n_heads, float scale2, const int32_t *kv_len, cudaStream_t st) {
    if (xattn_warps(NKT, C) == XM_WARPS)
        return launch_xattn_mma_w<HD, NKT, PACKED, XM_WARPS>(q, k, v, out, n_seq, Tq, Lk, C, n_heads, scale2, kv_len, st);
    return launch_xattn_mma_w<HD, NKT, PACKED, 4>(q, k, v, out, n_seq, Tq, Lk, C, n_heads, scale2, kv_len, st);
}

// tensor-core path: bf16 queries, head dim 32 / 64, <= 64 keys, 16-byte aligned rows
static bool xattn_mma_ok(const void *q, const void *out, int q_dtype, int n_seq, int Lk, int C, int n_heads) {
    if (q_dtype != DECAF_BF16 || C % 32 != 0 || n_heads <= 0 || C % n_heads != 0) return false;
    const int hd = C / n_heads, nkt = 2 * cdiv(Lk, 16);
    const bool al16 = ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    return (hd == 32 || hd == 64) && Lk >= 1 && Lk <= 64 && al16 && xattn_warps(nkt, C) > 0 && n_seq <= 65535;
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
    return decaf_local_attn_phase(q, k, v, out, dtype, n_seq, T, C, n_heads, window, mask, m_seq_stride, 0, stream);
}

extern "C" int decaf_local_attn_phase(const void *q, const void *k, const void *v, void *out, int32_t dtype,
                                      int32_t n_seq, int32_t T, int32_t C, int32_t n_heads, int32_t window,
                                      const uint8_t *mask, int64_t m_seq_stride, int32_t phase, void *stream) {
    DECAF_CHECK(q && k && v && out && mask, "decaf_local_attn: null pointers");
    DECAF_CHECK(phase >= 0 && phase < 16, "decaf_local_attn: phase must be in 0..15 (got %d)", phase);
    DECAF_CHECK(C % 32 == 0 && C % n_heads == 0, "decaf_local_attn: bad C/n_heads");
    DECAF_CHECK(window > 0 && window % 2 == 1, "decaf_local_attn: window must be odd and > 0");
    if (!m_seq_stride) m_seq_stride = T;
    const int64_t rows = (int64_t)n_seq * T;
    if (rows == 0) return 0;
    const int grid = cdiv(rows, AROWS);
    const float scale2 = 1.0f / sqrtf((float)(C / n_heads));
    cudaStream_t st = as_stream(stream);
    if (dtype == DECAF_BF16) {
        // tensor-core path: head dim 32 / 64, window <= 49 (16 + 2s keys per warp, padded to a multiple of 16)
        const int hd = C / n_heads, s_half = window / 2, nkt = 2 * cdiv(16 + 2 * s_half, 16);
        if ((hd == 32 || hd == 64) && nkt <= 8 && n_seq <= 65535 && n_heads <= 65535) {
#define LM(HD_, NKT_) if (hd == HD_ && nkt == NKT_) return launch_local_attn_mma<HD_, NKT_>((const bf16 *)q, (const bf16 *)k, (const bf16 *)v, (bf16 *)out, n_seq, T, C, n_heads, s_half, scale2, mask, m_seq_stride, phase, st);
            LM(64, 2) LM(64, 4) LM(64, 6) LM(64, 8) LM(32, 2) LM(32, 4) LM(32, 6) LM(32, 8)
#undef LM
        }
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
        // tensor-core path: head dim 32 / 64, <= 64 keys, K/V of a sequence staged in shared memory
        const int hd = C / n_heads, nkt = 2 * cdiv(Lk, 16);
        if (xattn_mma_ok(q, out, q_dtype, n_seq, Lk, C, n_heads) && (reinterpret_cast<uintptr_t>(k) & 15) == 0) {
#define XM(HD_, NKT_) if (hd == HD_ && nkt == NKT_) return launch_xattn_mma<HD_, NKT_, false>((const bf16 *)q, k, v, (bf16 *)out, n_seq, Tq, Lk, C, n_heads, scale2, kv_len, st);
            XM(64, 2) XM(64, 4) XM(64, 6) XM(64, 8) XM(32, 2) XM(32, 4) XM(32, 6) XM(32, 8)
#undef XM
        }
        DECAF_DISPATCH_LPH(n_heads, DECAF_DISPATCH_VEC_ATTN(C, (xattn_kernel<VEC, LPH, bf16, bf16><<<grid, 32 * AROWS, 0, st>>>(
            (const bf16 *)q, k, v, (bf16 *)out, n_seq, Tq, Lk, C, scale2, kv_len))));
    } else {
        DECAF_DISPATCH_LPH(n_heads, DECAF_DISPATCH_VEC_ATTN(C, (xattn_kernel<VEC, LPH, float, float><<<grid, 32 * AROWS, 0, st>>>(
            (const float *)q, k, v, (float *)out, n_seq, Tq, Lk, C, scale2, kv_len))));
    }
    DECAF_LAUNCH_CHECK();
    return 0;
}

extern "C" int64_t decaf_xattn_packed_elems(int32_t n_seq, int32_t Lk, int32_t C) {
    const int lkp = 16 * cdiv(Lk, 16);
    return (int64_t)n_seq * ((int64_t)lkp * (C + 8) + (int64_t)C * (lkp + 8));
}

extern "C" int decaf_xattn_packed_supported(int32_t Lk, int32_t C, int32_t n_heads) {
    return xattn_mma_ok(nullptr, nullptr, DECAF_BF16, 1, Lk, C, n_heads) ? 1 : 0;
}

extern "C" int decaf_xattn_pack_kv(const float *k, const float *v, const int32_t *kv_len, void *packed, int32_t n_seq,
                                   int32_t Lk, int32_t C, void *stream) {
    DECAF_CHECK(k && v && kv_len && packed, "decaf_xattn_pack_kv: null pointers");
    DECAF_CHECK(Lk >= 1 && Lk <= 64 && C % 32 == 0, "decaf_xattn_pack_kv: needs 1 <= Lk <= 64 and C %% 32 == 0 (Lk %d, C %d)", Lk, C);
    DECAF_CHECK(((reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(packed)) & 15) == 0, "decaf_xattn_pack_kv: unaligned buffers");
    if (n_seq == 0) return 0;
    const int lkp = 16 * cdiv(Lk, 16);
    dim3 grid(n_seq, lkp / 8);
    xattn_pack_kv_kernel<<<grid, 256, 0, as_stream(stream)>>>(k, v, kv_len, (bf16 *)packed, Lk, C, lkp);
    DECAF_LAUNCH_CHECK();
    return 0;
}

extern "C" int decaf_xattn_packed(const void *q, const void *packed, void *out, int32_t n_seq, int32_t Tq, int32_t Lk,
                                  int32_t C, int32_t n_heads, const int32_t *kv_len, void *stream) {
    DECAF_CHECK(q && packed && out && kv_len, "decaf_xattn_packed: null pointers");
    DECAF_CHECK(xattn_mma_ok(q, out, DECAF_BF16, n_seq, Lk, C, n_heads) && (reinterpret_cast<uintptr_t>(packed) & 15) == 0,
                "decaf_xattn_packed: shape not supported by the tensor-core kernel (see decaf_xattn_packed_supported)");
    if ((int64_t)n_seq * Tq == 0) return 0;
    const int hd = C / n_heads, nkt = 2 * cdiv(Lk, 16);
    const float scale2 = 1.0f / sqrtf((float)hd);
    cudaStream_t st = as_stream(stream);
#define XM(HD_, NKT_) if (hd == HD_ && nkt == NKT_) return launch_xattn_mma<HD_, NKT_, true>((const bf16 *)q, (const float *)packed, nullptr, (bf16 *)out, n_seq, Tq, Lk, C, n_heads, scale2, kv_len, st);
    XM(64, 2) XM(64, 4) XM(64, 6) XM(64, 8) XM(32, 2) XM(32, 4) XM(32, 6) XM(32, 8)
#undef XM
    decaf::set_error("decaf_xattn_packed: no instantiation for head dim %d, %d keys", hd, Lk);
    return 1;
}

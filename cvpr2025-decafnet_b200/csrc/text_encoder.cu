// Whole text encoder (+ the fusion layers' key/value projections) in ONE launch: a thread-block CLUSTER of 8
// CTAs per query.
//
// The reference runs TextTransformer once per query (libs/worker_v2.py:940-955 -> libs/modeling/text_net.py:
// 158-188): 1x1 embedding, background token, N global-attention encoder layers over <= 32 rows x 128
// channels — ~1 GFLOP per video but ~45 dependent tiny launches, i.e. pure launch latency (0.5 ms on the critical
// path of a 2.9 ms step).  Here every query owns a cluster; its activations (<= 32 rows) live in shared memory of
// the 8 CTAs as [channel][row], every GEMM is split by output channel across the 8 CTAs, and each CTA pushes its
// slice to the peers that need it through distributed shared memory (st.shared::cluster, coalesced 128-byte
// rows) followed by one cluster barrier.  DSMEM moves only ~20 bytes/cycle per SM, so slices go only where they
// are consumed: V stays local (CTA r owns channels [16 r, 16 r + 16) of q, k AND v, and produces exactly those
// attention-output channels), q / k go to the CTAs of the same head, only the residual stream, the attention output
// and the FFN hidden tensor are broadcast to all 8.
//
// What bounds a 32-row GEMM on CUDA cores is the shared-memory pipe, not the FMA pipe, so the dot product uses a
// 4 x 4 register tile per thread (one 16-byte load of 4 rows + one of 4 columns per 16 FMAs): activations
// [k][row] and weights [k][col] — the host packs every (stage, CTA) weight slice TRANSPOSED and contiguous into one
// blob, which cp.async copies straight into shared memory one stage ahead (weights do not depend on data).  Narrow
// stages split K over warps; partial tiles meet in shared memory.  All small per-layer vectors (LayerNorm affine,
// biases, LayerScale) come from a second blob, prefetched one layer ahead, so no stage waits on a global load.
//
// fp32 throughout (FMA on CUDA cores): the text path is 0.2 % of the FLOPs and feeds the softmax of the
// cross-attention, so it stays in the reference's arithmetic in both configurations.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace decaf {

constexpr int TE_CL = 8;                 // CTAs per cluster (= per query)
constexpr int TE_THREADS = 1024;
constexpr int TE_WARPS = TE_THREADS / 32;
constexpr int TE_ROWS = 32;              // rows per query (L1 = Lmax + 1 <= 32)
constexpr int TE_MAXC = 128;             // text embedding width
constexpr int TE_XLD = 33;               // row pitch of Xt: conflict-free for "lane = row" AND "lane = channel"
constexpr int TE_PLD = 33;               // row pitch of the partial-sum buffer (same reason)
constexpr int TE_KCHUNK = 128;           // token-feature channels per embedding chunk
constexpr int TE_TLD = TE_KCHUNK + 1;    // row pitch of a token chunk [row][k]
constexpr int TE_WBUF = 32 * 1024;       // one weight-slice buffer (two of them)
constexpr int TE_PART = 256;             // (k part, column) entries of the partial-sum buffer
constexpr int TE_MAXPB = 15 * TE_MAXC;   // floats of one parameter block

// shared-memory plan (bytes)
constexpr int TE_OFF_X = 0;                                        // Xt  [Ct][33]  residual stream
constexpr int TE_OFF_T = TE_OFF_X + TE_MAXC * TE_XLD * 4;          // Tt  [Ct][32]  LN output / attention output
constexpr int TE_OFF_BIG = TE_OFF_T + TE_MAXC * TE_ROWS * 4;       // Qt, Kt, Vt [Ct][32] | Ht [4Ct][32] | token + embd-weight chunks
constexpr int TE_BIG_BYTES = 4 * TE_MAXC * TE_ROWS * 4;            // 64 KB
constexpr int TE_OFF_W = TE_OFF_BIG + TE_BIG_BYTES;
constexpr int TE_OFF_PART = TE_OFF_W + 2 * TE_WBUF;
constexpr int TE_OFF_PB = TE_OFF_PART + TE_PART * TE_PLD * 4;      // two parameter blocks
constexpr int TE_OFF_S = TE_OFF_PB + 2 * TE_MAXPB * 4;             // attention scores [key][33] + LN statistics [2][32]
constexpr int TE_SMEM = TE_OFF_S + (TE_ROWS * TE_XLD + 2 * TE_ROWS) * 4;
constexpr int TE_CHUNK_FLOATS = TE_ROWS * TE_TLD + TE_KCHUNK * (TE_MAXC / TE_CL);   // one token chunk + its weight chunk
static_assert(2 * TE_CHUNK_FLOATS * 4 <= TE_BIG_BYTES, "embedding chunks must fit the BIG region");

// ---- blob layouts (floats); see decaf_text_encoder_t in include/decaf_b200.h
struct TeDims { int Ct, Ctok, H, cpc, L, F, C; };
__device__ __forceinline__ int64_t te_w_embd(const TeDims &d, int rank) { return (int64_t)rank * d.Ctok * d.cpc; }
__device__ __forceinline__ int64_t te_w_layer(const TeDims &d, int l) { return (int64_t)d.Ct * d.Ctok + (int64_t)l * 12 * d.Ct * d.Ct; }
__device__ __forceinline__ int64_t te_w_fusion(const TeDims &d, int f) {
    return te_w_layer(d, d.L) + (int64_t)f * 2 * d.C * d.Ct;
}
__device__ __forceinline__ int te_p_block(const TeDims &d, int b) {          // block 0 embd, 1..L layers, L+1.. fusion
    if (b == 0) return 0;
    if (b <= d.L) return 2 * d.Ct + (b - 1) * 15 * d.Ct;
    return 2 * d.Ct + d.L * 15 * d.Ct + (b - 1 - d.L) * (2 * d.Ct + 2 * d.C);
}
__device__ __forceinline__ int te_p_size(const TeDims &d, int b) {
    return b == 0 ? 2 * d.Ct : (b <= d.L ? 15 * d.Ct : 2 * d.Ct + 2 * d.C);
}

// cp.async copy of n floats (n % 4 == 0, 16-byte aligned source and destination)
__device__ __forceinline__ void te_copy(float *dst, const float *src, int n) {
    for (int i = threadIdx.x; i < n / 4; i += TE_THREADS) {
        const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst + i * 4);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + i * 4) : "memory");
    }
}
__device__ __forceinline__ void te_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void te_wait_all() {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
}

// Weight slice [K][ns] of GEMM stage s of CTA `rank`: 1 + 4 l + {0,1,2,3} = qkv / proj / fc / proj2 of layer l,
// 1 + 4 L + f = fusion layer f's key/value projection (stage 0, the embedding, streams its weights in chunks).
__device__ __noinline__ void te_prefetch_stage(const float *__restrict__ wblob, TeDims d, int s, int rank, float *wbuf) {
    const int n_stage = 1 + 4 * d.L + d.F;
    if (s >= 1 && s < n_stage) {
        const int64_t CC = (int64_t)d.Ct * d.Ct;
        if (s <= 4 * d.L) {
            const int l = (s - 1) >> 2, k = (s - 1) & 3;
            const float *base = wblob + te_w_layer(d, l);
            if (k == 0) te_copy(wbuf, base + (int64_t)rank * d.Ct * 3 * d.cpc, d.Ct * 3 * d.cpc);
            else if (k == 1) te_copy(wbuf, base + 3 * CC + (int64_t)rank * d.Ct * d.cpc, d.Ct * d.cpc);
            else if (k == 2) te_copy(wbuf, base + 4 * CC + (int64_t)rank * d.Ct * (d.H / TE_CL), d.Ct * (d.H / TE_CL));
            else te_copy(wbuf, base + 8 * CC + (int64_t)rank * d.H * d.cpc, d.H * d.cpc);
        } else {
            const int ns = 2 * d.C / TE_CL;
            te_copy(wbuf, wblob + te_w_fusion(d, s - 1 - 4 * d.L) + (int64_t)rank * d.Ct * ns, d.Ct * ns);
        }
    }
    te_commit();
}

// How a stage's ns output columns x K are spread over the 32 warps: every warp owns a 32-row x 16-column tile of one
// K part; narrow stages split K into `kparts` parts (power of two, kparts * ns <= TE_PART) added by te_part_sum.
struct TeSplit { int ctiles, kparts, klen; };
__device__ __forceinline__ TeSplit te_split(int ns, int K) {
    TeSplit s;
    s.ctiles = (ns + 15) / 16;
    int kp = 1;
    while (2 * kp * s.ctiles <= TE_WARPS && 2 * kp * ns <= TE_PART && K % (2 * kp) == 0) kp *= 2;
    s.kparts = kp;
    s.klen = K / kp;
    return s;
}

// part[(kpart * ns + col) * 33 + row] (+)= sum over the warp's K part of A[k][row] * Wt[k][col].
// Wt: [K][ns] in shared memory.  XT: activations are [k][32 rows] (one 16-byte load gives the thread's 4 rows);
// otherwise element (k, row) sits at At[row * sl + k] (token chunks).  4 x 4 register tile per thread: lane =
// (row group = lane / 4, column group = lane % 4).  Not inlined: the kernel body runs once per query, so its size is
// what the instruction cache sees.
template <bool XT>
__device__ __noinline__ void te_dot(const float *At, int sl, const float *Wt, int K, int ns, float *part, int accumulate) {
    __builtin_assume(__isShared(At));                  // tell the compiler these are LDS / STS, not generic accesses
    __builtin_assume(__isShared(Wt));
    __builtin_assume(__isShared(part));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const TeSplit sp = te_split(ns, K);
    const int kpart = warp / sp.ctiles, ct = warp % sp.ctiles;
    if (kpart >= sp.kparts) return;
    const int r0 = 4 * (lane >> 2);
    int c0 = ct * 16 + 4 * (lane & 3);
    const bool col_ok = c0 < ns;                        // ns % 4 == 0: a thread's 4 columns are all in or all out
    if (!col_ok) c0 = 0;
    const int k0 = kpart * sp.klen;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
    const float *wp = Wt + (int64_t)k0 * ns + c0;
    const float *ap = XT ? At + (int64_t)k0 * TE_ROWS + r0 : At + (int64_t)r0 * sl + k0;
#pragma unroll 4
    for (int k = 0; k < sp.klen; k++) {
        float x[4];
        if (XT) {
            const float4 x4 = *reinterpret_cast<const float4 *>(ap + k * TE_ROWS);
            x[0] = x4.x; x[1] = x4.y; x[2] = x4.z; x[3] = x4.w;
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) x[i] = ap[i * sl + k];
        }
        const float4 w4 = *reinterpret_cast<const float4 *>(wp + (int64_t)k * ns);
        const float w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j] = fmaf(x[i], w[j], acc[i][j]);
    }
    if (!col_ok) return;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        float *dst = part + (kpart * ns + c0 + j) * TE_PLD + r0;
#pragma unroll
        for (int i = 0; i < 4; i++) dst[i] = accumulate ? dst[i] + acc[i][j] : acc[i][j];
    }
}
__device__ __forceinline__ float te_part_sum(const float *part, int ns, int K, int col, int row) {
    const TeSplit sp = te_split(ns, K);
    float v = 0.f;
    for (int kp = 0; kp < sp.kparts; kp++) v += part[(kp * ns + col) * TE_PLD + row];
    return v;
}

// Channel LayerNorm of every row of Xt (two-pass, biased variance, eps inside the sqrt; libs/modeling/blocks.py:
// 125-131) -> Tt.  Computed redundantly by every CTA of the cluster.  Statistics: warp = row, lanes over channels
// (Xt's pitch of 33 makes that conflict-free); normalisation: lane = row, warps over channels.
__device__ __noinline__ void te_layernorm(const float *Xt, float *Tt, float *stat, int Ct, const float *w, const float *b, float eps) {
    __builtin_assume(__isShared(Xt));
    __builtin_assume(__isShared(Tt));
    __builtin_assume(__isShared(stat));
    __builtin_assume(__isShared(w));
    __builtin_assume(__isShared(b));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    {
        float x[TE_MAXC / 32], s = 0.f;
#pragma unroll
        for (int i = 0; i < TE_MAXC / 32; i++) {
            const int c = lane + 32 * i;
            x[i] = c < Ct ? Xt[c * TE_XLD + warp] : 0.f;
            s += x[i];
        }
        const float mean = warp_sum(s) / (float)Ct;
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < TE_MAXC / 32; i++) {
            const float dlt = lane + 32 * i < Ct ? x[i] - mean : 0.f;
            ss = fmaf(dlt, dlt, ss);
        }
        const float var = warp_sum(ss) / (float)Ct;
        if (lane == 0) { stat[warp] = mean; stat[TE_ROWS + warp] = rsqrtf(var + eps); }
    }
    __syncthreads();
    const float mean = stat[lane], r = stat[TE_ROWS + lane];
    for (int c = warp; c < Ct; c += TE_WARPS) Tt[c * TE_ROWS + lane] = (Xt[c * TE_XLD + lane] - mean) * r * w[c] + b[c];
    __syncthreads();
}

// store v at the same shared-memory offset in CTAs [r0, r0 + n) of the cluster
__device__ __forceinline__ void te_push(cg::cluster_group &cl, float *local_ptr, float v, int r0, int n) {
    for (int r = r0; r < r0 + n; r++) *cl.map_shared_rank(local_ptr, r) = v;
}

__global__ void __cluster_dims__(TE_CL, 1, 1) __launch_bounds__(TE_THREADS, 1)
text_encoder_kernel(const decaf_text_encoder_t p, unsigned long long *trace) {
    extern __shared__ __align__(16) uint8_t te_smem[];
    cg::cluster_group cl = cg::this_cluster();
    const int rank = (int)cl.block_rank();
    const int q = blockIdx.x / TE_CL;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    TeDims d;
    d.Ct = p.Ct; d.Ctok = p.Ctok; d.H = 4 * p.Ct; d.cpc = p.Ct / TE_CL; d.L = p.n_layers; d.F = p.n_fusion; d.C = p.C;
    const int Ct = d.Ct, L1 = p.Lmax + 1, H = d.H, cpc = d.cpc;
    const int col0 = rank * cpc;                                    // this CTA's channel slice of every Ct-wide tensor
    float *Xt = reinterpret_cast<float *>(te_smem + TE_OFF_X);
    float *Tt = reinterpret_cast<float *>(te_smem + TE_OFF_T);
    float *BIG = reinterpret_cast<float *>(te_smem + TE_OFF_BIG);
    float *Qt = BIG, *Kt = BIG + Ct * TE_ROWS, *Vt = BIG + 2 * Ct * TE_ROWS;    // all [channel][32 rows]
    float *Ht = BIG;
    float *wbuf[2] = {reinterpret_cast<float *>(te_smem + TE_OFF_W), reinterpret_cast<float *>(te_smem + TE_OFF_W + TE_WBUF)};
    float *part = reinterpret_cast<float *>(te_smem + TE_OFF_PART);
    float *pbuf[2] = {reinterpret_cast<float *>(te_smem + TE_OFF_PB), reinterpret_cast<float *>(te_smem + TE_OFF_PB) + TE_MAXPB};
    float *S = reinterpret_cast<float *>(te_smem + TE_OFF_S);      // [key][33]
    float *stat = S + TE_ROWS * TE_XLD;
    const int len = min(p.lens[q], p.Lmax);
    const int kv_len = len + 1;
    const float rowmask = lane < kv_len ? 1.f : 0.f;               // text mask of row `lane` (bkgd token + len words)
    if (rank == 0 && threadIdx.x == 0 && p.kv_len_out) p.kv_len_out[q] = kv_len;
    int trn = 0;
#define TE_STAMP() do { if (trace && blockIdx.x == 0 && threadIdx.x == 0 && trn < 256) trace[trn++] = clock64(); } while (0)
    TE_STAMP();

    int wb = 0, stage = 0, pb = 0;                                  // wbuf[wb] / pbuf[pb & 1]: the CURRENT stage's slice / block
    auto prefetch_params = [&](int b) {
        if (b <= d.L + d.F) te_copy(pbuf[b & 1], p.pblob + te_p_block(d, b), te_p_size(d, b));
    };
    // ---------------------------------------------------------------- stage 0: embedding (text_net.py:163-183)
    {
        // chunks of 128 token channels: tokens [32 rows][129] (4-byte cp.async, nothing staged through registers) and
        // the matching [128][cpc] rows of this CTA's transposed weight slice; two chunks in flight
        float *cbuf[2] = {BIG, BIG + TE_CHUNK_FLOATS};
        const float *tok = p.tokens + (int64_t)q * p.Lmax * p.Ctok;
        const float *wsl = p.wblob + te_w_embd(d, rank);
        const int n_chunk = (p.Ctok + TE_KCHUNK - 1) / TE_KCHUNK;
        auto stage_chunk = [&](int ci) {
            if (ci < n_chunk) {
                const int k0 = ci * TE_KCHUNK, kc = min(TE_KCHUNK, p.Ctok - k0);
                float *dst = cbuf[ci & 1];
                for (int i = threadIdx.x; i < p.Lmax * kc; i += TE_THREADS) {
                    const int l = i / kc, k = i - l * kc;
                    const uint32_t a = (uint32_t)__cvta_generic_to_shared(dst + l * TE_TLD + k);
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(a), "l"(tok + (int64_t)l * p.Ctok + k0 + k) : "memory");
                }
                te_copy(dst + TE_ROWS * TE_TLD, wsl + (int64_t)k0 * cpc, kc * cpc);
            }
            te_commit();
        };
        // rows >= Lmax of the token chunks are never copied: zero them once (their products are discarded anyway)
        for (int i = threadIdx.x; i < 2 * TE_CHUNK_FLOATS; i += TE_THREADS) BIG[i] = 0.f;
        __syncthreads();
        prefetch_params(0);
        prefetch_params(1);
        te_prefetch_stage(p.wblob, d, 1, rank, wbuf[0]);            // (one group: both parameter blocks + layer 0's qkv slice)
        stage_chunk(0);
        stage_chunk(1);
        for (int ci = 0; ci < n_chunk; ci++) {
            const int kc = min(TE_KCHUNK, p.Ctok - ci * TE_KCHUNK);
            asm volatile("cp.async.wait_group 1;" ::: "memory");       // everything up to chunk ci has landed
            __syncthreads();
            te_dot<false>(cbuf[ci & 1], TE_TLD, cbuf[ci & 1] + TE_ROWS * TE_TLD, kc, cpc, part, ci > 0);
            __syncthreads();
            stage_chunk(ci + 2);                                       // (an empty group past the last chunk)
        }
        te_wait_all();
        stage = 1;
        // row l + 1 <- word l; row 0 = background token (text_net.py:175-178).  lane = row.
        const float *pe0 = pbuf[0];                                   // [embd_b | bkgd]
        const int kc0 = min(TE_KCHUNK, p.Ctok);                       // every chunk is split like the first one
        for (int cs = warp; cs < cpc; cs += TE_WARPS) {
            const int n = col0 + cs;
            float v;
            if (lane == 0) v = pe0[Ct + n];
            else {
                v = 0.f;
                if (lane - 1 < len) {
                    v = te_part_sum(part, cpc, kc0, cs, lane - 1) + pe0[n];
                    if (p.pe) v += text_pe_value(p.pe, p.pe_rows, Ct, len, lane - 1, n);
                }
            }
            te_push(cl, Xt + n * TE_XLD + lane, v, 0, TE_CL);
        }
        pb = 1;
        cl.sync();
        TE_STAMP();
    }
    // ---------------------------------------------------------------- encoder layers (blocks.py:578-591, stride 0)
    const int hd = Ct / p.n_heads;
    const float scale = rsqrtf((float)hd);                          // (d^-1/4)^2: q and k are both scaled (blocks.py:379)
    const int gsz = hd / cpc, g0 = (rank / gsz) * gsz;              // the CTAs holding the channels of my head
    for (int layer = 0; layer < d.L; layer++) {
        // parameter block of this layer: [ln_attn w,b | q_b k_b v_b | proj_b | ls_attn | ln_ffn w,b | fc_b (4Ct) | proj2_b | ls_ffn]
        const float *P = pbuf[pb & 1];
        const float *qkv_b = P + 2 * Ct, *proj_b = P + 5 * Ct, *ls_attn = P + 6 * Ct, *ln2 = P + 7 * Ct;
        const float *fc_b = P + 9 * Ct, *proj2_b = P + 13 * Ct, *ls_ffn = P + 14 * Ct;
        // --- LN -> q, k, v of this CTA's channels (global attention, no depthwise convs in the text encoder)
        {
            te_wait_all();                                          // this stage's weights (+ this layer's parameters)
            prefetch_params(pb + 1);
            te_prefetch_stage(p.wblob, d, ++stage, rank, wbuf[wb ^ 1]);
            te_layernorm(Xt, Tt, stat, Ct, P, P + Ct, p.eps);
            te_dot<true>(Tt, 0, wbuf[wb], Ct, 3 * cpc, part, 0);
            __syncthreads();
            for (int cs = warp; cs < 3 * cpc; cs += TE_WARPS) {
                const int j = cs / cpc, n = col0 + cs % cpc;        // j: 0 q, 1 k, 2 v
                const float v = te_part_sum(part, 3 * cpc, Ct, cs, lane) + qkv_b[j * Ct + n];
                float *dst = BIG + (j * Ct + n) * TE_ROWS + lane;
                if (j < 2) te_push(cl, dst, v, g0, gsz);            // q, k: to the CTAs of this head
                else *dst = v;                                      // v: consumed here only
            }
            wb ^= 1;
            cl.sync();
            TE_STAMP();
        }
        // --- attention (blocks.py:374-389: -inf on masked keys) for this CTA's head, output channels [col0, col0 + cpc):
        // scores once per (row, key), softmax with warp = row, then P V with lane = row
        {
            const int h = col0 / hd;
            if (warp < TE_ROWS / 4) {                               // warp = group of 4 keys, lane = query row
                const float *qh = Qt + h * hd * TE_ROWS + lane, *kh = Kt + h * hd * TE_ROWS + 4 * warp;
                float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
                for (int dd = 0; dd < hd; dd++) {
                    const float qv = qh[dd * TE_ROWS];
                    const float4 k4 = *reinterpret_cast<const float4 *>(kh + dd * TE_ROWS);     // broadcast
                    s[0] = fmaf(qv, k4.x, s[0]); s[1] = fmaf(qv, k4.y, s[1]); s[2] = fmaf(qv, k4.z, s[2]); s[3] = fmaf(qv, k4.w, s[3]);
                }
#pragma unroll
                for (int j = 0; j < 4; j++) S[(4 * warp + j) * TE_XLD + lane] = 4 * warp + j < kv_len ? s[j] * scale : -INFINITY;
            }
            __syncthreads();
            {                                                       // warp = query row, lane = key
                const float s = S[lane * TE_XLD + warp];
                const float m = warp_max(s);                        // key 0 (the bkgd token) is always valid: m is finite
                const float e = __expf(s - m);
                S[lane * TE_XLD + warp] = e / warp_sum(e);
            }
            __syncthreads();
            const int half = warp / cpc, cs = warp % cpc;           // 2 warps per channel, each over half of the keys
            if (half < 2) {
                const int j0 = half * (TE_ROWS / 2);
                const float *vr = Vt + (col0 + cs) * TE_ROWS + j0;
                float o = 0.f;
#pragma unroll
                for (int j4 = 0; j4 < TE_ROWS / 8; j4++) {
                    const float4 v4 = *reinterpret_cast<const float4 *>(vr + 4 * j4);           // broadcast
                    const float *sp = S + (j0 + 4 * j4) * TE_XLD + lane;
                    o = fmaf(sp[0], v4.x, o); o = fmaf(sp[TE_XLD], v4.y, o);
                    o = fmaf(sp[2 * TE_XLD], v4.z, o); o = fmaf(sp[3 * TE_XLD], v4.w, o);
                }
                part[(half * cpc + cs) * TE_PLD + lane] = o;
            }
            __syncthreads();
            for (int c = warp; c < cpc; c += TE_WARPS)
                te_push(cl, Tt + (col0 + c) * TE_ROWS + lane, part[c * TE_PLD + lane] + part[(cpc + c) * TE_PLD + lane], 0, TE_CL);
            cl.sync();
            TE_STAMP();
        }
        // --- proj + LayerScale + residual + mask (blocks.py:586)
        {
            te_wait_all();
            te_prefetch_stage(p.wblob, d, ++stage, rank, wbuf[wb ^ 1]);
            te_dot<true>(Tt, 0, wbuf[wb], Ct, cpc, part, 0);
            __syncthreads();
            for (int cs = warp; cs < cpc; cs += TE_WARPS) {
                const int n = col0 + cs;
                const float a = te_part_sum(part, cpc, Ct, cs, lane);
                const float v = (Xt[n * TE_XLD + lane] + ls_attn[n] * (a + proj_b[n])) * rowmask;
                te_push(cl, Xt + n * TE_XLD + lane, v, 0, TE_CL);
            }
            wb ^= 1;
            cl.sync();
            TE_STAMP();
        }
        // --- LN -> fc -> GELU (blocks.py:535-538)
        {
            const int ns = H / TE_CL;
            te_wait_all();
            te_prefetch_stage(p.wblob, d, ++stage, rank, wbuf[wb ^ 1]);
            te_layernorm(Xt, Tt, stat, Ct, ln2, ln2 + Ct, p.eps);
            te_dot<true>(Tt, 0, wbuf[wb], Ct, ns, part, 0);
            __syncthreads();
            for (int cs = warp; cs < ns; cs += TE_WARPS) {
                const int n = rank * ns + cs;
                te_push(cl, Ht + n * TE_ROWS + lane, gelu_erf(te_part_sum(part, ns, Ct, cs, lane) + fc_b[n]), 0, TE_CL);
            }
            wb ^= 1;
            cl.sync();
            TE_STAMP();
        }
        // --- proj2 + LayerScale + residual + mask (blocks.py:589-590)
        {
            te_wait_all();
            te_prefetch_stage(p.wblob, d, ++stage, rank, wbuf[wb ^ 1]);
            te_dot<true>(Ht, 0, wbuf[wb], H, cpc, part, 0);
            __syncthreads();
            for (int cs = warp; cs < cpc; cs += TE_WARPS) {
                const int n = col0 + cs;
                const float a = te_part_sum(part, cpc, H, cs, lane);
                const float v = (Xt[n * TE_XLD + lane] + ls_ffn[n] * (a + proj2_b[n])) * rowmask;
                te_push(cl, Xt + n * TE_XLD + lane, v, 0, TE_CL);
            }
            wb ^= 1;
            pb++;
            cl.sync();
            TE_STAMP();
        }
    }
    // ---------------------------------------------------------------- outputs
    // text (n, L1, Ct) fp32, row-major (each CTA writes its channel slice)
    if (p.text_out) {
        for (int i = threadIdx.x; i < L1 * cpc; i += TE_THREADS) {
            const int r = i / cpc, n = col0 + i % cpc;
            p.text_out[((int64_t)q * L1 + r) * Ct + n] = Xt[n * TE_XLD + r];
        }
    }
    // fusion key/value projections: kv_out[f][0|1][q * L1 + row][C] = LN(text; lnkv_f) W_{k|v}^T + b
    // (TransformerDecoder.ln_xattn_kv + MaskedMHA key/value, blocks.py:640-641, 348-350); warp = row, lane = column:
    // coalesced global stores
    for (int f = 0; f < d.F; f++) {
        const int C = d.C, ns = 2 * C / TE_CL, n0 = rank * ns;
        const float *P = pbuf[pb & 1];                              // [lnkv_w | lnkv_b | k_b v_b (2C)]
        te_wait_all();
        prefetch_params(pb + 1);
        te_prefetch_stage(p.wblob, d, ++stage, rank, wbuf[wb ^ 1]);
        te_layernorm(Xt, Tt, stat, Ct, P, P + Ct, p.eps);
        te_dot<true>(Tt, 0, wbuf[wb], Ct, ns, part, 0);
        __syncthreads();
        for (int r = warp; r < L1; r += TE_WARPS) {
            for (int c = lane; c < ns; c += 32) {
                const int n = n0 + c;                               // 0..2C: [k | v]
                float *dst = p.kv_out + ((int64_t)f * 2 + n / C) * p.n_query * L1 * C;
                dst[((int64_t)q * L1 + r) * C + n % C] = te_part_sum(part, ns, Ct, c, r) + P[2 * Ct + n];
            }
        }
        __syncthreads();
        wb ^= 1;
        pb++;
    }
    TE_STAMP();
    cl.sync();                                                     // no CTA exits while peers may still write its smem
}

}  // namespace decaf

using namespace decaf;

static unsigned long long *g_te_trace = nullptr;
// debug only: clock64 stamps of CTA 0 / thread 0 after every stage of the following launches (NULL = off)
extern "C" int decaf_debug_text_trace(unsigned long long *buf) {
    g_te_trace = buf;
    return 0;
}

// debug only: how many clusters of the text-encoder kernel can be resident at once on the current device
extern "C" int decaf_debug_text_max_clusters(void) {
    cudaFuncSetAttribute(text_encoder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TE_SMEM);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(16 * TE_CL); cfg.blockDim = dim3(TE_THREADS); cfg.dynamicSmemBytes = TE_SMEM;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = TE_CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = -1;
    if (cudaOccupancyMaxActiveClusters(&n, text_encoder_kernel, &cfg) != cudaSuccess) return -1;
    return n;
}

extern "C" int decaf_text_encoder_supported(int32_t Lmax, int32_t Ct, int32_t Ctok, int32_t n_heads, int32_t n_layers,
                                            int32_t C, int32_t n_fusion) {
    if (Lmax + 1 > TE_ROWS || Lmax < 1) return 0;
    if (Ct > TE_MAXC || Ct % (4 * TE_CL) != 0 || n_heads < 1 || Ct % n_heads != 0) return 0;
    const int cpc = Ct / TE_CL, hd = Ct / n_heads, H = 4 * Ct;
    if (hd % cpc != 0) return 0;                                   // a CTA's channel slice must lie inside one head
    if (Ctok % 4 != 0 || (Ctok > TE_KCHUNK && Ctok % TE_KCHUNK != 0)) return 0;   // equal embedding chunks
    if (n_layers < 1 || n_fusion < 0) return 0;
    // every weight slice must fit one 32 KB buffer, every stage's partial sums the partial-sum buffer
    if ((int64_t)3 * cpc * Ct * 4 > TE_WBUF || (int64_t)(H / TE_CL) * Ct * 4 > TE_WBUF || (int64_t)cpc * H * 4 > TE_WBUF) return 0;
    if (H / TE_CL > TE_PART || 3 * cpc > TE_PART) return 0;
    if (n_fusion > 0 && (C % (4 * TE_CL) != 0 || (int64_t)(2 * C / TE_CL) * Ct * 4 > TE_WBUF || 2 * C / TE_CL > TE_PART ||
                         2 * Ct + 2 * C > TE_MAXPB)) return 0;
    return 1;
}

extern "C" int64_t decaf_text_encoder_wblob_floats(int32_t Ct, int32_t Ctok, int32_t n_layers, int32_t C, int32_t n_fusion) {
    return (int64_t)Ct * Ctok + (int64_t)n_layers * 12 * Ct * Ct + (int64_t)n_fusion * 2 * C * Ct;
}
extern "C" int64_t decaf_text_encoder_pblob_floats(int32_t Ct, int32_t n_layers, int32_t C, int32_t n_fusion) {
    return 2 * (int64_t)Ct + (int64_t)n_layers * 15 * Ct + (int64_t)n_fusion * (2 * Ct + 2 * C);
}

extern "C" int decaf_text_encoder(const decaf_text_encoder_t *pp, void *stream) {
    DECAF_CHECK(pp && pp->tokens && pp->lens && pp->wblob && pp->pblob, "decaf_text_encoder: null pointers");
    DECAF_CHECK(decaf_text_encoder_supported(pp->Lmax, pp->Ct, pp->Ctok, pp->n_heads, pp->n_layers, pp->C, pp->n_fusion),
                "decaf_text_encoder: unsupported shape (Lmax %d Ct %d Ctok %d heads %d layers %d C %d fusion %d)", pp->Lmax,
                pp->Ct, pp->Ctok, pp->n_heads, pp->n_layers, pp->C, pp->n_fusion);
    DECAF_CHECK(pp->n_fusion == 0 || pp->kv_out, "decaf_text_encoder: kv_out is required when n_fusion > 0");
    DECAF_CHECK(((reinterpret_cast<uintptr_t>(pp->wblob) | reinterpret_cast<uintptr_t>(pp->pblob)) & 15) == 0,
                "decaf_text_encoder: blobs must be 16-byte aligned");
    if (pp->n_query == 0) return 0;
    static bool attr_set = false;
    if (!attr_set) {
        DECAF_CUDA(cudaFuncSetAttribute(text_encoder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TE_SMEM));
        attr_set = true;
    }
    text_encoder_kernel<<<pp->n_query * TE_CL, TE_THREADS, TE_SMEM, as_stream(stream)>>>(*pp, g_te_trace);
    DECAF_LAUNCH_CHECK();
    return 0;
}

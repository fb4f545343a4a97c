// Segment decoding (exact top-k) and 1D (soft-)NMS + voting, one CTA per query.
// reference: Evaluator._collect_segments (libs/worker_v2.py:1131-1187), libs/nms/nms.py,
// libs/nms/src/nms_cpu.cpp.  Integer/index results (candidate order, keep-sets) are exact;
// see SURVEY.md A.5/A.6 for the tie rules reproduced here.
#include "common.cuh"

namespace decaf {

constexpr int DEC_THREADS = 1024;
constexpr int NMS_THREADS = 1024;

// order-preserving float -> uint32 (larger float <=> larger key)
__device__ __forceinline__ uint32_t f2key(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k);
}

// exclusive prefix sum of one int per thread over the CTA; `total` = CTA sum.
// scratch: >= 33 ints of shared memory.  Ends with a barrier (scratch reusable).
__device__ __forceinline__ int block_excl_scan(int v, int *scratch, int &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) scratch[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < nwarp ? scratch[lane] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += n;
        }
        scratch[lane] = winc - w;                  // exclusive warp offsets
        if (lane == 31) scratch[32] = winc;
    }
    __syncthreads();
    const int res = inc - v + scratch[warp];
    total = scratch[32];
    __syncthreads();
    return res;
}

// in-place ascending bitonic sort of n_pow2 uint64 keys in shared memory
__device__ __forceinline__ void block_bitonic_sort(unsigned long long *a, int n_pow2) {
    for (int k = 2; k <= n_pow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long x = a[i], y = a[ixj];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { a[i] = y; a[ixj] = x; }
                }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ int level_of(const decaf_levels_t &lv, int r) {
    for (int l = 0; l < lv.n_levels; l++)
        if (r >= lv.off[l] && r < lv.off[l] + lv.len[l]) return l;
    return -1;
}

// ------------------------------------------------------------------------------- decode
__global__ void __launch_bounds__(DEC_THREADS)
decode_kernel(const float *__restrict__ logits, const float *__restrict__ offsets, const uint8_t *__restrict__ hmask,
              decaf_levels_t lv, int from_logits, float thresh, int topk, int ns_pow2, float seg_len_thresh,
              float *__restrict__ cand_segs, float *__restrict__ cand_scores, int32_t *__restrict__ cand_idx,
              int32_t *__restrict__ cand_count, decaf_decode_window_t win) {
    extern __shared__ unsigned long long sel_list[];          // [ns_pow2]
    __shared__ int hist[256];
    __shared__ int scratch[34];
    __shared__ uint32_t s_prefix, s_pmask;
    __shared__ int s_remaining;
    const int q = blockIdx.x, tid = threadIdx.x;
    const int Pp = lv.Pp;
    const int64_t base = (int64_t)q * Pp;

    auto row_key = [&](int r, uint32_t &key) -> bool {
        const int l = level_of(lv, r);
        if (l < 0) return false;
        if (win.own_hi > 0) {                                 // time shard: only the points this shard owns
            const int t = r - lv.off[l];
            if (t < (win.own_lo >> l) || t >= (win.own_hi >> l)) return false;
        }
        float s = logits[base + r];
        if (from_logits) s = 1.0f / (1.0f + expf(-s));       // torch.sigmoid
        s *= (float)hmask[base + r];                          // scores *= masks.float()
        key = f2key(s);
        return s > thresh;
    };

    // 1) count candidates
    int cnt = 0;
    for (int r = tid; r < Pp; r += DEC_THREADS) { uint32_t k; cnt += row_key(r, k) ? 1 : 0; }
    int total;
    block_excl_scan(cnt, scratch, total);
    const int ksel = min(topk, total);
    if (ksel == 0) {
        if (tid == 0) cand_count[q] = 0;
        return;
    }
    // 2) radix select: key of the ksel-th largest candidate
    if (tid == 0) { s_prefix = 0; s_pmask = 0; s_remaining = ksel; }
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = tid; i < 256; i += DEC_THREADS) hist[i] = 0;
        __syncthreads();
        const uint32_t prefix = s_prefix, pmask = s_pmask;
        for (int r = tid; r < Pp; r += DEC_THREADS) {
            uint32_t k;
            if (row_key(r, k) && (k & pmask) == prefix) atomicAdd(&hist[(k >> shift) & 255], 1);
        }
        __syncthreads();
        if (tid == 0) {
            int rem = s_remaining, b = 255;
            for (; b > 0; b--) {
                if (hist[b] >= rem) break;
                rem -= hist[b];
            }
            s_remaining = rem;
            s_prefix = prefix | ((uint32_t)b << shift);
            s_pmask = pmask | (255u << shift);
        }
        __syncthreads();
    }
    const uint32_t kth = s_prefix;
    const int need_eq = s_remaining;                          // # of candidates == kth to take (>= 1)
    // 3) ordered compaction (ascending row index): key > kth, or the first need_eq with key == kth
    int eq_base = 0, out_base = 0;
    for (int r0 = 0; r0 < Pp; r0 += DEC_THREADS) {
        const int r = r0 + tid;
        uint32_t k = 0;
        const bool cand = r < Pp && row_key(r, k);
        const int feq = cand && k == kth;
        int teq;
        const int eq_rank = eq_base + block_excl_scan(feq, scratch, teq);
        const int take = cand && (k > kth || (feq && eq_rank < need_eq));
        int ttake;
        const int pos = out_base + block_excl_scan(take, scratch, ttake);
        if (take) sel_list[pos] = ((unsigned long long)(~k) << 32) | (uint32_t)r;
        eq_base += teq;
        out_base += ttake;
    }
    for (int i = ksel + tid; i < ns_pow2; i += DEC_THREADS) sel_list[i] = ~0ull;
    __syncthreads();
    // 4) sort: descending score, ties by ascending row (level-major flat order)
    block_bitonic_sort(sel_list, ns_pow2);
    // 5) decode + length filter, order preserving
    int wbase = 0;
    for (int i0 = 0; i0 < ksel; i0 += DEC_THREADS) {
        const int i = i0 + tid;
        float left = 0.f, right = 0.f, score = 0.f;
        int flat = 0, keep = 0;
        if (i < ksel) {
            const unsigned long long e = sel_list[i];
            const int r = (int)(e & 0xffffffffu);
            score = key2f(~(uint32_t)(e >> 32));
            const int l = level_of(lv, r);
            const int t = r - lv.off[l];
            const float stride = (float)(1 << l);
            // PtGenerator: tics[::stride]; a time shard adds its window origin (exact: integers < 2^24)
            const float ctr = (float)t * stride + (float)win.t0;
            const float o0 = offsets[(base + r) * 2], o1 = offsets[(base + r) * 2 + 1];
            left = __fsub_rn(ctr, __fmul_rn(o0, stride));
            right = __fadd_rn(ctr, __fmul_rn(o1, stride));
            keep = __fsub_rn(right, left) > seg_len_thresh;
            flat = r - (l + 1);                               // drop the pad rows before level l
            if (win.T_global > 0) {                           // level-major flat index in the WHOLE timeline
                flat = (win.t0 >> l) + t;
                for (int ll = 0; ll < l; ll++) flat += win.T_global >> ll;
            }
        }
        int tk;
        const int pos = wbase + block_excl_scan(keep, scratch, tk);
        if (keep) {
            const int64_t o = (int64_t)q * topk + pos;
            cand_segs[o * 2] = left; cand_segs[o * 2 + 1] = right;
            cand_scores[o] = score;
            cand_idx[o] = flat;
        }
        wbase += tk;
    }
    if (tid == 0) cand_count[q] = wbase;
}

// ------------------------------------------------------------------------------- soft-NMS
// exp used by the gaussian decay.  The reference calls glibc expf (std::exp(float),
// nms_cpu.cpp:150); CUDA expf is within 2 ulp of it, so decayed scores agree to ulps while the
// selection order is exact on tie-free inputs.
__device__ __forceinline__ float nms_expf(float x) { return expf(x); }

struct NmsState { float *x1, *x2, *sc, *ar; int *id; int *list; };

__device__ __forceinline__ float seg_iou(float ix1, float ix2, float iarea, float jx1, float jx2, float jarea) {
    const float xx1 = fmaxf(ix1, jx1), xx2 = fminf(ix2, jx2);
    const float inter = fmaxf(0.f, __fsub_rn(xx2, xx1));
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(iarea, jarea), inter));
}

__global__ void __launch_bounds__(NMS_THREADS)
softnms_kernel(const float *__restrict__ segs, const float *__restrict__ scores, const int32_t *__restrict__ n_in,
               int cand_stride, float *__restrict__ dets, int32_t *__restrict__ inds, int32_t *__restrict__ n_out,
               float iou_thresh, float sigma, float min_score, int method, int max_iters, int smem_cap,
               char *__restrict__ workspace) {
    extern __shared__ __align__(16) char smem_raw[];
    __shared__ int scratch[34];
    __shared__ float red_s[32];
    __shared__ int red_p[32];
    __shared__ float s_ix1, s_ix2, s_iarea;
    const int q = blockIdx.x, tid = threadIdx.x;
    int n = n_in[q];
    n = max(0, min(n, cand_stride));
    NmsState st;
    if (n <= smem_cap) {
        float *f = reinterpret_cast<float *>(smem_raw);
        st.x1 = f; st.x2 = f + smem_cap; st.sc = f + 2 * smem_cap; st.ar = f + 3 * smem_cap;
        st.id = reinterpret_cast<int *>(f + 4 * smem_cap);
        st.list = st.id + smem_cap;
    } else {
        float *f = reinterpret_cast<float *>(workspace) + (int64_t)q * cand_stride * 6;
        st.x1 = f; st.x2 = f + cand_stride; st.sc = f + 2 * (int64_t)cand_stride; st.ar = f + 3 * (int64_t)cand_stride;
        st.id = reinterpret_cast<int *>(f + 4 * (int64_t)cand_stride);
        st.list = st.id + cand_stride;
    }
    const float *sg = segs + (int64_t)q * cand_stride * 2;
    const float *sc_in = scores + (int64_t)q * cand_stride;
    for (int i = tid; i < n; i += NMS_THREADS) {
        const float a = sg[2 * i], b = sg[2 * i + 1];
        st.x1[i] = a; st.x2[i] = b; st.sc[i] = sc_in[i];
        st.ar[i] = __fadd_rn(__fsub_rn(b, a), 1e-6f);         // areas = x2 - x1 + 1e-6
        st.id[i] = i;
    }
    __syncthreads();
    float *d_out = dets + (int64_t)q * cand_stride * 3;
    int32_t *i_out = inds + (int64_t)q * cand_stride;

    int it = 0;
    for (;; it++) {
        if (it >= n) break;
        if (max_iters > 0 && it >= max_iters) break;
        // (a) arg-max over [it, n): first position wins ties (strict '<' in the reference)
        float best = -INFINITY; int bpos = 0x7fffffff;
        for (int p = it + tid; p < n; p += NMS_THREADS) {
            const float s = st.sc[p];
            if (bpos == 0x7fffffff || s > best) { best = s; bpos = p; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float os = __shfl_xor_sync(0xffffffffu, best, o);
            const int op = __shfl_xor_sync(0xffffffffu, bpos, o);
            if (op != 0x7fffffff && (bpos == 0x7fffffff || os > best || (os == best && op < bpos))) { best = os; bpos = op; }
        }
        if ((tid & 31) == 0) { red_s[tid >> 5] = best; red_p[tid >> 5] = bpos; }
        __syncthreads();
        if (tid < 32) {
            best = red_s[tid]; bpos = red_p[tid];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float os = __shfl_xor_sync(0xffffffffu, best, o);
                const int op = __shfl_xor_sync(0xffffffffu, bpos, o);
                if (op != 0x7fffffff && (bpos == 0x7fffffff || os > best || (os == best && op < bpos))) { best = os; bpos = op; }
            }
            if (tid == 0) {
                // (b) swap position `it` with the arg-max; emit the detection
                const int mp = bpos;
                const float ix1 = st.x1[mp], ix2 = st.x2[mp], isc = st.sc[mp], iar = st.ar[mp];
                const int iid = st.id[mp];
                st.x1[mp] = st.x1[it]; st.x2[mp] = st.x2[it]; st.sc[mp] = st.sc[it]; st.ar[mp] = st.ar[it]; st.id[mp] = st.id[it];
                st.x1[it] = ix1; st.x2[it] = ix2; st.sc[it] = isc; st.ar[it] = iar; st.id[it] = iid;
                d_out[it * 3] = ix1; d_out[it * 3 + 1] = ix2; d_out[it * 3 + 2] = isc;
                i_out[it] = iid;
                s_ix1 = ix1; s_ix2 = ix2; s_iarea = iar;
            }
        }
        __syncthreads();
        const float ix1 = s_ix1, ix2 = s_ix2, iar = s_iarea;
        // (c) decay every later element once; count survivors
        int surv = 0;
        for (int p = it + 1 + tid; p < n; p += NMS_THREADS) {
            const float ovr = seg_iou(ix1, ix2, iar, st.x1[p], st.x2[p], st.ar[p]);
            float w = 1.f;
            if (method == 0) { if (ovr >= iou_thresh) w = 0.f; }
            else if (method == 1) { if (ovr >= iou_thresh) w = __fsub_rn(1.f, ovr); }
            else if (method == 2) { w = nms_expf(__fdiv_rn(-__fmul_rn(ovr, ovr), sigma)); }
            const float s = __fmul_rn(st.sc[p], w);
            st.sc[p] = s;
            surv += !(s < min_score);
        }
        int S;
        block_excl_scan(surv, scratch, S);
        const int n_final = it + 1 + S;
        if (n_final != n) {
            // (d) the reference prunes by "overwrite with the last element, shrink, re-examine":
            // net effect = holes (pruned, position < n_final) are filled in ascending order by the
            // survivors at positions >= n_final taken from the end backwards.
            int hbase = 0;
            for (int p0 = it + 1; p0 < n_final; p0 += NMS_THREADS) {
                const int p = p0 + tid;
                const int hole = p < n_final && st.sc[p] < min_score;
                int th;
                const int r = hbase + block_excl_scan(hole, scratch, th);
                if (hole) st.list[r] = p;
                hbase += th;
            }
            __syncthreads();
            int mbase = 0;
            for (int p0 = n - 1; p0 >= n_final; p0 -= NMS_THREADS) {
                const int p = p0 - tid;
                const int mover = p >= n_final && !(st.sc[p] < min_score);
                int tm;
                const int r = mbase + block_excl_scan(mover, scratch, tm);
                if (mover) {
                    const int d = st.list[r];
                    st.x1[d] = st.x1[p]; st.x2[d] = st.x2[p]; st.sc[d] = st.sc[p]; st.ar[d] = st.ar[p]; st.id[d] = st.id[p];
                }
                mbase += tm;
            }
            n = n_final;
        }
        __syncthreads();
    }
    if (tid == 0) {
        // reference returns inds[:nsegs] after running to completion; with max_iters the first
        // `it` rows are the ones libs/nms/nms.py consumes
        n_out[q] = it;                                     // full run: it == n == nsegs
    }
}

// ------------------------------------------------------------------------------- hard NMS
__global__ void __launch_bounds__(NMS_THREADS)
hardnms_kernel(const float *__restrict__ segs, const float *__restrict__ scores, const int32_t *__restrict__ n_in,
               int cand_stride, int32_t *__restrict__ keep, int32_t *__restrict__ n_out, float iou_thresh,
               float min_score, int max_keep, int ns_pow2) {
    extern __shared__ __align__(16) char smem_raw[];
    unsigned long long *order = reinterpret_cast<unsigned long long *>(smem_raw);     // [ns_pow2]
    float *kx1 = reinterpret_cast<float *>(order + ns_pow2);                            // kept list [ns_pow2] x3
    float *kx2 = kx1 + ns_pow2, *kar = kx2 + ns_pow2;
    const int q = blockIdx.x, tid = threadIdx.x;
    int n = n_in[q];
    n = max(0, min(n, cand_stride));
    const float *sg = segs + (int64_t)q * cand_stride * 2;
    const float *sc = scores + (int64_t)q * cand_stride;
    for (int i = tid; i < ns_pow2; i += NMS_THREADS) {
        unsigned long long e = ~0ull;
        if (i < n) {
            const float s = sc[i];
            if (!(min_score > 0.f) || s > min_score) e = ((unsigned long long)(~f2key(s)) << 32) | (uint32_t)i;
        }
        order[i] = e;
    }
    __syncthreads();
    block_bitonic_sort(order, ns_pow2);
    if (tid >= 32) return;
    // greedy suppression by one warp, 32 candidates at a time
    const int lane = tid;
    int kept = 0;
    int32_t *kout = keep + (int64_t)q * cand_stride;
    for (int b0 = 0; b0 < n; b0 += 32) {
        const unsigned long long e = order[min(b0 + lane, ns_pow2 - 1)];
        bool alive = (b0 + lane) < n && e != ~0ull;
        const int idx = (int)(e & 0xffffffffu);
        float x1 = 0.f, x2 = 0.f, ar = 0.f;
        if (alive) {
            x1 = sg[2 * idx]; x2 = sg[2 * idx + 1];
            ar = __fadd_rn(__fsub_rn(x2, x1), 1e-6f);
            for (int j = 0; j < kept && alive; j++)
                if (seg_iou(kx1[j], kx2[j], kar[j], x1, x2, ar) >= iou_thresh) alive = false;
        }
        for (;;) {
            const unsigned ball = __ballot_sync(0xffffffffu, alive);
            if (!ball) break;
            const int f = __ffs(ball) - 1;
            const float fx1 = __shfl_sync(0xffffffffu, x1, f), fx2 = __shfl_sync(0xffffffffu, x2, f);
            const float far_ = __shfl_sync(0xffffffffu, ar, f);
            const int fidx = __shfl_sync(0xffffffffu, idx, f);
            if (lane == 0) { kx1[kept] = fx1; kx2[kept] = fx2; kar[kept] = far_; kout[kept] = fidx; }
            kept++;
            if (lane == f) alive = false;
            else if (alive && lane > f && seg_iou(fx1, fx2, far_, x1, x2, ar) >= iou_thresh) alive = false;
            if (max_keep > 0 && kept >= max_keep) break;
        }
        __syncwarp();
        if (max_keep > 0 && kept >= max_keep) break;
    }
    if (lane == 0) n_out[q] = kept;
}

// Hard NMS for candidate lists beyond the shared-memory sort (> 4096): only the first max_keep survivors are ever
// consumed (libs/nms/nms.py:25-26), so instead of sorting, repeat max_keep times: block arg-max over the live
// candidates (ties: lower index first = the stable order of the sorted variant), emit it, kill everything with
// IoU >= thresh.  O(max_keep * n), live scores in the caller's workspace.
__global__ void __launch_bounds__(NMS_THREADS)
hardnms_large_kernel(const float *__restrict__ segs, const float *__restrict__ scores, const int32_t *__restrict__ n_in,
                     int cand_stride, int32_t *__restrict__ keep, int32_t *__restrict__ n_out, float iou_thresh,
                     float min_score, int max_keep, float *__restrict__ work) {
    __shared__ float red_s[32];
    __shared__ int red_p[32];
    __shared__ float s_x1, s_x2, s_ar;
    __shared__ int s_idx;
    const int q = blockIdx.x, tid = threadIdx.x;
    int n = n_in[q];
    n = max(0, min(n, cand_stride));
    const float *sg = segs + (int64_t)q * cand_stride * 2;
    const float *sc_in = scores + (int64_t)q * cand_stride;
    float *live = work + (int64_t)q * cand_stride;
    for (int i = tid; i < n; i += NMS_THREADS) {
        const float s = sc_in[i];
        live[i] = (!(min_score > 0.f) || s > min_score) ? s : -INFINITY;
    }
    __syncthreads();
    int32_t *kout = keep + (int64_t)q * cand_stride;
    int kept = 0;
    while (kept < max_keep) {
        float best = -INFINITY; int bpos = 0x7fffffff;
        for (int i = tid; i < n; i += NMS_THREADS) {
            const float s = live[i];
            if (s > best) { best = s; bpos = i; }             // ascending i per thread: the first maximum stays
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float os = __shfl_xor_sync(0xffffffffu, best, o);
            const int op = __shfl_xor_sync(0xffffffffu, bpos, o);
            if (os > best || (os == best && op < bpos)) { best = os; bpos = op; }
        }
        if ((tid & 31) == 0) { red_s[tid >> 5] = best; red_p[tid >> 5] = bpos; }
        __syncthreads();
        if (tid < 32) {
            best = tid < NMS_THREADS / 32 ? red_s[tid] : -INFINITY;
            bpos = tid < NMS_THREADS / 32 ? red_p[tid] : 0x7fffffff;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float os = __shfl_xor_sync(0xffffffffu, best, o);
                const int op = __shfl_xor_sync(0xffffffffu, bpos, o);
                if (os > best || (os == best && op < bpos)) { best = os; bpos = op; }
            }
            if (tid == 0) {
                s_idx = best > -INFINITY ? bpos : -1;
                if (s_idx >= 0) {
                    s_x1 = sg[2 * bpos]; s_x2 = sg[2 * bpos + 1];
                    s_ar = __fadd_rn(__fsub_rn(s_x2, s_x1), 1e-6f);
                    kout[kept] = bpos;
                }
            }
        }
        __syncthreads();
        const int idx = s_idx;
        if (idx < 0) break;
        kept++;
        const float x1 = s_x1, x2 = s_x2, ar = s_ar;
        for (int i = tid; i < n; i += NMS_THREADS) {
            if (live[i] > -INFINITY) {
                const float a = sg[2 * i], b = sg[2 * i + 1];
                if (i == idx || seg_iou(x1, x2, ar, a, b, __fadd_rn(__fsub_rn(b, a), 1e-6f)) >= iou_thresh) live[i] = -INFINITY;
            }
        }
        __syncthreads();
    }
    if (tid == 0) n_out[q] = kept;
}

// ------------------------------------------------------------------------------- voting + finalize
// k rows per query (from soft-NMS dets or from hard-NMS keep indices) -> segment voting over all
// input candidates -> stable descending sort -> seconds.
__global__ void __launch_bounds__(256)
nms_finalize_kernel(const float *__restrict__ segs, const float *__restrict__ scores, const int32_t *__restrict__ n_in,
                    int cand_stride, const float *__restrict__ dets, const int32_t *__restrict__ keep,
                    const int32_t *__restrict__ k_in, decaf_nms_params_t prm, int max_out,
                    float *__restrict__ out_segs, float *__restrict__ out_scores, int32_t *__restrict__ out_count,
                    float *__restrict__ tmp /* (n_query, cand_stride, 3) voted rows */) {
    __shared__ double red[3][8];
    const int q = blockIdx.x, tid = threadIdx.x;
    int n = n_in[q];
    n = max(0, min(n, cand_stride));
    int k = n > 0 ? k_in[q] : 0;
    if (prm.max_num_segs > 0) k = min(k, prm.max_num_segs);
    const float *sg = segs + (int64_t)q * cand_stride * 2;
    const float *sc = scores + (int64_t)q * cand_stride;
    float *row = tmp + (int64_t)q * cand_stride * 3;
    for (int j = 0; j < k; j++) {
        float a, b, s;
        if (dets) { const float *d = dets + ((int64_t)q * cand_stride + j) * 3; a = d[0]; b = d[1]; s = d[2]; }
        else { const int i = keep[(int64_t)q * cand_stride + j]; a = sg[2 * i]; b = sg[2 * i + 1]; s = sc[i]; }
        if (prm.voting_thresh > 0.f) {
            // segment_voting (libs/nms/nms.py:82-101); sums in double
            double sw = 0.0, s1 = 0.0, s2 = 0.0;
            const float la = __fsub_rn(b, a);
            for (int i = tid; i < n; i += blockDim.x) {
                const float c = sg[2 * i], d = sg[2 * i + 1];
                const float left = fmaxf(a, c), right = fminf(b, d);
                const float ov = fmaxf(__fsub_rn(right, left), 0.f);
                const float uni = __fsub_rn(__fadd_rn(la, __fsub_rn(d, c)), ov);
                const float iou = __fdiv_rn(ov, uni);
                if (iou >= prm.voting_thresh) {
                    const double w = (double)sc[i];
                    sw += w; s1 += w * (double)c; s2 += w * (double)d;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                sw += __shfl_xor_sync(0xffffffffu, sw, o);
                s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            }
            if ((tid & 31) == 0) { red[0][tid >> 5] = sw; red[1][tid >> 5] = s1; red[2][tid >> 5] = s2; }
            __syncthreads();
            if (tid == 0) {
                double tw = 0, t1 = 0, t2 = 0;
                for (int w = 0; w < 8; w++) { tw += red[0][w]; t1 += red[1][w]; t2 += red[2][w]; }
                row[j * 3] = (float)(t1 / tw); row[j * 3 + 1] = (float)(t2 / tw); row[j * 3 + 2] = s;
            }
            __syncthreads();
        } else if (tid == 0) {
            row[j * 3] = a; row[j * 3 + 1] = b; row[j * 3 + 2] = s;
        }
    }
    __syncthreads();
    // stable descending rank sort + seconds conversion
    const int kout = min(k, max_out);
    for (int j = tid; j < k; j += blockDim.x) {
        const float s = row[j * 3 + 2];
        int rank = 0;
        for (int i = 0; i < k; i++) {
            const float t = row[i * 3 + 2];
            rank += (t > s) || (t == s && i < j);
        }
        if (rank < kout) {
            float a = row[j * 3], b = row[j * 3 + 1];
            if (prm.to_seconds) {
                float vs = prm.vid_stride, cst = prm.clip_stride, hc = prm.half_clip_size, fps = prm.fps, dur = prm.duration;
                if (prm.video_meta) {
                    vs = prm.video_meta[0]; cst = prm.video_meta[1]; hc = prm.video_meta[2]; fps = prm.video_meta[3];
                    dur = prm.video_meta[4];
                }
                a = __fdiv_rn(__fadd_rn(__fmul_rn(__fmul_rn(a, vs), cst), hc), fps);
                b = __fdiv_rn(__fadd_rn(__fmul_rn(__fmul_rn(b, vs), cst), hc), fps);
                a = fminf(fmaxf(a, 0.f), dur);
                b = fminf(fmaxf(b, 0.f), dur);
            }
            const int64_t o = (int64_t)q * max_out + rank;
            out_segs[o * 2] = a; out_segs[o * 2 + 1] = b; out_scores[o] = s;
        }
    }
    if (tid == 0) out_count[q] = kout;
}

static int pow2_at_least(int n) { int p = 32; while (p < n) p <<= 1; return p; }

}  // namespace decaf

using namespace decaf;

static int decode_launch(const float *logits, const float *offsets, const uint8_t *hmask, const decaf_levels_t *lv,
                         int32_t n_query, int32_t from_logits, float pre_nms_thresh, int32_t topk, float seg_len_thresh,
                         float *cand_segs, float *cand_scores, int32_t *cand_idx, int32_t *cand_count,
                         const decaf_decode_window_t &win, void *stream) {
    DECAF_CHECK(logits && offsets && hmask && lv && cand_segs && cand_scores && cand_idx && cand_count,
                "decaf_decode: null pointers");
    DECAF_CHECK(topk > 0 && topk <= 4096, "decaf_decode: topk must be in 1..4096 (got %d)", topk);
    if (n_query == 0) return 0;
    const int ns = pow2_at_least(topk);
    const size_t smem = (size_t)ns * sizeof(unsigned long long);
    if (smem > 32 * 1024)
        DECAF_CUDA(cudaFuncSetAttribute(decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    decode_kernel<<<n_query, DEC_THREADS, smem, as_stream(stream)>>>(logits, offsets, hmask, *lv, from_logits, pre_nms_thresh,
                                                                      topk, ns, seg_len_thresh, cand_segs, cand_scores,
                                                                      cand_idx, cand_count, win);
    DECAF_LAUNCH_CHECK();
    return 0;
}

extern "C" int decaf_decode(const float *logits, const float *offsets, const uint8_t *hmask,
                            const decaf_levels_t *lv, int32_t n_query, int32_t from_logits, float pre_nms_thresh,
                            int32_t topk, float seg_len_thresh, float *cand_segs, float *cand_scores,
                            int32_t *cand_idx, int32_t *cand_count, void *stream) {
    decaf_decode_window_t win = {0, 0, 0, 0};
    return decode_launch(logits, offsets, hmask, lv, n_query, from_logits, pre_nms_thresh, topk, seg_len_thresh, cand_segs,
                         cand_scores, cand_idx, cand_count, win, stream);
}

extern "C" int decaf_decode_window(const float *logits, const float *offsets, const uint8_t *hmask,
                                   const decaf_levels_t *lv, int32_t n_query, int32_t from_logits, float pre_nms_thresh,
                                   int32_t topk, float seg_len_thresh, const decaf_decode_window_t *win, float *cand_segs,
                                   float *cand_scores, int32_t *cand_idx, int32_t *cand_count, void *stream) {
    DECAF_CHECK(win, "decaf_decode_window: null window");
    const int align = 1 << (lv ? lv->n_levels - 1 : 0);
    DECAF_CHECK(win->t0 % align == 0 && win->own_lo % align == 0 && win->own_hi % align == 0,
                "decaf_decode_window: window origin / owned range must be multiples of 2^(levels-1) = %d", align);
    DECAF_CHECK(win->T_global < (1 << 24), "decaf_decode_window: timeline too long for exact fp32 coordinates");
    // global flat point indices feed decaf_merge_candidates, whose 64-bit sort key carries them in 18 bits
    int64_t p_global = 0;
    for (int l = 0; lv && l < lv->n_levels; l++) p_global += win->T_global >> l;
    DECAF_CHECK(p_global < (1 << 18), "decaf_decode_window: %lld global points do not fit the 18-bit index field of the "
                "candidate merge key (T_global %d)", (long long)p_global, win->T_global);
    return decode_launch(logits, offsets, hmask, lv, n_query, from_logits, pre_nms_thresh, topk, seg_len_thresh, cand_segs,
                         cand_scores, cand_idx, cand_count, *win, stream);
}

// ------------------------------------------------------------------------------- eval-time loss statistics
// Evaluator._calc_loss (libs/worker_v2.py:1029-1061): per query, over its points (level-major) —
//   labels / GT offsets of annotate_points_per_video (:93-133): inside the centre-sampling window AND event within the
//   level's regression range; focal loss (label smoothing 0.2, alpha 0.5, gamma 2; libs/modeling/loss.py:6-58) summed over
//   the valid points; 1 - IoU of (predicted, GT) offsets (ctr_giou_loss, :61-108; the reference passes reg_loss='iou')
//   summed over the positive valid points; the number of positives.
// One CTA per query, fixed-order block reduction (deterministic).  out[q] = {cls_sum, reg_sum, n_pos}.
constexpr int EL_THREADS = 512;
__global__ void __launch_bounds__(EL_THREADS)
eval_loss_kernel(const float *__restrict__ logits, const float *__restrict__ offsets, const uint8_t *__restrict__ hmask,
                 decaf_levels_t lv, const float *__restrict__ targets, const float *__restrict__ reg_range, int center_sampling,
                 float radius_mul, float smoothing, float alpha, float *__restrict__ out) {
    __shared__ float red[3][EL_THREADS / 32];
    const int q = blockIdx.x, tid = threadIdx.x;
    const int64_t base = (int64_t)q * lv.Pp;
    const float t0 = targets[2 * q], t1 = targets[2 * q + 1];
    float cls = 0.f, reg = 0.f, npos = 0.f;
    for (int r = tid; r < lv.Pp; r += EL_THREADS) {
        const int l = level_of(lv, r);
        if (l < 0 || !hmask[base + r]) continue;
        const float stride = (float)(1 << l);
        const float pt = (float)(r - lv.off[l]) * stride;
        const float d0 = pt - t0, d1 = t1 - pt;                 // distance to the segment boundaries
        bool inside;
        if (center_sampling) {
            const float ctr = 0.5f * (t0 + t1), rad = stride * radius_mul;
            const float lo = fmaxf(ctr - rad, t0), hi = fminf(ctr + rad, t1);
            inside = (pt - lo > 0.f) && (hi - pt > 0.f);
        } else {
            inside = d0 > 0.f && d1 > 0.f;
        }
        const float md = fmaxf(d0, d1);
        const bool pos = inside && md >= reg_range[2 * l] && md < reg_range[2 * l + 1];
        // focal loss with smoothed targets
        const float x = logits[base + r];
        const float tg = (pos ? 1.f : 0.f) * (1.f - smoothing) + 0.5f * smoothing;
        const float pr = 1.0f / (1.0f + expf(-x));
        const float p_t = pr * tg + (1.f - pr) * (1.f - tg);
        const float ce = fmaxf(x, 0.f) - x * tg + log1pf(expf(-fabsf(x)));
        float lo_ = ce * (1.f - p_t) * (1.f - p_t);
        if (alpha >= 0.f) lo_ *= (tg >= 0.5f) ? alpha : (1.f - alpha);
        cls += lo_;
        if (pos) {
            const float lp = offsets[2 * (base + r)], rp = offsets[2 * (base + r) + 1];
            const float lg = d0 / stride, rg = d1 / stride;
            const float inter = fminf(lp, lg) + fminf(rp, rg);
            const float uni = (lp + rp) + (lg + rg) - inter;
            reg += 1.f - inter / fmaxf(uni, 1e-8f);
            npos += 1.f;
        }
    }
    float v[3] = {cls, reg, npos};
#pragma unroll
    for (int k = 0; k < 3; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
        if ((tid & 31) == 0) red[k][tid >> 5] = v[k];
    }
    __syncthreads();
    if (tid < 3) {
        float s = 0.f;
        for (int w = 0; w < EL_THREADS / 32; w++) s += red[tid][w];
        out[3 * q + tid] = s;
    }
}

extern "C" int decaf_eval_loss(const float *logits, const float *offsets, const uint8_t *hmask, const decaf_levels_t *lv,
                               int32_t n_query, const float *targets, const float *reg_range, int32_t center_sampling,
                               float radius_mul, float smoothing, float alpha, float *out, void *stream) {
    DECAF_CHECK(logits && offsets && hmask && lv && targets && reg_range && out, "decaf_eval_loss: null pointers");
    if (n_query == 0) return 0;
    eval_loss_kernel<<<n_query, EL_THREADS, 0, as_stream(stream)>>>(logits, offsets, hmask, *lv, targets, reg_range, center_sampling,
                                                                    radius_mul, smoothing, alpha, out);
    DECAF_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------- candidate merge (time shards)
// Per query: the union of `n_src` per-shard candidate lists (each the shard's own top-k) -> global top-k by
// (score descending, global flat point index ascending) = exactly the order the unsharded decode produces.
// key = ~scorekey : idx (18 bits) : source slot (14 bits); one CTA per query, bitonic sort in shared memory.
__global__ void __launch_bounds__(DEC_THREADS)
merge_candidates_kernel(const float *__restrict__ segs, const float *__restrict__ scores, const int32_t *__restrict__ idx,
                        const int32_t *__restrict__ count, int n_src, int n_query, int topk, int ns_pow2,
                        float *__restrict__ out_segs, float *__restrict__ out_scores, int32_t *__restrict__ out_idx,
                        int32_t *__restrict__ out_count) {
    extern __shared__ unsigned long long mkeys[];             // [ns_pow2]
    __shared__ int s_total;
    const int q = blockIdx.x, tid = threadIdx.x;
    const int slots = n_src * topk;
    int total = 0;
    for (int i = tid; i < ns_pow2; i += DEC_THREADS) {
        unsigned long long key = ~0ull;
        if (i < slots) {
            const int src = i / topk, j = i % topk;
            const int64_t o = ((int64_t)src * n_query + q) * topk + j;
            if (j < count[src * n_query + q]) {
                key = ((unsigned long long)(~f2key(scores[o])) << 32) | ((unsigned long long)(uint32_t)idx[o] << 14) | (uint32_t)i;
                total++;
            }
        }
        mkeys[i] = key;
    }
    total = warp_sum_i(total);
    if (tid == 0) s_total = 0;
    __syncthreads();
    if ((tid & 31) == 0) atomicAdd(&s_total, total);
    __syncthreads();
    block_bitonic_sort(mkeys, ns_pow2);
    const int n_out = min(s_total, topk);
    for (int i = tid; i < n_out; i += DEC_THREADS) {
        const unsigned long long e = mkeys[i];
        const int slot = (int)(e & 0x3fffu);
        const int src = slot / topk, j = slot % topk;
        const int64_t o = ((int64_t)src * n_query + q) * topk + j, w = (int64_t)q * topk + i;
        out_segs[w * 2] = segs[o * 2]; out_segs[w * 2 + 1] = segs[o * 2 + 1];
        out_scores[w] = scores[o];
        out_idx[w] = idx[o];
    }
    if (tid == 0) out_count[q] = n_out;
}

extern "C" int decaf_merge_candidates(const float *segs, const float *scores, const int32_t *idx, const int32_t *count,
                                      int32_t n_src, int32_t n_query, int32_t topk, float *out_segs, float *out_scores,
                                      int32_t *out_idx, int32_t *out_count, void *stream) {
    DECAF_CHECK(segs && scores && idx && count && out_segs && out_scores && out_idx && out_count,
                "decaf_merge_candidates: null pointers");
    DECAF_CHECK(n_src >= 1 && topk >= 1 && (int64_t)n_src * topk <= 16384,
                "decaf_merge_candidates: n_src * topk must be <= 16384 (got %d x %d)", n_src, topk);
    if (n_query == 0) return 0;
    const int ns = pow2_at_least(n_src * topk);
    const size_t smem = (size_t)ns * sizeof(unsigned long long);
    if (smem > 32 * 1024)
        DECAF_CUDA(cudaFuncSetAttribute(merge_candidates_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    merge_candidates_kernel<<<n_query, DEC_THREADS, smem, as_stream(stream)>>>(segs, scores, idx, count, n_src, n_query, topk, ns,
                                                                               out_segs, out_scores, out_idx, out_count);
    DECAF_LAUNCH_CHECK();
    return 0;
}

extern "C" int64_t decaf_nms_workspace_bytes(int32_t n_query, int32_t max_n) {
    // soft-NMS state (6 words / candidate) + dets/inds/voted rows for the fused call + counters
    return (int64_t)n_query * max_n * (6 + 3 + 1 + 3) * 4 + (int64_t)n_query * 16 + 256;
}

static int softnms_launch(const float *segs, const float *scores, const int32_t *n, int n_query, int cand_stride,
                          float *dets, int32_t *inds, int32_t *n_out, float iou_thresh, float sigma, float min_score,
                          int method, int max_iters, char *ws_state, cudaStream_t st) {
    int cap = pow2_at_least(cand_stride);
    size_t smem = 0;
    if (cap <= 4096) smem = (size_t)cap * 6 * 4; else cap = 0;
    if (smem > 32 * 1024)      // static shared memory (scan scratch) counts towards the 48 KB default limit
        DECAF_CUDA(cudaFuncSetAttribute(softnms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    softnms_kernel<<<n_query, NMS_THREADS, smem, st>>>(segs, scores, n, cand_stride, dets, inds, n_out, iou_thresh, sigma,
                                                       min_score, method, max_iters, cap, ws_state);
    DECAF_LAUNCH_CHECK();
    return 0;
}

extern "C" int decaf_softnms_1d(const float *segs, const float *scores, const int32_t *n, int32_t n_query,
                                int32_t cand_stride, float *dets, int32_t *inds, int32_t *n_out, float iou_thresh,
                                float sigma, float min_score, int32_t method, int32_t max_iters, void *workspace,
                                void *stream) {
    DECAF_CHECK(segs && scores && n && dets && inds && n_out, "decaf_softnms_1d: null pointers");
    DECAF_CHECK(method >= 0 && method <= 2, "decaf_softnms_1d: method must be 0, 1 or 2");
    DECAF_CHECK(cand_stride > 0, "decaf_softnms_1d: cand_stride must be > 0");
    DECAF_CHECK(cand_stride <= 4096 || workspace, "decaf_softnms_1d: workspace required for > 4096 candidates");
    if (n_query == 0) return 0;
    return softnms_launch(segs, scores, n, n_query, cand_stride, dets, inds, n_out, iou_thresh, sigma, min_score, method,
                          max_iters, (char *)workspace, as_stream(stream));
}

static int hardnms_launch(const float *segs, const float *scores, const int32_t *n, int n_query, int cand_stride,
                          int32_t *keep, int32_t *n_out, float iou_thresh, float min_score, int max_keep, float *work,
                          cudaStream_t st) {
    if (cand_stride > 4096) {
        DECAF_CHECK(max_keep > 0 && work, "decaf_nms_1d: more than 4096 candidates per query need max_keep > 0 and a workspace "
                                          "of n_query * cand_stride floats (got %d candidates)", cand_stride);
        hardnms_large_kernel<<<n_query, NMS_THREADS, 0, st>>>(segs, scores, n, cand_stride, keep, n_out, iou_thresh, min_score,
                                                              max_keep, work);
        DECAF_LAUNCH_CHECK();
        return 0;
    }
    const int ns = pow2_at_least(cand_stride);
    const size_t smem = (size_t)ns * (8 + 12);
    if (smem > 32 * 1024)
        DECAF_CUDA(cudaFuncSetAttribute(hardnms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    hardnms_kernel<<<n_query, NMS_THREADS, smem, st>>>(segs, scores, n, cand_stride, keep, n_out, iou_thresh, min_score,
                                                       max_keep, ns);
    DECAF_LAUNCH_CHECK();
    return 0;
}

extern "C" int decaf_nms_1d(const float *segs, const float *scores, const int32_t *n, int32_t n_query,
                            int32_t cand_stride, int32_t *keep, int32_t *n_out, float iou_thresh, float min_score,
                            int32_t max_keep, void *workspace, void *stream) {
    DECAF_CHECK(segs && scores && n && keep && n_out, "decaf_nms_1d: null pointers");
    DECAF_CHECK(cand_stride > 0, "decaf_nms_1d: cand_stride must be > 0");
    if (n_query == 0) return 0;
    return hardnms_launch(segs, scores, n, n_query, cand_stride, keep, n_out, iou_thresh, min_score, max_keep,
                          reinterpret_cast<float *>(workspace), as_stream(stream));
}

extern "C" int decaf_batched_nms(const float *segs, const float *scores, const int32_t *n, int32_t n_query,
                                 int32_t cand_stride, const decaf_nms_params_t *prm, float *out_segs,
                                 float *out_scores, int32_t *out_count, void *workspace, void *stream) {
    DECAF_CHECK(segs && scores && n && prm && out_segs && out_scores && out_count && workspace,
                "decaf_batched_nms: null pointers");
    DECAF_CHECK(prm->mode >= 0 && prm->mode <= 2, "decaf_batched_nms: invalid NMS mode %d", prm->mode);
    DECAF_CHECK(cand_stride > 0, "decaf_batched_nms: cand_stride must be > 0");
    if (n_query == 0) return 0;
    cudaStream_t st = as_stream(stream);
    // workspace carve-up (see decaf_nms_workspace_bytes)
    float *ws = reinterpret_cast<float *>(workspace);
    const int64_t per = (int64_t)n_query * cand_stride;
    char *state = reinterpret_cast<char *>(ws);                     // 6 words / candidate
    float *dets = ws + per * 6;                                      // 3 words
    int32_t *inds = reinterpret_cast<int32_t *>(ws + per * 9);       // 1 word
    float *voted = ws + per * 10;                                    // 3 words
    int32_t *cnt = reinterpret_cast<int32_t *>(ws + per * 13);
    const int max_out = prm->max_num_segs > 0 ? prm->max_num_segs : cand_stride;
    if (prm->mode == 2) {
        if (softnms_launch(segs, scores, n, n_query, cand_stride, dets, inds, cnt, prm->iou_thresh, prm->sigma,
                           prm->min_score, 2, prm->max_num_segs > 0 ? prm->max_num_segs : 0, state, st))
            return 1;
        nms_finalize_kernel<<<n_query, 256, 0, st>>>(segs, scores, n, cand_stride, dets, nullptr, cnt, *prm, max_out,
                                                     out_segs, out_scores, out_count, voted);
    } else {
        // mode 1: hard NMS; mode 0 (no NMS): same kernel with nothing suppressed = stable sort by score
        const float thr = prm->mode == 1 ? prm->iou_thresh : INFINITY;
        const float ms = prm->mode == 1 ? prm->min_score : 0.f;
        if (hardnms_launch(segs, scores, n, n_query, cand_stride, inds, cnt, thr, ms,
                           prm->max_num_segs > 0 ? prm->max_num_segs : 0, reinterpret_cast<float *>(state), st))
            return 1;
        decaf_nms_params_t p2 = *prm;
        if (prm->mode == 0) p2.voting_thresh = 0.f;              // voting only runs when mode is not None
        nms_finalize_kernel<<<n_query, 256, 0, st>>>(segs, scores, n, cand_stride, nullptr, inds, cnt, p2, max_out,
                                                     out_segs, out_scores, out_count, voted);
    }
    DECAF_LAUNCH_CHECK();
    return 0;
}

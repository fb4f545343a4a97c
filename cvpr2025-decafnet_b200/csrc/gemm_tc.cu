// tcgen05 / TMEM / TMA bf16 GEMM + conv1d(k=3) for sm_100a — persistent, warp-specialised, TMA in AND out.
//
//   out[seq, t, n] = epi( sum_{tap, k} A[seq, t + (tap - taps/2) * dil, k] * W[n, tap, k] )
//
// One CTA (640 threads) per SM loops over 128 (time steps) x BN (output channels) tiles:
//   warp 0      : TMA producer — per k-iteration one A box (64 ch x 128 rows, 128B-swizzled) and the W
//                 boxes (64 ch x BN rows).  A k=3 convolution is an implicit GEMM: the three taps are three
//                 time-shifted TMA loads of the SAME activation tensor accumulating into the same TMEM
//                 tile; rows outside the sequence are zero-filled by TMA (= conv zero padding).
//                 Weight-resident mode: when the CTA's (group, n tile) weight block fits next to the A ring it
//                 is loaded once and the ring carries activations only.
//   warp 1      : allocates TMEM (512 columns = up to four accumulator stages), issues tcgen05.mma
//                 (kind::f16, bf16 x bf16 -> fp32, M = 128, N <= 256 per instruction, K = 16);
//                 tcgen05.commit releases each smem stage and signals the accumulator stage.
//   warp 2      : "C producer" — TMA-loads the fp32 addend of the epilogue (residual stream or PE table) chunk
//                 by chunk into the epilogue teams' shared-memory slabs, ahead of the epilogue.
//   warps 4..19 : epilogue, four TEAMS of four warps (one warp per TMEM lane quarter).  A team owns one
//                 128-row x 32-column chunk at a time: tcgen05.ld gives "thread = output row", the per-column
//                 parameters come from shared memory (broadcast reads), the fp32 addend from the team's slab
//                 (the TMA swizzle makes the row-per-thread access conflict-free), the results are written back
//                 into the slab(s) and leave through ONE TMA store per output tensor (cp.async.bulk.tensor,
//                 hardware-coalesced and clipped at the tensor edges).  No epilogue thread ever waits on a
//                 global load: the first version of this kernel fetched residual rows with ld.global from
//                 the epilogue warps and spent ~5 us per 32-column chunk in exposed latency.
//                 Optional fused channel LayerNorm (two-pass, like libs/modeling/blocks.py:125-131): the whole
//                 output row lives in one TMEM lane, so mean / variance are thread-local sums over the team's
//                 chunks, exchanged once per pass between the four warps that share a lane quarter.
// mbarrier rings: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue), slab full/empty
// (C producer <-> epilogue team).
// Streaming shapes (W not resident: k=3 convs with fused LayerNorm, K = 1024 FFN proj2) run in CTA-PAIR mode: a cluster of
// two CTAs issues tcgen05.mma.cta_group::2 (M = 256 over the two SMs of a TPC), each CTA loading its own 128 rows and half
// of every W block; see TcSched::pair and the PAIR template parameter below.
#include "tc_ptx.cuh"

namespace decaf {

constexpr int TBM = 128, TBK = 64;
constexpr int N_TEAMS = 4;
constexpr int EPI_WARP0 = 4;                          // first epilogue warp (multiple of 4: warp & 3 = TMEM lane quarter)
constexpr int EPI_WARPS = 4 * N_TEAMS;
constexpr int TC_THREADS = 32 * (EPI_WARP0 + EPI_WARPS);   // 640
constexpr int MAX_GROUP = 3;
constexpr int MAX_STAGES = 8;
constexpr int MAX_ACC = 4;
constexpr int A_BYTES = TBM * TBK * 2;                // 16 KB
constexpr int SLAB_F32 = TBM * 32 * 4;                // 128 rows x 32 fp32 (128-byte rows, SWIZZLE_128B)
constexpr int SLAB_B16 = TBM * 32 * 2;                // 128 rows x 32 bf16 (64-byte rows, SWIZZLE_64B)
constexpr int LNX_BYTES = 2 * N_TEAMS * TBM * 4;      // [pass][team][row]
constexpr int MAX_WB = 16;                            // per-k-block barriers of the resident W (blocks >= MAX_WB - 1 share the last)
constexpr int BAR_BYTES = (2 * MAX_STAGES + 2 * MAX_ACC + MAX_WB + 2 * N_TEAMS) * 8 + 16;
constexpr int SMEM_LIMIT = TC_SMEM_LIMIT;             // 227 KB
constexpr int MAX_PARAM_COLS = 2304;                  // bias columns (n_group * N) / colscale columns staged in smem (embd 512: FFN fc N = 2048)

struct TcMaps {
    CUtensorMap a[MAX_GROUP], w[MAX_GROUP];           // operands
    CUtensorMap of[MAX_GROUP], ob[MAX_GROUP];         // fp32 / bf16 outputs
    CUtensorMap add;                                  // fp32 addend: residual stream or PE table
};

struct TcSched {
    int BN;             // CTA tile width (output channels)
    int n_mma;          // MMAs per k-step (BN / n_mma columns each, <= 256)
    int acc_stages;     // TMEM accumulator stages
    int acc_stride;     // TMEM columns between stages
    int stages;         // smem pipeline depth
    int flat, tiles_per_seq, kb_per_tap;
    int k_tail_steps;   // K = 16 MMA steps of the LAST 64-channel k-block of a tap (4 unless K % 64 != 0: the zero-filled tail of a
                        // 288-channel row is not multiplied — 18 instead of 20 steps per tap)
    int m_tiles, n_tiles, n_group, total_tiles;
    int w_res;          // 1: the CTA's (group, n tile) weights stay resident in smem, the ring carries A only
    int nb16;           // bf16 slab buffers per team (2 = double buffered)
    int commit_every;   // G: the MMA thread commits (= releases ring stages) after every G-th k-block only — a tcgen05.commit holds
                        // the tensor pipe for ~190 cycles (measured: 724 cycles per k-block of 8 MMAs = 8 x 66.5 + 192), so one
                        // per k-block costs a quarter of the MMA time; the producer waits on the barrier of the stage that
                        // carries the group's commit
    int debug;          // timing experiments only (DECAF_GEMM_DBG, results are wrong): 1 skip the MMAs, 2 skip the W loads, 4 skip the A loads (pair mode)
    int lnr;            // > 0: register-resident LayerNorm epilogue, columns per team (kernel template parameter LNR)
    int add_is_pe;      // the addend is the PE table (indexed by the row inside the sequence, shared by all sequences)
    int cl;             // thread-block cluster size (1, 2 or 4): the CTAs of a cluster work on `cl` consecutive m tiles of the
                        // same (group, n tile) and share its weight blocks — each CTA loads 1/cl of every W block and
                        // TMA-multicasts it to all of them (the streaming mode is L2 -> SM bandwidth bound)
    int items;          // work items = ceil(m_tiles / cl) * n_tiles * n_group
    int pair;           // 1: CTA-pair mode (cl == 2): tcgen05.mma.cta_group::2, M = 256 over the two SMs of a TPC.  Each CTA
                        // loads its own 128 activation rows and HALF of every W block (the tensor core reads the other half
                        // from the peer's shared memory), which halves the W rows an SM has to ingest per k-block — the
                        // streaming shapes are paced by that ingest (~1 128-byte row per 2 cycles), not by the MMAs
    int off_w, off_f32, off_b16, off_lnx, off_bias, off_cs, off_lnw, off_lnb, off_bar;   // bytes from the aligned base
    unsigned long long *trace;   // debug: per-role clock64 stamps of CTA 0 (decaf_debug_gemm_trace), else NULL
};

constexpr int TRACE_SLOTS = 2048;                     // per role
__device__ __forceinline__ void trace_put(unsigned long long *tr, int role, int &n) {
    if (tr != nullptr && blockIdx.x == 0 && n < TRACE_SLOTS) tr[role * TRACE_SLOTS + n++] = clock64();
}

// ---------------------------------------------------------------------------------- kernel
// EPI >= 0 fixes the epilogue at compile time (bit 0 LN, bits 1-2 act, bit 3 colscale, bit 4 fp32 out,
// bit 5 bf16 out, bit 6 fp32 addend (residual or PE)); EPI < 0 is the generic variant that reads the flags
// from GemmArgs.
constexpr int epi_code(bool ln, int act, bool cs, bool f32, bool b16, bool add) {
    return (ln ? 1 : 0) | (act << 1) | (cs ? 8 : 0) | (f32 ? 16 : 0) | (b16 ? 32 : 0) | (add ? 64 : 0);
}

// Tile id -> (m tile, n tile, group).  combo = (group, n tile) is the fastest index so that (a) CTAs that
// run at the same time share the A tile through L2 and (b) with gridDim.x a multiple of `combos` every CTA
// keeps one combo for its whole life — the weight-resident mode relies on that.
struct TileIdx { int mt, nt, g; };
// `item` = work item of the CTA's cluster, `crank` = rank of the CTA inside it: the cluster takes cl consecutive m tiles
// of one combo; an m tile >= m_tiles is a phantom (its loads are zero-filled and its stores clipped by TMA) that keeps the
// cluster's pipelines in lockstep.  With cl == 1 an item is a tile.
__device__ __forceinline__ TileIdx decode_tile(const TcSched &sc, int item, int crank) {
    const int combos = sc.n_tiles * sc.n_group;
    const int combo = item % combos;
    TileIdx t;
    t.mt = (item / combos) * sc.cl + crank; t.nt = combo % sc.n_tiles; t.g = combo / sc.n_tiles;
    return t;
}
// first row of the tile as TMA coordinates (row inside the sequence / flat row, sequence)
__device__ __forceinline__ void tile_rows(const TcSched &sc, int mt, int &t0, int &seq) {
    if (sc.flat) { seq = 0; t0 = mt * TBM; }
    else { seq = mt / sc.tiles_per_seq; t0 = (mt % sc.tiles_per_seq) * TBM; }
}

// PAIR: CTA-pair instantiation (cta_group::2 instructions make a kernel launchable only as a cluster, so the
// single-CTA kernel must not contain them)
// LNR > 0: register-resident LayerNorm epilogue (conv -> LN -> ReLU -> bf16 with 256 < N = 4 * LNR <= 512, i.e. ONE TMEM
// accumulator stage): every epilogue team takes LNR contiguous columns of the row into registers with a single TMEM pass
// and hands the stage back to the MMA warp at once, so statistics, normalisation and the store of tile i run under the
// MMAs of tile i + 1 (the chunked epilogue below holds the stage for its whole duration: ~7.8k of 17.4k cycles per tile of
// the 288-channel head convolutions were exposed).
template <int EPI, bool PAIR, int LNR>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ TcMaps maps, const __grid_constant__ GemmArgs p, const __grid_constant__ TcSched sc) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment (SWIZZLE_128B atoms) without laundering the pointer through an integer, so
    // the compiler keeps the shared address space (LDS/STS instead of generic accesses)
    uint8_t *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int b_bytes = sc.BN * TBK * 2;
    const int b_stage = PAIR ? b_bytes / 2 : b_bytes;   // W bytes per ring stage of THIS CTA (pair mode: its half of the block)
    const int n_iters = p.taps * sc.kb_per_tap;
    uint8_t *smem_a = base;
    uint8_t *smem_b = base + sc.off_w;                  // W ring (stages blocks) or the resident W (n_iters blocks)
    float *lnx = reinterpret_cast<float *>(base + sc.off_lnx);
    const float *bias_s = reinterpret_cast<const float *>(base + sc.off_bias);
    const float *cs_s = reinterpret_cast<const float *>(base + sc.off_cs);
    const float *lnw_s = reinterpret_cast<const float *>(base + sc.off_lnw);
    const float *lnb_s = reinterpret_cast<const float *>(base + sc.off_lnb);
    uint64_t *full = reinterpret_cast<uint64_t *>(base + sc.off_bar);
    uint64_t *empty = full + MAX_STAGES;
    uint64_t *tmem_full = empty + MAX_STAGES;
    uint64_t *tmem_empty = tmem_full + MAX_ACC;
    uint64_t *w_full = tmem_empty + MAX_ACC;
    uint64_t *slab_full = w_full + MAX_WB;
    uint64_t *slab_empty = slab_full + N_TEAMS;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(slab_empty + N_TEAMS);

    constexpr bool G = EPI < 0;
    const bool f_ln = G ? (p.ln != 0) : ((EPI & 1) != 0);
    const int f_act = G ? p.act : ((EPI >> 1) & 3);
    const bool f_cs = G ? (p.colscale != nullptr) : (((EPI >> 3) & 1) != 0);
    const bool f_f32 = G ? (p.out_f32 != nullptr) : (((EPI >> 4) & 1) != 0);
    const bool f_b16 = G ? (p.out_act != nullptr) : (((EPI >> 5) & 1) != 0);
    const bool f_add = G ? (p.resid != nullptr || p.pe != nullptr) : (((EPI >> 6) & 1) != 0);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int crank = sc.cl > 1 ? (int)cluster_ctarank() : 0;
    const int cid = (int)blockIdx.x / sc.cl, ncl = (int)gridDim.x / sc.cl;       // cluster index / number of clusters
    const uint16_t cmask = (uint16_t)((1u << sc.cl) - 1u);
    const int bn_mma = sc.BN / sc.n_mma;
    const int nch = (sc.BN + 31) / 32;                  // 32-column chunks per tile

    if (warp == 0 && lane == 0) {
        for (int g = 0; g < sc.n_group; g++) {
            prefetch_tmap(&maps.a[g]); prefetch_tmap(&maps.w[g]);
            if (f_f32) prefetch_tmap(&maps.of[g]);
            if (f_b16) prefetch_tmap(&maps.ob[g]);
        }
        if (LNR > 0) prefetch_tmap(&maps.ob[1]);
        if (f_add) prefetch_tmap(&maps.add);
        for (int s = 0; s < sc.stages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], sc.pair ? 1 : sc.cl); }
        for (int s = 0; s < MAX_ACC; s++) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], sc.pair ? 2 * EPI_WARPS : EPI_WARPS); }
        for (int s = 0; s < MAX_WB; s++) mbar_init(&w_full[s], 1);
        for (int s = 0; s < N_TEAMS; s++) { mbar_init(&slab_full[s], 1); mbar_init(&slab_empty[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if constexpr (PAIR) {                           // one warp of EACH CTA of the pair
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    if (warp >= 2) {
        // per-column epilogue parameters -> shared memory (read back as broadcasts, thread = row)
        const int t = threadIdx.x - 64, nt = TC_THREADS - 64;
        float *bw = reinterpret_cast<float *>(base + sc.off_bias);
        if (p.bias) {
            for (int i = t; i < sc.n_group * p.N; i += nt) bw[i] = p.bias[(int64_t)(i / p.N) * p.g_stride_bias + i % p.N];
        } else {
            for (int i = t; i < sc.n_group * p.N; i += nt) bw[i] = 0.f;
        }
        if (f_cs) {
            float *cw = reinterpret_cast<float *>(base + sc.off_cs);
            for (int i = t; i < p.N; i += nt) cw[i] = p.colscale[i];
        }
        if (f_ln) {
            float *ww = reinterpret_cast<float *>(base + sc.off_lnw), *wb = reinterpret_cast<float *>(base + sc.off_lnb);
            for (int i = t; i < p.N; i += nt) { ww[i] = p.ln_w ? p.ln_w[i] : 1.f; wb[i] = p.ln_b ? p.ln_b[i] : 0.f; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (sc.cl > 1) cluster_sync_all();                  // peers' barriers are initialised before anything remote touches them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            // ------------------------------------------------ TMA producer (operands)
            const uint32_t tx_bytes = (uint32_t)(sc.w_res ? A_BYTES : A_BYTES + b_bytes);
            // single-thread role: no divisions in the loop (a dependent 32-bit division costs ~150 cycles)
            int s = 0, trn = 0;
            // stage release: k-block m (counted over the CTA's whole life) is covered by the commit issued after k-block
            // m - m % G + G - 1, which arrives on THAT k-block's stage barrier; `eph` bit t = parity of the next completion of
            // empty[t], flipped by the last member of a group (every completion is waited on by all G members)
            uint32_t eph = 0;
            int filled = 0, gr = 0;                     // k-blocks issued so far (saturating at stages), (m % G) of the stage being reused
            const int G = sc.commit_every;
            const int bn_sl = bn_mma / sc.cl;          // W rows of one multicast slice
            for (int item = cid; item < sc.items; item += ncl) {
                const TileIdx t = decode_tile(sc, item, crank);
                int seq_c, t0;
                tile_rows(sc, t.mt, t0, seq_c);
                const int n0 = t.nt * sc.BN;
                for (int tap = 0; tap < p.taps; tap++) {
                    const int shift = (tap - p.taps / 2) * p.dil;
                    for (int kb = 0; kb < sc.kb_per_tap; kb++) {
                        if (filled < sc.stages) {
                            filled++;                   // first pass over the ring: every stage is free
                        } else {
                            int tb = s + (G - 1 - gr);
                            if (tb >= sc.stages) tb -= sc.stages;
                            mbar_wait(&empty[tb], (eph >> tb) & 1u);
                            if (++gr == G) { gr = 0; eph ^= 1u << tb; }
                        }
                        trace_put(sc.trace, 0, trn);
                        if constexpr (PAIR) {
                            // CTA-pair mode: my 128 activation rows and my half of the W block land in MY shared memory, the
                            // bytes are counted on the LEADER's full[s] (it expects both CTAs' bytes and issues the MMAs)
                            const uint32_t lbar = mapa_rank(smem_u32(&full[s]), 0);
                            if (crank == 0) mbar_expect_tx(&full[s], (uint32_t)(((sc.debug & 4) ? 0 : 2 * A_BYTES) + ((sc.debug & 2) ? 0 : b_bytes)));
                            if (!(sc.debug & 4)) tma_load_3d_pair(&maps.a[t.g], lbar, smem_a + s * A_BYTES, kb * TBK, t0 + shift, seq_c);
                            for (int j = 0; j < sc.n_mma && !(sc.debug & 2); j++)
                                tma_load_3d_pair(&maps.w[t.g], lbar, smem_b + s * b_stage + j * bn_sl * TBK * 2, kb * TBK, tap,
                                                 n0 + j * bn_mma + crank * bn_sl);
                            if (++s == sc.stages) s = 0;
                            continue;
                        }
                        if (sc.w_res && item == cid) {
                            // weight-resident mode: this CTA's (group, n tile) never changes -> its W is loaded once, block by
                            // block in front of the first tile's activation blocks, each on its own barrier so that the MMAs
                            // start after the first W block instead of after the whole matrix (all CTAs pull their W from L2 at
                            // the same moment: a 19 MB burst at the start of every launch)
                            const int it = tap * sc.kb_per_tap + kb;
                            const int wb = it < MAX_WB - 1 ? it : MAX_WB - 1;
                            if (it <= MAX_WB - 1) mbar_expect_tx(&w_full[wb], (uint32_t)((it < MAX_WB - 1 ? 1 : n_iters - (MAX_WB - 1)) * b_bytes));
                            tma_load_3d(&maps.w[t.g], &w_full[wb], smem_b + it * b_bytes, kb * TBK, tap, n0);
                        }
                        mbar_expect_tx(&full[s], tx_bytes);
                        tma_load_3d(&maps.a[t.g], &full[s], smem_a + s * A_BYTES, kb * TBK, t0 + shift, seq_c);
                        if (!sc.w_res) {
                            if (sc.cl == 1) {
                                for (int j = 0; j < sc.n_mma; j++)
                                    tma_load_3d(&maps.w[t.g], &full[s], smem_b + s * b_bytes + j * bn_mma * TBK * 2, kb * TBK, tap,
                                                n0 + j * bn_mma);
                            } else {
                                // my 1/cl of every W box, delivered to the same stage of every CTA of the cluster (their
                                // full[s] barriers count the bytes; empty[s] has collected all cl consumers' releases)
                                for (int j = 0; j < sc.n_mma; j++)
                                    tma_load_3d_mc(&maps.w[t.g], &full[s],
                                                   smem_b + s * b_bytes + (j * bn_mma + crank * bn_sl) * TBK * 2, kb * TBK, tap,
                                                   n0 + j * bn_mma + crank * bn_sl, cmask);
                            }
                        }
                        if (++s == sc.stages) s = 0;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (!PAIR || crank == 0) {
            // ------------------------------------------------ MMA issuer (pair mode: the leader CTA issues for both SMs)
            // The WHOLE warp runs the loop and one elected lane issues the MMAs and the commit of a k-block: with uniform
            // control flow ptxas keeps the stage index, the barrier addresses and the operand descriptors in the uniform
            // datapath.  As a single-lane role the same loop took ~600 cycles of thread time per k-block (R2UR / VOTEU
            // chains, S2R + constant loads to rebuild shared-memory addresses, issue slots shared with four epilogue warps)
            // next to 8 x 71 cycles of tensor-pipe time, i.e. the issuing thread — not TMA, not the MMAs — paced the kernel.
            // instruction descriptor: D fp32, A/B bf16, both K-major, N = bn_mma, M = 128 (256 over a CTA pair)
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn_mma >> 3) << 17) |
                                   ((uint32_t)((sc.pair ? 2 * TBM : TBM) >> 4) << 24);
            const int n_stages = sc.stages, kb_per_tap = sc.kb_per_tap, k_tail = sc.k_tail_steps, n_mma = sc.n_mma, G = sc.commit_every;
            const bool skip_mma = (sc.debug & 1) != 0, w_res = sc.w_res != 0;
            // shared-memory descriptors (umma_desc_sw128): low word = 16-byte address | LBO, high word constant
            const uint64_t desc_hi = (uint64_t)(umma_desc_sw128(0) >> 32) << 32;
            const uint32_t a_lo0 = (uint32_t)umma_desc_sw128(smem_u32(smem_a)), b_lo0 = (uint32_t)umma_desc_sw128(smem_u32(smem_b));
            const uint32_t w_piece = (uint32_t)((PAIR ? bn_mma / 2 : bn_mma) * TBK * 2) >> 4;      // one MMA's W rows held by this CTA (16-byte units)
            const uint32_t b_step = (uint32_t)(w_res ? b_bytes : b_stage) >> 4;
            int s = 0, as = 0, trn = 0, cg = 0;          // cg: k-blocks since the last commit
            uint32_t ph = 0, aph = 0;
            for (int item = cid; item < sc.items; item += ncl) {
                mbar_wait(&tmem_empty[as], aph ^ 1u);
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(as * sc.acc_stride);
                const bool first = w_res && item == cid;
                int kb = 0;
                for (int it = 0; it < n_iters; it++) {
                    const int ks = (kb == kb_per_tap - 1) ? k_tail : TBK / 16;
                    if (++kb == kb_per_tap) kb = 0;
                    mbar_wait(&full[s], ph);
                    if (first) mbar_wait(&w_full[it < MAX_WB - 1 ? it : MAX_WB - 1], 0);   // first tile: W block `it` landed
                    tc_fence_after();
                    if (sc.trace != nullptr && lane == 0) trace_put(sc.trace, 1, trn);
                    const uint32_t a_lo = a_lo0 + (uint32_t)s * (A_BYTES >> 4);
                    const uint32_t b_lo = b_lo0 + (uint32_t)(w_res ? it : s) * b_step;
                    const bool commit = ++cg == G;
                    if (commit) cg = 0;
                    if (elect_one()) {
                        for (int j = 0; j < n_mma; j++) {
                            // pair mode: each CTA holds bn_mma / 2 rows of every W piece at the same shared-memory offset
#pragma unroll
                            for (int k = 0; k < TBK / 16; k++) {   // +32 B (= 2 x 16 B units) per K = 16 slice inside the swizzle row
                                if (k < ks && !skip_mma) {
                                    const uint64_t ad = desc_hi | (uint64_t)(a_lo + 2 * k), bd = desc_hi | (uint64_t)(b_lo + j * w_piece + 2 * k);
                                    if constexpr (PAIR) umma_bf16_pair(tacc + (uint32_t)(j * bn_mma), ad, bd, idesc, (it > 0 || k > 0) ? 1u : 0u);
                                    else umma_bf16(tacc + (uint32_t)(j * bn_mma), ad, bd, idesc, (it > 0 || k > 0) ? 1u : 0u);
                                }
                            }
                        }
                        if (commit) {                        // frees the group's smem stages once the MMAs above retire
                            if constexpr (PAIR) umma_commit_pair(&empty[s]);          // ... in BOTH CTAs
                            else if (sc.cl == 1) umma_commit(&empty[s]);
                            else umma_commit_mc(&empty[s], cmask);    // ... in every CTA of the cluster (their W slices land here)
                        }
                    }
                    __syncwarp();
                    if (++s == n_stages) { s = 0; ph ^= 1u; }
                }
                if (elect_one()) {
                    if constexpr (PAIR) umma_commit_pair(&tmem_full[as]);   // accumulator stage complete (both CTAs' epilogues)
                    else umma_commit(&tmem_full[as]);
                }
                __syncwarp();
                if (++as == sc.acc_stages) { as = 0; aph ^= 1u; }
            }
        }
    } else if (warp == 2) {
        if (f_add && elect_one()) {
            // ------------------------------------------------ C producer: fp32 addend chunks -> team slabs
            uint32_t eph = 0;                           // bit `team`: parity of that team's slab_empty barrier
            for (int item = cid; item < sc.items; item += ncl) {
                const TileIdx t = decode_tile(sc, item, crank);
                int seq_c, t0;
                tile_rows(sc, t.mt, t0, seq_c);
                if (sc.add_is_pe) seq_c = 0;
                const int n0 = t.nt * sc.BN;
                for (int c = 0; c < nch; c++) {
                    const int team = c & (N_TEAMS - 1);
                    mbar_wait(&slab_empty[team], ((eph >> team) & 1u) ^ 1u);
                    eph ^= 1u << team;
                    mbar_expect_tx(&slab_full[team], SLAB_F32);
                    tma_load_3d(&maps.add, &slab_full[team], base + sc.off_f32 + team * SLAB_F32, n0 + c * 32, t0, seq_c);
                }
            }
        }
    } else if (warp >= EPI_WARP0) {
        // ---------------------------------------------------- epilogue teams
        const int team = (warp - EPI_WARP0) >> 2;
        const int q = warp & 3;                         // TMEM lane quarter this warp may access
        const int r_tile = q * 32 + lane;               // this thread's row inside the tile
        const bool leader = (q == 0 && lane == 0);
        uint8_t *slab_f = base + sc.off_f32 + team * SLAB_F32;
        uint8_t *slab_b = base + sc.off_b16 + team * sc.nb16 * SLAB_B16;
        // swizzled row bases of this thread inside the slabs (TMA SWIZZLE_128B / SWIZZLE_64B patterns)
        uint8_t *rowf = slab_f + r_tile * 128;
        const int xf = r_tile & 7;                      // 16-byte chunk j of the row lives at (j ^ xf)
        const int xb = (r_tile >> 1) & 3;
        float *lxw = lnx + team * TBM + r_tile;         // [pass][team][row]
        const float *lx0 = lnx + r_tile;
        const int team_bar = 5 + team, q_bar = 1 + q;
        int as = 0, trn = 0, bbuf = 0;
        uint32_t aph = 0, sph = 0;
        for (int item = cid; item < sc.items; item += ncl) {
            const TileIdx ti = decode_tile(sc, item, crank);
            const int g = ti.g;
            const int n0 = ti.nt * sc.BN;
            int t0, seq_c;
            tile_rows(sc, ti.mt, t0, seq_c);
            float rm = 1.f;
            if (p.rowmask) {
                const int t = t0 + r_tile;
                const bool ok = sc.flat ? ((int64_t)t < (int64_t)p.n_seq * p.rows_per_seq) : (t < p.rows_per_seq);
                if (ok && ti.mt < sc.m_tiles) rm = (float)p.rowmask[(int64_t)seq_c * p.m_seq_stride + t];
            }
            const float *bias_t = bias_s + g * p.N + n0;

            if (team == 0 && leader) trace_put(sc.trace, 2, trn);       // tile setup done, waiting for the accumulator
            mbar_wait(&tmem_full[as], aph);
            tc_fence_after();
            if (team == 0 && leader) trace_put(sc.trace, 2, trn);       // accumulator ready
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * sc.acc_stride);

            if constexpr (LNR > 0) {
                float v[LNR];
                tmem_ld72(trow + (uint32_t)(team * LNR), v);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { if constexpr (PAIR) mbar_arrive_remote(mapa_rank(smem_u32(&tmem_empty[as]), 0)); else mbar_arrive(&tmem_empty[as]); }
                if (team == 0 && leader) trace_put(sc.trace, 2, trn);   // TMEM read done, stage released
                const int c0 = team * LNR;              // first column of this team inside the row (n0 == 0: one n tile)
                float2 s2 = make_float2(0.f, 0.f), q2 = make_float2(0.f, 0.f);
#pragma unroll
                for (int i = 0; i < LNR; i += 2) {
                    const float2 x = __fadd2_rn(make_float2(v[i], v[i + 1]), *reinterpret_cast<const float2 *>(bias_t + c0 + i));
                    v[i] = x.x; v[i + 1] = x.y;
                    s2 = __fadd2_rn(s2, x);
                    q2 = __ffma2_rn(x, x, q2);
                }
                lxw[0] = s2.x + s2.y;
                lxw[N_TEAMS * TBM] = q2.x + q2.y;
                named_barrier(q_bar, 128);
                const float *l1 = lx0 + N_TEAMS * TBM;
                const float inv_n = 1.0f / (float)p.N;
                const float mean = (lx0[0] + lx0[TBM] + lx0[2 * TBM] + lx0[3 * TBM]) * inv_n;
                const float var = fmaxf((l1[0] + l1[TBM] + l1[2 * TBM] + l1[3 * TBM]) * inv_n - mean * mean, 0.f);
                named_barrier(q_bar, 128);              // everyone has read the exchange buffer before the next tile rewrites it
                const float rstd = rsqrtf(var + p.ln_eps);
                const float2 rstd2 = make_float2(rstd, rstd), nmr2 = make_float2(-mean * rstd, -mean * rstd), rm2 = make_float2(rm, rm);
                // normalise -> affine -> ReLU -> mask -> bf16 of 8 consecutive columns starting at team column i0
                auto out8 = [&](int i0) -> uint4 {
                    uint4 pk;
                    __nv_bfloat162 *hp = reinterpret_cast<__nv_bfloat162 *>(&pk);
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const float4 w4 = *reinterpret_cast<const float4 *>(lnw_s + c0 + i0 + 4 * h);
                        const float4 c4 = *reinterpret_cast<const float4 *>(lnb_s + c0 + i0 + 4 * h);
                        float2 x0 = make_float2(v[i0 + 4 * h], v[i0 + 4 * h + 1]), x1 = make_float2(v[i0 + 4 * h + 2], v[i0 + 4 * h + 3]);
                        x0 = __ffma2_rn(__ffma2_rn(x0, rstd2, nmr2), make_float2(w4.x, w4.y), make_float2(c4.x, c4.y));
                        x1 = __ffma2_rn(__ffma2_rn(x1, rstd2, nmr2), make_float2(w4.z, w4.w), make_float2(c4.z, c4.w));
                        if (f_act == DECAF_ACT_RELU) { x0.x = fmaxf(x0.x, 0.f); x0.y = fmaxf(x0.y, 0.f); x1.x = fmaxf(x1.x, 0.f); x1.y = fmaxf(x1.y, 0.f); }
                        if (p.rowmask) { x0 = __fmul2_rn(x0, rm2); x1 = __fmul2_rn(x1, rm2); }
                        hp[2 * h] = __floats2bfloat162_rn(x0.x, x0.y);
                        hp[2 * h + 1] = __floats2bfloat162_rn(x1.x, x1.y);
                    }
                    return pk;
                };
                // The team's 72 columns leave in two TMA stores out of ONE 10 KB slab (a whole-row slab per team would cost the
                // operand ring its fifth stage, and the ring depth is what paces the MMAs of the streaming shapes):
                //   part 1: columns [0, P1) as dense 80-byte rows (16-byte stores of 8 consecutive rows hit 8 bank groups),
                //   part 2: columns [P1, LNR) as 64-byte SWIZZLE_64B rows, once part 1's store has read the slab.
                constexpr int P1 = LNR - 32;
                uint8_t *slab = base + sc.off_b16 + team * (TBM * P1 * 2);
                if (q == 0 && elect_one()) bulk_wait_read<0>();         // the previous tile's store has read the slab
                named_barrier(team_bar, 128);
                uint8_t *rowb = slab + r_tile * (P1 * 2);
#pragma unroll
                for (int j = 0; j < P1 / 8; j++) *reinterpret_cast<uint4 *>(rowb + 16 * j) = out8(8 * j);
                fence_proxy_async();
                named_barrier(team_bar, 128);
                if (q == 0 && elect_one()) {
                    tma_store_3d(&maps.ob[0], slab, c0, t0, seq_c);
                    bulk_commit();
                    bulk_wait_read<0>();
                }
                named_barrier(team_bar, 128);
                rowb = slab + r_tile * 64;
#pragma unroll
                for (int j = 0; j < 4; j++) *reinterpret_cast<uint4 *>(rowb + ((j ^ xb) << 4)) = out8(P1 + 8 * j);
                fence_proxy_async();
                named_barrier(team_bar, 128);
                if (q == 0 && elect_one()) {
                    tma_store_3d(&maps.ob[1], slab, c0 + P1, t0, seq_c);
                    bulk_commit();
                }
                if (team == 0 && leader) trace_put(sc.trace, 2, trn);   // tile stored
                if (++as == sc.acc_stages) { as = 0; aph ^= 1u; }
                continue;
            }

            float rstd = 1.f, nmr = 0.f;                // LN: y = (v + bias) * rstd + nmr,  nmr = -mean * rstd
            if (f_ln) {
                // statistics over this thread's row in ONE pass over TMEM (sum and sum of squares of x = acc + bias;
                // var = E[x^2] - mean^2 in fp32 over <= 512 O(1) values: ~1e-6 relative to the two-pass form of
                // libs/modeling/blocks.py:125-131, far below the bf16 rounding of this path's operands).  The epilogue of
                // the 288-channel head convs is not overlapped with MMAs (single TMEM stage), so a TMEM pass is ~8 % of
                // their tile time.  Partial sums over the team's chunks, exchanged between the four warps of the quarter.
                float s = 0.f, ss = 0.f;
                for (int c = team; c < nch; c += N_TEAMS) {
                    float v[32];
                    tmem_ld32(trow + (uint32_t)(c * 32), v);
                    const float *bs = bias_t + c * 32;
                    if (c * 32 + 32 <= p.N) {
                        float2 s2 = make_float2(0.f, 0.f), q2 = make_float2(0.f, 0.f);
#pragma unroll
                        for (int i = 0; i < 32; i += 2) {
                            const float2 x = __fadd2_rn(make_float2(v[i], v[i + 1]), *reinterpret_cast<const float2 *>(bs + i));
                            s2 = __fadd2_rn(s2, x);
                            q2 = __ffma2_rn(x, x, q2);
                        }
                        s += s2.x + s2.y;
                        ss += q2.x + q2.y;
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; i++) {
                            const float x = (c * 32 + i < p.N) ? v[i] + bs[i] : 0.f;
                            s += x;
                            ss = fmaf(x, x, ss);
                        }
                    }
                }
                lxw[0] = s;
                lxw[N_TEAMS * TBM] = ss;
                named_barrier(q_bar, 128);
                const float *l1 = lx0 + N_TEAMS * TBM;
                const float inv_n = 1.0f / (float)p.N;
                const float mean = (lx0[0] + lx0[TBM] + lx0[2 * TBM] + lx0[3 * TBM]) * inv_n;
                const float var = fmaxf((l1[0] + l1[TBM] + l1[2 * TBM] + l1[3 * TBM]) * inv_n - mean * mean, 0.f);
                named_barrier(q_bar, 128);              // everyone has read the exchange buffer before the next tile rewrites it
                rstd = rsqrtf(var + p.ln_eps);
                nmr = -mean * rstd;
            }

            if (team >= nch) {                          // narrow tiles: this team owns no chunk
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { if constexpr (PAIR) mbar_arrive_remote(mapa_rank(smem_u32(&tmem_empty[as]), 0)); else mbar_arrive(&tmem_empty[as]); }
            }
            for (int c = team; c < nch; c += N_TEAMS) {
                float v[32];
                tmem_ld32(trow + (uint32_t)(c * 32), v);
                if (c + N_TEAMS >= nch) {               // last TMEM read of this tile: hand the stage back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) { if constexpr (PAIR) mbar_arrive_remote(mapa_rank(smem_u32(&tmem_empty[as]), 0)); else mbar_arrive(&tmem_empty[as]); }
                }
                if (team == 0 && leader) trace_put(sc.trace, 2, trn);   // chunk: TMEM read done
                const int cn = c * 32;                  // first column of the chunk inside the tile
                // v = acc + bias; LN; act; * colscale   (per-column parameters: shared-memory broadcasts).  Packed fp32
                // (FADD2 / FMUL2 / FFMA2, two columns per instruction): the epilogue warps are issue-bound
                const float2 rstd2 = make_float2(rstd, rstd), nmr2 = make_float2(nmr, nmr);
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const float4 b4 = *reinterpret_cast<const float4 *>(bias_t + cn + 4 * j);
                    float2 x0 = __fadd2_rn(make_float2(v[4 * j], v[4 * j + 1]), make_float2(b4.x, b4.y));
                    float2 x1 = __fadd2_rn(make_float2(v[4 * j + 2], v[4 * j + 3]), make_float2(b4.z, b4.w));
                    if (f_ln) {
                        const float4 w4 = *reinterpret_cast<const float4 *>(lnw_s + n0 + cn + 4 * j);
                        const float4 c4 = *reinterpret_cast<const float4 *>(lnb_s + n0 + cn + 4 * j);
                        x0 = __ffma2_rn(__ffma2_rn(x0, rstd2, nmr2), make_float2(w4.x, w4.y), make_float2(c4.x, c4.y));
                        x1 = __ffma2_rn(__ffma2_rn(x1, rstd2, nmr2), make_float2(w4.z, w4.w), make_float2(c4.z, c4.w));
                    }
                    if (f_act == DECAF_ACT_RELU) {
                        x0.x = fmaxf(x0.x, 0.f); x0.y = fmaxf(x0.y, 0.f); x1.x = fmaxf(x1.x, 0.f); x1.y = fmaxf(x1.y, 0.f);
                    } else if (f_act == DECAF_ACT_GELU) {
                        if constexpr (kGeluTanhForBf16 && EPI >= 0 && ((EPI >> 4) & 1) == 0) {
                            x0 = gelu_tanh2(x0); x1 = gelu_tanh2(x1);
                        } else {
                            x0.x = gelu_erf_fast(x0.x); x0.y = gelu_erf_fast(x0.y); x1.x = gelu_erf_fast(x1.x); x1.y = gelu_erf_fast(x1.y);
                        }
                    }
                    if (f_cs) {
                        const float4 s4 = *reinterpret_cast<const float4 *>(cs_s + n0 + cn + 4 * j);
                        x0 = __fmul2_rn(x0, make_float2(s4.x, s4.y)); x1 = __fmul2_rn(x1, make_float2(s4.z, s4.w));
                    }
                    v[4 * j] = x0.x; v[4 * j + 1] = x0.y; v[4 * j + 2] = x1.x; v[4 * j + 3] = x1.y;
                }
                if (f_add) {
                    // the addend chunk was TMA-loaded into this team's fp32 slab by the C producer
                    mbar_wait(&slab_full[team], sph);
                    sph ^= 1u;
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const float4 a4 = *reinterpret_cast<const float4 *>(rowf + ((j ^ xf) << 4));
                        const float2 y0 = __fadd2_rn(make_float2(v[4 * j], v[4 * j + 1]), make_float2(a4.x, a4.y));
                        const float2 y1 = __fadd2_rn(make_float2(v[4 * j + 2], v[4 * j + 3]), make_float2(a4.z, a4.w));
                        v[4 * j] = y0.x; v[4 * j + 1] = y0.y; v[4 * j + 2] = y1.x; v[4 * j + 3] = y1.y;
                    }
                } else {
                    // slab reuse: the previous TMA store out of the buffer about to be overwritten must have been read
                    if (q == 0 && elect_one()) {           // (the team's store thread: elect.sync picks lane 0 of a full warp)
                        if (f_f32 || sc.nb16 == 1) bulk_wait_read<0>(); else bulk_wait_read<1>();
                    }
                    named_barrier(team_bar, 128);
                }
                if (p.rowmask) {
                    const float2 rm2 = make_float2(rm, rm);
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        const float2 y = __fmul2_rn(make_float2(v[i], v[i + 1]), rm2);
                        v[i] = y.x; v[i + 1] = y.y;
                    }
                }
                uint8_t *rowb = slab_b + bbuf * SLAB_B16 + r_tile * 64;
                if (f_f32) {
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        *reinterpret_cast<float4 *>(rowf + ((j ^ xf) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
                if (f_b16) {
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        uint4 pk;
                        __nv_bfloat162 *hp = reinterpret_cast<__nv_bfloat162 *>(&pk);
                        hp[0] = __floats2bfloat162_rn(v[8 * j], v[8 * j + 1]);
                        hp[1] = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]);
                        hp[2] = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]);
                        hp[3] = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
                        *reinterpret_cast<uint4 *>(rowb + ((j ^ xb) << 4)) = pk;
                    }
                }
                fence_proxy_async();                    // generic-proxy smem writes -> visible to the TMA store
                named_barrier(team_bar, 128);
                if (q == 0 && elect_one()) {
                    if (f_f32) tma_store_3d(&maps.of[g], slab_f, n0 + cn, t0, seq_c);
                    if (f_b16) tma_store_3d(&maps.ob[g], slab_b + bbuf * SLAB_B16, n0 + cn, t0, seq_c);
                    bulk_commit();
                    if (f_add) {                        // the slab goes back to the C producer once the store has read it
                        bulk_wait_read<0>();
                        mbar_arrive(&slab_empty[team]);
                    }
                }
                if (sc.nb16 == 2) bbuf ^= 1;
                if (team == 0 && leader) trace_put(sc.trace, 2, trn);   // chunk stored
            }
            if (++as == sc.acc_stages) { as = 0; aph ^= 1u; }
        }
        if (q == 0 && elect_one()) bulk_wait_all();     // all TMA stores of this thread are complete before exit
    }
    tc_fence_before();
    __syncthreads();
    if (sc.cl > 1) cluster_sync_all();                  // no CTA leaves while a peer may still multicast into it / arrive on it
    if (warp == 1) {
        tc_fence_after();
        if constexpr (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// ---------------------------------------------------------------------------------- host side
static inline int align_up(int x, int a) { return (x + a - 1) / a * a; }

const char *gemm_tc_why_not(const GemmArgs &a, int dtype) {
    static int sm100 = -1;
    if (sm100 < 0) sm100 = decaf_device_is_sm100();
    if (!sm100) return "device is not sm_100";
    if (dtype != DECAF_BF16) return "activation dtype is not bf16";
    if (a.K % 8 != 0) return "K is not a multiple of 8 (16-byte TMA pitch)";
    if (a.lda % 8 != 0) return "lda is not a multiple of 8";
    if (a.N % 8 != 0) return "N is not a multiple of 8";
    if ((reinterpret_cast<uintptr_t>(a.A) & 15) || (reinterpret_cast<uintptr_t>(a.W) & 15)) return "operands not 16-byte aligned";
    if ((a.a_seq_stride * a.lda) % 8 != 0) return "sequence pitch not 16-byte aligned";
    if ((int64_t)a.n_seq * a.rows_per_seq < 64) return "too few rows for a 128-row tensor-core tile";
    if (a.out_f32 && ((reinterpret_cast<uintptr_t>(a.out_f32) & 15) || a.ldo % 4 || a.g_stride_out_f32 % 4)) return "out_f32 not 16-byte aligned";
    if (a.out_act && ((reinterpret_cast<uintptr_t>(a.out_act) & 15) || a.ldo2 % 8 || a.g_stride_out_act % 8)) return "out_act not 16-byte aligned";
    if (a.resid && ((reinterpret_cast<uintptr_t>(a.resid) & 15) || a.ldr % 4)) return "resid not 16-byte aligned";
    if (a.resid && a.pe) return "resid and pe together (one fp32 addend per launch)";
    if (a.pe && (reinterpret_cast<uintptr_t>(a.pe) & 15)) return "pe not 16-byte aligned";
    if (a.N > MAX_PARAM_COLS) return "N > 2304 (per-column parameters are staged in shared memory)";
    if (a.ln && a.N > 512) return "fused LayerNorm needs N <= 512";
    if (get_encode() == nullptr) return "cuTensorMapEncodeTiled not available";
    return nullptr;
}

static unsigned long long *g_trace = nullptr;

// the one epilogue / width with a register-resident LayerNorm instantiation: conv -> LN -> ReLU -> bf16 over 4 x 72 = 288
// channels (the refined heads' towers, libs/modeling/head.py:33-44 on embd_dim + 32 channels)
constexpr int LNR_CODE = epi_code(true, DECAF_ACT_RELU, false, false, true, false);
constexpr int LNR_COLS = 72;

static constexpr bool code_has_pair(int code) {
    return code == epi_code(true, DECAF_ACT_RELU, false, false, true, false) || code == epi_code(true, DECAF_ACT_RELU, false, true, false, true) ||
           code == epi_code(false, DECAF_ACT_NONE, true, true, false, true) || code == epi_code(false, DECAF_ACT_NONE, true, true, true, true);
}

template <int CODE, bool PAIR, int LNR = 0>
static int launch_kernel(int grid, size_t smem, cudaStream_t st, const TcMaps &maps, const GemmArgs &a, const TcSched &sc) {
    static bool attr_set = false;
    if (!attr_set) {
        DECAF_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<CODE, PAIR, LNR>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        attr_set = true;
    }
    if (sc.cl == 1) {
        gemm_tc_kernel<CODE, PAIR, LNR><<<grid, TC_THREADS, smem, st>>>(maps, a, sc);
    } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = sc.cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        DECAF_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<CODE, PAIR, LNR>, maps, a, sc));
    }
    DECAF_LAUNCH_CHECK();
    return 0;
}

// streaming-mode shapes on the path (k=3 convs with fused LayerNorm, K = 1024 FFN proj2) have pair instantiations; any other
// epilogue falls back to the single-CTA kernel
template <int CODE>
static int launch_variant(int grid, size_t smem, cudaStream_t st, const TcMaps &maps, const GemmArgs &a, const TcSched &sc) {
    constexpr bool has_pair = code_has_pair(CODE);
    if constexpr (CODE == LNR_CODE) {
        if (sc.lnr == LNR_COLS) return sc.pair ? launch_kernel<CODE, true, LNR_COLS>(grid, smem, st, maps, a, sc)
                                                : launch_kernel<CODE, false, LNR_COLS>(grid, smem, st, maps, a, sc);
    }
    if constexpr (has_pair) {
        if (sc.pair) return launch_kernel<CODE, true>(grid, smem, st, maps, a, sc);
    }
    return launch_kernel<CODE, false>(grid, smem, st, maps, a, sc);
}

int gemm_tc_launch(const GemmArgs &a, int n_group, cudaStream_t st) {
    DECAF_CHECK(n_group <= MAX_GROUP, "decaf_gemm(tcgen05): at most %d groups", MAX_GROUP);
    DECAF_CHECK(!a.ln || n_group == 1, "decaf_gemm(tcgen05): fused LayerNorm does not support grouped launches");
    DECAF_CHECK((!a.resid && !a.pe && !a.colscale) || n_group == 1,
                "decaf_gemm(tcgen05): residual / PE / colscale epilogues do not support grouped launches");
    DECAF_CHECK((int64_t)n_group * a.N <= MAX_PARAM_COLS, "decaf_gemm(tcgen05): n_group * N > %d", MAX_PARAM_COLS);
    const bool f_f32 = a.out_f32 != nullptr, f_b16 = a.out_act != nullptr, f_add = a.resid != nullptr || a.pe != nullptr;
    const bool f_cs = a.colscale != nullptr, f_ln = a.ln != 0;
    const int64_t M = (int64_t)a.n_seq * a.rows_per_seq;
    DECAF_CHECK(M < (1ll << 31) - TBM, "decaf_gemm(tcgen05): too many rows");
    TcSched sc;
    memset(&sc, 0, sizeof(sc));
    // flat tiling (128-row tiles over all rows) needs every row-indexed tensor to be contiguous over sequences
    sc.flat = (a.taps == 1 && a.a_seq_stride == a.rows_per_seq && (!a.out_f32 || a.o_seq_stride == a.rows_per_seq) &&
               (!a.out_act || a.o2_seq_stride == a.rows_per_seq) && (!a.resid || a.r_seq_stride == a.rows_per_seq) &&
               (!a.rowmask || a.m_seq_stride == a.rows_per_seq) && !a.pe) ? 1 : 0;
    sc.tiles_per_seq = cdiv(a.rows_per_seq, TBM);
    sc.m_tiles = sc.flat ? cdiv(M, TBM) : a.n_seq * sc.tiles_per_seq;
    sc.kb_per_tap = cdiv(a.K, TBK);
    sc.k_tail_steps = cdiv(a.K - (sc.kb_per_tap - 1) * TBK, 16);
    sc.n_group = n_group;
    sc.add_is_pe = a.pe != nullptr;
    const int n_iters = a.taps * sc.kb_per_tap;

    // ---- shared-memory plan: [A ring][W ring | resident W][fp32 slabs][bf16 slabs][LN exchange][params][barriers]
    const int need_f32 = (f_f32 || f_add) ? 1 : 0;
    static int want_lnr = -1;
    if (want_lnr < 0) {
        const char *e = getenv("DECAF_GEMM_LNR");
        want_lnr = e ? atoi(e) : 1;
    }
    sc.lnr = (want_lnr && epi_code(f_ln, a.act, f_cs, f_f32, f_b16, f_add) == LNR_CODE && a.N == 4 * LNR_COLS) ? LNR_COLS : 0;
    const int slab_b16 = sc.lnr ? TBM * (sc.lnr - 32) * 2 : SLAB_B16;       // bytes of one team's bf16 slab
    const int params = align_up(n_group * a.N * 4, 16) + (f_cs ? align_up(a.N * 4, 16) : 0) + (f_ln ? 2 * align_up(a.N * 4, 16) : 0);
    const int fixed = 1024 + (f_ln ? LNX_BYTES : 0) + params + BAR_BYTES;
    const int slabs1 = need_f32 * N_TEAMS * SLAB_F32 + (f_b16 ? N_TEAMS * slab_b16 : 0);
    const int avail = SMEM_LIMIT - fixed - slabs1;             // for operands, with single-buffered bf16 slabs
    const int kpad = sc.kb_per_tap * TBK;
    sc.w_res = 0;
    if (f_ln) {
        // LN needs the full row in one CTA: N <= 512 split into n_mma instructions of <= 256 columns
        sc.n_mma = a.N <= 256 ? 1 : 2;
        sc.BN = align_up(a.N, 16 * sc.n_mma);
    } else {
        // prefer the widest BN (multiple of 32 = whole epilogue chunks, <= 256) whose (taps x K x BN) weight block fits
        // next to a 3-stage A ring ("weight resident": the TMA unit issues ~1 128-byte row per 2 cycles, so
        // re-loading a 256-row weight box per k-block costs twice the activation traffic); allow up to 2x more n
        // tiles than necessary (A is then re-read from L2); otherwise stream W through the ring with the widest BN
        // that leaves >= 3 stages.
        sc.n_mma = 1;
        const int n_min = cdiv(a.N, 256);
        for (int n_tiles = n_min; n_tiles <= 2 * n_min && !sc.w_res; n_tiles++) {
            const int bn = align_up(cdiv(a.N, n_tiles), 32);
            if (bn > 256) continue;
            if ((int64_t)a.taps * kpad * bn * 2 + 3 * A_BYTES <= avail && n_tiles * n_group <= num_sms()) { sc.BN = bn; sc.w_res = 1; }
        }
        if (!sc.w_res) {
            sc.BN = 0;
            for (int n_tiles = n_min; n_tiles <= 8 * n_min; n_tiles++) {
                const int bn = align_up(cdiv(a.N, n_tiles), 32);
                if (bn > 256) continue;
                if (3 * (A_BYTES + bn * TBK * 2) <= avail) { sc.BN = bn; break; }
            }
            DECAF_CHECK(sc.BN > 0, "decaf_gemm(tcgen05): no tile shape fits shared memory (N %d K %d)", a.N, a.K);
        }
    }
    sc.n_tiles = cdiv(a.N, sc.BN);
    // streaming mode, optional (DECAF_GEMM_CLUSTER=2|4): the CTAs of a thread-block cluster split every W block and
    // TMA-multicast their slices to each other; a slice must be whole 8-row swizzle atoms.  Measured on B200
    // (profiles/README.md): parity-green but no faster than cluster size 1 (head conv 66.5 vs 65.5 us, FFN proj2 48.0 vs
    // 46.9 us) and slower at 4 (fewer co-resident clusters) — the streaming shapes are not bound by L2 -> SM reads of W,
    // each SM still ingests the whole block.  Off by default; halving the ingest needs cta_group::2 MMAs.
    sc.cl = 1;
    if (!sc.w_res) {
        static int want = -1;
        if (want < 0) {
            const char *e = getenv("DECAF_GEMM_CLUSTER");
            want = e ? atoi(e) : 1;
            if (want != 1 && want != 2 && want != 4) want = 1;
        }
        int cl = want;
        while (cl > 1 && ((sc.BN / sc.n_mma) % (8 * cl) != 0 || num_sms() % cl != 0)) cl >>= 1;
        sc.cl = cl;
        // CTA-pair MMAs (cta_group::2): W halves of whole 8-row swizzle atoms, N of an instruction a multiple of 16
        static int want_pair = -1;
        if (want_pair < 0) {
            const char *e = getenv("DECAF_GEMM_PAIR");
            want_pair = e ? atoi(e) : 1;
        }
        if (want_pair && code_has_pair(epi_code(f_ln, a.act, f_cs, f_f32, f_b16, f_add)) && (sc.BN / sc.n_mma) % 16 == 0 && num_sms() % 2 == 0) {
            sc.cl = 2; sc.pair = 1;
        }
    }
    sc.acc_stride = sc.BN <= 128 ? 128 : (sc.BN <= 256 ? 256 : 512);
    sc.acc_stages = 512 / sc.acc_stride;
    const int combos = sc.n_tiles * n_group;
    sc.total_tiles = sc.m_tiles * combos;
    sc.items = cdiv(sc.m_tiles, sc.cl) * combos;
    const int b_bytes = sc.BN * TBK * 2;
    int op_bytes;
    if (sc.w_res) {
        const int w_bytes = n_iters * b_bytes;
        sc.stages = (avail - w_bytes) / A_BYTES;
        if (sc.stages > MAX_STAGES) sc.stages = MAX_STAGES;
        op_bytes = w_bytes + sc.stages * A_BYTES;
    } else {
        const int b_stage = sc.pair ? b_bytes / 2 : b_bytes;     // pair mode: a CTA holds half of every W block -> deeper ring
        sc.stages = avail / (A_BYTES + b_stage);
        if (sc.stages > MAX_STAGES) sc.stages = MAX_STAGES;
        op_bytes = sc.stages * (A_BYTES + b_stage);
    }
    DECAF_CHECK(sc.stages >= 2, "decaf_gemm(tcgen05): tile does not fit shared memory (BN %d)", sc.BN);
    static int want_commit = -1;
    if (want_commit < 0) {
        const char *e = getenv("DECAF_GEMM_COMMIT_EVERY");
        want_commit = e ? atoi(e) : 1;
        if (want_commit < 1) want_commit = 1;
    }
    sc.commit_every = want_commit < sc.stages - 1 ? want_commit : (sc.stages > 2 ? sc.stages - 2 : 1);   // >= 2 stages of look-ahead stay
    sc.nb16 = (f_b16 && !sc.lnr && avail - op_bytes >= N_TEAMS * SLAB_B16) ? 2 : 1;
    int off = sc.stages * A_BYTES;
    sc.off_w = off;            off = op_bytes;
    sc.off_f32 = off;          off += need_f32 * N_TEAMS * SLAB_F32;
    sc.off_b16 = off;          off += f_b16 ? N_TEAMS * sc.nb16 * slab_b16 : 0;
    sc.off_lnx = off;          off += f_ln ? LNX_BYTES : 0;
    sc.off_bias = off;         off += align_up(n_group * a.N * 4, 16);
    sc.off_cs = off;           off += f_cs ? align_up(a.N * 4, 16) : 0;
    sc.off_lnw = off;          off += f_ln ? align_up(a.N * 4, 16) : 0;
    sc.off_lnb = off;          off += f_ln ? align_up(a.N * 4, 16) : 0;
    sc.off_bar = off;          off += BAR_BYTES;
    const size_t smem = (size_t)off + 1024;
    DECAF_CHECK(smem <= SMEM_LIMIT, "decaf_gemm(tcgen05): shared-memory plan overflows (%zu bytes)", smem);
    sc.trace = g_trace;
    static int dbg = -1;
    if (dbg < 0) { dbg = getenv("DECAF_GEMM_DEBUG") ? 1 : 0; }
    { const char *e = getenv("DECAF_GEMM_DBG"); sc.debug = e ? atoi(e) : 0; }
    if (dbg)
        fprintf(stderr, "gemm_tc plan: M %lld K %d N %d taps %d groups %d | BN %d n_mma %d w_res %d pair %d cl %d stages %d acc_stages %d nb16 %d lnr %d commit_every %d smem %zu\n",
                (long long)M, a.K, a.N, a.taps, n_group, sc.BN, sc.n_mma, sc.w_res, sc.pair, sc.cl, sc.stages, sc.acc_stages, sc.nb16, sc.lnr, sc.commit_every, smem);

    const int bn_mma = sc.BN / sc.n_mma;
    TcMaps maps;
    const CUtensorMapDataType BF = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, FP = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    const CUtensorMapSwizzle S128 = CU_TENSOR_MAP_SWIZZLE_128B, S64 = CU_TENSOR_MAP_SWIZZLE_64B;
    const uint64_t rows_d1 = sc.flat ? (uint64_t)M : (uint64_t)a.rows_per_seq;
    const uint64_t seqs_d2 = sc.flat ? 1 : (uint64_t)a.n_seq;
    for (int g = 0; g < n_group; g++) {
        const bf16 *A = reinterpret_cast<const bf16 *>(a.A) + (int64_t)g * a.g_stride_a;
        const bf16 *W = reinterpret_cast<const bf16 *>(a.W) + (int64_t)g * a.g_stride_w;
        if (encode_3d(&maps.a[g], BF, S128, A, a.K, rows_d1, seqs_d2, a.lda * 2,
                      (uint64_t)(sc.flat ? M : a.a_seq_stride) * a.lda * 2, TBK, TBM, 1)) return 1;
        if (encode_3d(&maps.w[g], BF, S128, W, a.K, a.taps, a.N, (uint64_t)a.K * 2, (uint64_t)a.taps * a.K * 2, TBK, 1, bn_mma / sc.cl)) return 1;
        if (f_f32) {
            float *O = a.out_f32 + (int64_t)g * a.g_stride_out_f32;
            if (encode_3d(&maps.of[g], FP, S128, O, a.N, rows_d1, seqs_d2, a.ldo * 4,
                          (uint64_t)(sc.flat ? M : a.o_seq_stride) * a.ldo * 4, 32, TBM, 1)) return 1;
        }
        if (f_b16) {
            bf16 *O = reinterpret_cast<bf16 *>(a.out_act) + (int64_t)g * a.g_stride_out_act;
            if (sc.lnr) {                                // register-resident LN: [0] = the first lnr - 32 columns of a team (dense rows),
                if (encode_3d(&maps.ob[0], BF, CU_TENSOR_MAP_SWIZZLE_NONE, O, a.N, rows_d1, seqs_d2, a.ldo2 * 2,      // [1] = its last 32
                              (uint64_t)(sc.flat ? M : a.o2_seq_stride) * a.ldo2 * 2, sc.lnr - 32, TBM, 1)) return 1;
                if (encode_3d(&maps.ob[1], BF, S64, O, a.N, rows_d1, seqs_d2, a.ldo2 * 2,
                              (uint64_t)(sc.flat ? M : a.o2_seq_stride) * a.ldo2 * 2, 32, TBM, 1)) return 1;
                continue;
            }
            if (encode_3d(&maps.ob[g], BF, S64, O, a.N, rows_d1, seqs_d2, a.ldo2 * 2,
                          (uint64_t)(sc.flat ? M : a.o2_seq_stride) * a.ldo2 * 2, 32, TBM, 1)) return 1;
        }
    }
    if (a.resid) {
        if (encode_3d(&maps.add, FP, S128, a.resid, a.N, rows_d1, seqs_d2, a.ldr * 4,
                      (uint64_t)(sc.flat ? M : a.r_seq_stride) * a.ldr * 4, 32, TBM, 1)) return 1;
    } else if (a.pe) {
        if (encode_3d(&maps.add, FP, S128, a.pe, a.N, a.rows_per_seq, 1, (uint64_t)a.N * 4,
                      (uint64_t)a.rows_per_seq * a.N * 4, 32, TBM, 1)) return 1;
    }
    // persistent grid: one CTA per SM; a multiple of `combos` in weight-resident mode so that items
    // blockIdx.x + i * gridDim.x keep the CTA's (group, n tile); whole clusters in multicast mode
    int grid = sc.items < num_sms() / sc.cl ? sc.items * sc.cl : num_sms() / sc.cl * sc.cl;
    if (sc.w_res) {
        int per = num_sms() / combos;
        if (per > sc.m_tiles) per = sc.m_tiles;
        grid = per * combos;
    }
    const int code = epi_code(f_ln, a.act, f_cs, f_f32, f_b16, f_add);
#define TC_VARIANT(CODE)                                                                                          \
    if (code == (CODE)) return launch_variant<CODE>(grid, smem, st, maps, a, sc);
    TC_VARIANT(epi_code(false, DECAF_ACT_NONE, false, false, true, false))   // q/k/v, embd, AdaLN scale-shift projections
    TC_VARIANT(epi_code(false, DECAF_ACT_GELU, false, false, true, false))   // FFN fc
    TC_VARIANT(epi_code(false, DECAF_ACT_NONE, true, true, false, true))     // attention proj / FFN proj -> residual stream
    TC_VARIANT(epi_code(false, DECAF_ACT_NONE, true, true, true, true))      // FFN proj -> residual stream + FPN level copy
    TC_VARIANT(epi_code(false, DECAF_ACT_NONE, false, true, false, false))   // vid_map
    TC_VARIANT(epi_code(true, DECAF_ACT_RELU, false, false, true, false))    // conv -> LN -> ReLU (heads, embed convs)
    TC_VARIANT(epi_code(true, DECAF_ACT_RELU, false, true, false, true))     // last embed conv: + PE, fp32 residual stream
#undef TC_VARIANT
    return launch_variant<-1>(grid, smem, st, maps, a, sc);
}

}  // namespace decaf

// Debug hook (not part of the product path): the next tcgen05 GEMM launches write clock64 stamps of CTA 0 into
// buf[3][2048] (role 0 producer: stage acquired; 1 MMA: stage full; 2 epilogue team 0 leader: tile setup done /
// accumulator ready / per chunk: TMEM read done, chunk stored).  Pass NULL to switch it off.
extern "C" int decaf_debug_gemm_trace(unsigned long long *buf) {
    decaf::g_trace = buf;
    return 0;
}

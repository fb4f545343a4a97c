"""Time sharding of hour-long videos across GPUs (BASELINE.json configs[2]).

The reference evaluates a video on one GPU (libs/worker_v2.py:922-923) and explicitly does not support
sliding windows (:932-933).  An hour-long MAD video is ~70k clips x 64 queries — activations of tens of GB — so
here the padded timeline is split along time into `world` contiguous shards (aligned to 2^(L-1) steps so that
every FPN level splits at an integer index).  Each rank runs the grounder on its shard plus a HALO on both
sides, in one of two modes:

  * halo_mode='exchange' (default): a window-sized halo (exchange_halo: (1 + half window) rows of the COARSEST level, + 1
    row of margin = 1408 level-0 steps for 8 levels / window 19) that is REFRESHED from the neighbours after every encoder
    output: the rows next to a window edge that a k = 3 convolution / the local attention window computed from
    zero padding instead of real neighbours (1 + 9 rows at the level's own resolution) are overwritten with the
    neighbour's exact values — its OWN rows — by a grouped NCCL send/recv pair per level (or a copy between plans when
    several shards live in one process); the heads, the refinement TCN (dilations up to 2^(L-1)) and the pooling pyramid
    then fit inside the halo without further exchange.  Receptive-field sources: libs/modeling/blocks.py:357-373 (window),
    :462-473 (depthwise k3), libs/modeling/tcn.py:21-38 (dilations), libs/modeling/head.py:55-60 (k3 towers).
  * halo_mode='recompute': a one-shot halo covering the receptive field of the WHOLE network (receptive_halo: 3840 steps),
    no exchange inside the forward — 86 % extra work per shard at 8 shards of a MAD video, against 31.5 % for 'exchange'.

Either way every point a rank owns sees exactly the inputs it would see in the unsharded run, so candidates and final
segments are bit-identical to the unsharded path.  Three more things are global and need an exchange:

  1. the saliency top-k selection (libs/modeling/model.py:531-541) ranks blocks of the WHOLE valid length:
     every rank scores its own steps, the per-step scores are all-gathered (n x T floats), and every rank runs
     the (cheap, exact) selection kernel on the full row and keeps its window's columns;
  2. the absolute positional encoding is a function of the global index (video_net.py:143-152): each rank
     reads rows [w0, w0 + T_w) of the global table;
  3. the candidate top-k (libs/worker_v2.py:1169-1173) is over all points: each rank decodes the top-k of the
     points it OWNS (decaf_decode_window: global coordinates and global flat indices), the lists are
     all-gathered (ONE buffer of n x topk x 4 words + n counts per rank) and merged by decaf_merge_candidates into
     exactly the order the unsharded decode produces; NMS then runs on the merged list.

`TimeShardedEvaluator(evaluator, rank, world, group)` is the multi-process form (one process per GPU, NCCL over
NVLink); `emulate=S` runs S shards in a single process, advanced in lockstep through the exchange points (tests,
1-GPU parity checks against the unsharded path).
"""
import torch

from . import _cabi as cabi


def exchange_halo(opt, margin_rows=1):
    """Halo (level-0 steps, a multiple of 2^(L-1)) for halo_mode='exchange'.  With the FPN outputs refreshed after every
    encoder, what must fit in the halo AT EVERY LEVEL'S OWN RESOLUTION is (a) one encoder: depthwise k3 (1 row) + half
    window; (b) everything after the last exchange: first classification tower (n_layers + 1 k3 convs at each level),
    nearest expansion to level 0, the TCN (dilations 1 .. 2^(L-1): 2^L - 1 steps), the max-pool pyramid back down (<= 2
    rows per level) and the second towers (n_layers + 1 rows).  The coarsest level is the binding one."""
    m = opt['model']
    vn = m['vid_net']
    L = int(vn['arch'][2])
    top = 2 ** (L - 1)
    s = int(vn['mha_win_size']) // 2
    enc = 1 + s
    h1 = int(m['cls_head']['n_layers']) + 1
    h2 = max(int(m['cls_head']['n_layers']), int(m['reg_head']['n_layers'])) + 1
    tail_steps = h1 * top + (2 ** L - 1)                     # level-0 steps invalid after expansion + TCN
    tail = -(-tail_steps // top) + 2 + h2                    # rows of the coarsest level after pyramid + second towers
    return (max(enc, tail) + int(margin_rows)) * top


def halo_rows(shard, level, rows, T):
    """Row ranges (at `level`, window coordinates) of one halo exchange for a shard {'own': (a, e), 'win': (w0, w1)} of a
    timeline of T steps: ((recv_left, send_left), (recv_right, send_right)); a side is None at the ends of the timeline.
    recv_*: the outermost `rows` rows of that side's halo, overwritten with the neighbour's values; send_*: my OWN rows
    that are the outermost rows of the neighbour's halo on my side — the neighbour's window reaches as far into my range
    as mine reaches into its range (same halo h on every interior edge)."""
    (a, e), (w0, w1) = shard['own'], shard['win']
    n_l = (w1 - w0) >> level
    left = right = None
    if a > 0:                                                # interior left edge: my window starts at a - h
        h = a - w0
        assert h > 0 and (h >> level) >= rows, f'halo of {h} steps is narrower than {rows} rows at level {level}'
        s0 = ((a + h - w0) >> level) - rows                  # the left neighbour's window ends at a + h
        left = ((0, rows), (s0, s0 + rows))
    if e < T:                                                # interior right edge: my window ends at e + h
        h = w1 - e
        assert h > 0 and (h >> level) >= rows, f'halo of {h} steps is narrower than {rows} rows at level {level}'
        s0 = (e - h - w0) >> level                           # the right neighbour's window starts at e - h
        assert s0 >= ((a - w0) >> level), 'shard shorter than the halo'
        right = ((n_l - rows, n_l), (s0, s0 + rows))
    return left, right


def receptive_halo(opt):
    """Conservative receptive-field radius of the grounder in level-0 steps, rounded up to a multiple of
    2^(L-1).  (Measured on the reference by perturbing one input step: 3.3k / 3.6k steps for L = 8, window 19,
    SURVEY.md section 8(e); this bound gives 3840.)"""
    m = opt['model']
    vn = m['vid_net']
    L = int(vn['arch'][2])
    s = int(vn['mha_win_size']) // 2
    r = int(m['fusion']['n_layers'])                       # one depthwise k3 conv per fusion layer (blocks.py:513-516)
    r += int(vn['arch'][0])                                # embedding k3 convs (video_net.py:139-142)
    r += int(vn['arch'][1]) * (1 + s)                      # stem encoders at level 0
    for l in range(L):                                     # branch encoders: depthwise k3 (stride 2 from level 1) + window
        r += (1 if l == 0 else 2 ** (l - 1)) + s * 2 ** l
    top = 2 ** (L - 1)
    r += (int(m['cls_head']['n_layers']) + 1) * top        # first cls head: k3 convs at the coarsest level (head.py:55-60)
    r += top                                               # nearest expansion of level-l logits to level 0 (model.py:449-455)
    r += 2 ** L - 1                                        # TCN: dilations 1, 2, ..., 2^(L-1) (tcn.py:66-84)
    r += 2 ** L - 1                                        # max-pool pyramid back down (model.py:459-466)
    r += (max(int(m['cls_head']['n_layers']), int(m['reg_head']['n_layers'])) + 1) * top   # second heads
    return (r + top - 1) // top * top


def plan_shards(T, world, n_levels, halo):
    """Contiguous shards of the padded timeline [0, T): own = [s, e) (multiples of 2^(L-1)), window =
    [max(0, s - halo), min(T, e + halo))."""
    align = 2 ** (n_levels - 1)
    assert T % align == 0, f'T={T} is not a multiple of 2^(L-1)={align}'
    assert halo % align == 0
    units = T // align
    assert units >= world, f'{units} alignment units cannot be split over {world} ranks'
    base, extra = divmod(units, world)
    shards, s = [], 0
    for r in range(world):
        n = (base + (1 if r < extra else 0)) * align
        shards.append({'rank': r, 'own': (s, s + n), 'win': (max(0, s - halo), min(T, s + n + halo))})
        s += n
    return shards


def merge_candidates_reference(segs, scores, idx, count, topk):
    """torch statement of decaf_merge_candidates (tests; CPU or GPU tensors): inputs (n_src, n, topk[, 2])."""
    n_src, n = scores.shape[:2]
    out = []
    for q in range(n):
        ss, sc, ii = [], [], []
        for r in range(n_src):
            c = int(count[r, q])
            ss.append(segs[r, q, :c]); sc.append(scores[r, q, :c]); ii.append(idx[r, q, :c])
        ss, sc, ii = torch.cat(ss), torch.cat(sc), torch.cat(ii)
        o = torch.argsort(ii, stable=True)
        ss, sc, ii = ss[o], sc[o], ii[o]
        o = torch.argsort(sc, descending=True, stable=True)[:topk]
        out.append((ss[o], sc[o], ii[o]))
    return out


class _LocalComm:
    """All shards live in this process (emulation): "all-gather" = the list itself."""

    def __init__(self, world):
        self.world = world

    def all_gather(self, per_shard):
        return list(per_shard)


class _DistComm:
    """One shard per process; NCCL (or gloo) all-gather of equally shaped tensors and grouped neighbour send/recv."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def all_gather(self, per_shard):
        (t, ) = per_shard
        t = t.contiguous()
        if t.is_cuda:                                    # NCCL: one flat receive buffer
            out = torch.empty((self.world, ) + tuple(t.shape), dtype=t.dtype, device=t.device)
            self.dist.all_gather_into_tensor(out, t, group=self.group)
            return list(out.unbind(0))
        out = [torch.empty_like(t) for _ in range(self.world)]          # gloo (CPU tests of the host logic)
        self.dist.all_gather(out, t, group=self.group)
        return out

    def exchange(self, send_left, recv_left, send_right, recv_right):
        """One grouped send/recv with both neighbours (None = no neighbour on that side)."""
        dist, ops = self.dist, []
        peer = lambda r: dist.get_global_rank(self.group, r) if self.group is not None else r
        if send_left is not None:
            ops += [dist.P2POp(dist.isend, send_left, peer(self.rank - 1), self.group),
                    dist.P2POp(dist.irecv, recv_left, peer(self.rank - 1), self.group)]
        if send_right is not None:
            ops += [dist.P2POp(dist.isend, send_right, peer(self.rank + 1), self.group),
                    dist.P2POp(dist.irecv, recv_right, peer(self.rank + 1), self.group)]
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()


def assemble_rows(comm, shards, local_ids, own_chunks, T):
    """own_chunks[i]: (n, own length of shard local_ids[i]) -> the full (n, T) rows on every rank.  Chunks are
    padded to the longest shard so the all-gather is over equal shapes."""
    n = own_chunks[0].shape[0]
    mx = max(s['own'][1] - s['own'][0] for s in shards)
    padded = []
    for c in own_chunks:
        b = c.new_zeros(n, mx)
        b[:, :c.shape[1]] = c
        padded.append(b)
    gathered = comm.all_gather(padded)
    full = own_chunks[0].new_zeros(n, T)
    for s, g in zip(shards, gathered):
        a, e = s['own']
        full[:, a:e] = g[:, :e - a]
    return full


class TimeShardedEvaluator:
    def __init__(self, evaluator, rank=0, world=1, group=None, emulate=None, halo=None, halo_mode='exchange'):
        self.ev = evaluator
        self.eng = evaluator.model.engine()
        self.opt = evaluator.opt
        assert halo_mode in ('exchange', 'recompute')
        self.halo_mode = halo_mode
        if emulate is not None:
            self.world, self.local_ids, self.comm = int(emulate), list(range(int(emulate))), _LocalComm(int(emulate))
        else:
            self.world, self.local_ids = int(world), [int(rank)]
            self.comm = _DistComm(group) if world > 1 else _LocalComm(1)
        if halo is None:
            halo = exchange_halo(self.opt) if halo_mode == 'exchange' else receptive_halo(self.opt)
        self.halo = int(halo)
        self._buf = {}                       # persistent staging / exchange / result buffers, keyed by role and shape
        self.exchange_bytes = 0              # bytes this process sent in halo exchanges during the last predict_video

    def _cached(self, key, make):
        b = self._buf.get(key)
        if b is None:
            b = self._buf[key] = make()
        return b

    # ------------------------------------------------------------------ staging
    def _stage_window(self, data, T, w0, w1, slot=0):
        """Columns [w0, w1) of the video's features -> device, through pinned host buffers that persist across videos."""
        vid, shallow = data['vid'], data['shallow_vid']
        vid_len = vid.size(-1)
        Tw = w1 - w0
        out = []
        for name, src in (('vid', vid), ('sh', shallow)):
            C = src.size(0)
            d = self._cached(('d', name, slot, C, Tw), lambda: torch.zeros(C, Tw, device='cuda'))
            hi = min(w1, vid_len)
            k = max(hi - w0, 0)
            prev = self._buf.get(('len', name, slot, C, Tw), 0)          # columns of d that may be non-zero
            if src.is_pinned() and src.dtype == torch.float32 and src.stride(1) == 1:
                # features already in pinned host memory: the window's columns go up with ONE strided asynchronous copy
                if k < prev:
                    d[:, k:prev].zero_()
                if k:
                    cabi.upload_2d(d[:, :k], src[:, w0:hi])
            else:
                h = self._cached(('h', name, slot, C, Tw), lambda: torch.zeros(C, Tw).pin_memory())
                if k:
                    h[:, :k] = src[:, w0:hi]
                hp = self._buf.get(('hlen', name, slot, C, Tw), 0)
                if k < hp:
                    h[:, k:hp] = 0
                self._buf[('hlen', name, slot, C, Tw)] = k
                d.copy_(h, non_blocking=True)
            self._buf[('len', name, slot, C, Tw)] = k
            out.append(d)
        return out

    def _stage_text(self, data):
        ev = self.ev
        tokens = data['text']
        if not isinstance(tokens, tuple):
            tokens = (tokens, )
        n = len(tokens)
        Lm = max(t.size(-1) for t in tokens)
        Lmax = (Lm + ev.text_len_bucket - 1) // ev.text_len_bucket * ev.text_len_bucket
        padded = torch.nn.utils.rnn.pad_sequence([t.t() for t in tokens], batch_first=True)
        tok = torch.zeros(n, Lmax, tokens[0].size(0))
        tok[:, :Lm] = padded
        lens = torch.tensor([t.size(-1) for t in tokens], dtype=torch.int32)
        return tok.cuda(non_blocking=True), lens.cuda(non_blocking=True), data['text_cls'].float().cuda(non_blocking=True)

    # ------------------------------------------------------------------ halo exchange
    def _exchange_step(self, shards, mine, items, T):
        """items[i] = (level, X, cat, rows) yielded by shard mine[i]'s forward at the same point: refresh the outermost
        `rows` halo rows of X (fp32 residual stream / FPN output) from the neighbours and mirror them into cat (the bf16
        copy the heads read)."""
        level, _, _, rows = items[0]
        plans = [halo_rows(s, level, rows, T) for s in mine]
        if isinstance(self.comm, _LocalComm):
            if self.comm.world == 1:
                return
            by_rank = {s['rank']: (it, pl) for s, it, pl in zip(mine, items, plans)}
            for s, it, (left, right) in zip(mine, items, plans):
                X = it[1]
                if left is not None:
                    (r0, r1), _ = left
                    nb_X, nb_plan = by_rank[s['rank'] - 1][0][1], by_rank[s['rank'] - 1][1]
                    s0, s1 = nb_plan[1][1]                      # the left neighbour's send-right rows
                    X[:, r0:r1].copy_(nb_X[:, s0:s1])
                if right is not None:
                    (r0, r1), _ = right
                    nb_X, nb_plan = by_rank[s['rank'] + 1][0][1], by_rank[s['rank'] + 1][1]
                    s0, s1 = nb_plan[0][1]                      # the right neighbour's send-left rows
                    X[:, r0:r1].copy_(nb_X[:, s0:s1])
        else:
            (it, ), ((left, right), ) = items, plans
            X = it[1]
            B, _, C = X.shape
            mk = lambda tag: self._cached(('x', tag, level, B, rows, C), lambda: torch.empty(B, rows, C, device=X.device))
            sl = rl = sr = rr = None
            if left is not None:
                sl, rl = mk('sl'), mk('rl')
                sl.copy_(X[:, left[1][0]:left[1][1]])
            if right is not None:
                sr, rr = mk('sr'), mk('rr')
                sr.copy_(X[:, right[1][0]:right[1][1]])
            self.comm.exchange(sl, rl, sr, rr)
            if left is not None:
                X[:, left[0][0]:left[0][1]].copy_(rl)
                self.exchange_bytes += sl.numel() * 4
            if right is not None:
                X[:, right[0][0]:right[0][1]].copy_(rr)
                self.exchange_bytes += sr.numel() * 4
        for it, (left, right) in zip(items, plans):             # the heads' copy of the refreshed rows (same RN rounding
            X, cat = it[1], it[2]                               # as the GEMM epilogue that wrote the others)
            if cat is None:
                continue
            for side in (left, right):
                if side is not None:
                    r0, r1 = side[0]
                    cat[:, r0:r1].copy_(X[:, r0:r1])

    # ------------------------------------------------------------------ predict
    @torch.no_grad()
    def predict_video(self, data, return_candidates=False):
        ev, eng, opt = self.ev, self.eng, self.opt
        vid_len = data['vid'].size(-1)
        T = ev.padded_len(vid_len)
        shards = plan_shards(T, self.world, eng.L, self.halo)
        mine = [shards[i] for i in self.local_ids]
        exchange = self.halo_mode == 'exchange' and self.world > 1
        if exchange:
            for s in shards:
                assert s['own'][1] - s['own'][0] >= self.halo, f'shards of {s["own"][1] - s["own"][0]} steps are shorter than the {self.halo}-step halo'
        self.exchange_bytes = 0
        tok, lens, text_cls = self._stage_text(data)
        n = tok.size(0)
        eng.lane = 0
        text, kv_len, kv = eng.encode_text_batch(tok, lens)                   # replicated: n x <= 32 rows
        dev = text.device
        # (1) per-step saliency of the owned steps -> all-gather -> global selection on every rank
        wins, own_chunks = [], []
        for i, s in enumerate(mine):
            w0, w1 = s['win']
            dv, ds = self._stage_window(data, T, w0, w1, slot=i)
            corr = self._cached(('corr', i, n, w1 - w0), lambda: torch.empty(n, w1 - w0, device=dev))
            cabi.saliency(ds, text_cls, corr, ds.shape[0], w1 - w0, n, eng.norm)
            wins.append((dv, ds, corr))
            own_chunks.append(corr[:, s['own'][0] - w0:s['own'][1] - w0].contiguous())
        correl = assemble_rows(self.comm, shards, self.local_ids, own_chunks, T)
        vm = self._cached(('vm', T), lambda: torch.empty(T, dtype=torch.uint8, device=dev))
        vm.copy_((torch.arange(T, device=dev) < vid_len))
        sel = self._cached(('sel', n, T), lambda: torch.empty(n, T, dtype=torch.uint8, device=dev))
        mask0 = self._cached(('mask0', n, T), lambda: torch.empty(n, T, dtype=torch.uint8, device=dev))
        max_blocks = (T + eng.sn - 1) // eng.sn
        pooled = self._cached(('pooled', n, max_blocks), lambda: torch.empty(n, max_blocks, device=dev))
        cabi.select(correl, vm, sel, mask0, pooled, max_blocks, T, n, eng.sn, eng.sratio, and_mask=not eng.msf)
        # (2) the grounder on every window, with the global selection / PE rows; with halo exchange the local shards advance
        # in lockstep through the exchange points (one per encoder output)
        ev_opt = opt['eval']
        topk = int(ev_opt['pre_nms_topk'])
        gens, plans = [], []
        for i, (s, (dv, ds, corr)) in enumerate(zip(mine, wins)):
            w0, w1 = s['win']
            eng.lane = 100 + i if len(mine) > 1 else 0                    # one workspace set per local shard
            p = eng.plan(n, w1 - w0)
            p.correl.copy_(correl[:, w0:w1])
            p.sel.copy_(sel[:, w0:w1])
            p.mask0.copy_(mask0[:, w0:w1])
            g = eng.forward_steps(dv, ds, vm[w0:w1], text, kv_len, text_cls, text_kv=kv, window=(T, w0), halo_steps=exchange)
            gens.append(g)
            plans.append(p)
        done = [False] * len(gens)
        while not all(done):
            items = []
            for i, g in enumerate(gens):
                eng.lane = 100 + i if len(mine) > 1 else 0
                try:
                    items.append(next(g))
                except StopIteration as e:
                    done[i] = True
                    assert e.value is plans[i]
            if items:
                assert len(items) == len(gens), 'shards left the exchange points out of step'
                self._exchange_step(shards, mine, items, T)
        eng.lane = 0
        # (3) per-shard candidates of the OWNED points -> one all-gather -> merge into the global top-k -> NMS
        packs = []
        for i, (s, p) in enumerate(zip(mine, plans)):
            w0 = s['win'][0]
            pk = self._cached(('pack', i, n, topk), lambda: torch.zeros(n * topk * 4 + n, device=dev))
            segs = pk[:n * topk * 2].view(n, topk, 2)
            scores = pk[n * topk * 2:n * topk * 3].view(n, topk)
            idx = pk[n * topk * 3:n * topk * 4].view(torch.int32).view(n, topk)
            cnt = pk[n * topk * 4:].view(torch.int32)
            cabi.decode_window(p.logits2, p.offsets, p.hmask, p.lv, n, True, float(ev_opt['pre_nms_thresh']), topk,
                               float(ev_opt['seg_len_thresh']), w0, s['own'][0] - w0, s['own'][1] - w0, T, segs, scores, idx, cnt)
            packs.append(pk)
        g_all = torch.stack(self.comm.all_gather(packs)).contiguous()          # (world, n * topk * 4 + n)
        W_ = g_all.shape[0]
        g_segs = g_all[:, :n * topk * 2].reshape(W_, n, topk, 2).contiguous()
        g_scores = g_all[:, n * topk * 2:n * topk * 3].reshape(W_, n, topk).contiguous()
        g_idx = g_all[:, n * topk * 3:n * topk * 4].contiguous().view(torch.int32).view(W_, n, topk)
        g_cnt = g_all[:, n * topk * 4:].contiguous().view(torch.int32).view(W_, n)
        m_segs = self._cached(('m_segs', n, topk), lambda: torch.zeros(n, topk, 2, device=dev))
        m_scores = self._cached(('m_scores', n, topk), lambda: torch.zeros(n, topk, device=dev))
        m_idx = self._cached(('m_idx', n, topk), lambda: torch.zeros(n, topk, dtype=torch.int32, device=dev))
        m_cnt = self._cached(('m_cnt', n), lambda: torch.zeros(n, dtype=torch.int32, device=dev))
        cabi.merge_candidates(g_segs, g_scores, g_idx, g_cnt, self.world, n, topk, m_segs, m_scores, m_idx, m_cnt)
        nm = opt['nms']
        prm = cabi.NmsParams()
        prm.mode = {None: 0, 'nms': 1, 'soft_nms': 2}[nm['mode']]
        prm.iou_thresh, prm.sigma, prm.min_score = float(nm['iou_thresh']), float(nm['sigma']), float(nm['min_score'])
        prm.max_num_segs, prm.voting_thresh = int(nm['max_num_segs']), float(nm['voting_thresh'])
        prm.to_seconds = 1
        prm.vid_stride = float(ev.vid_stride)
        prm.clip_stride, prm.half_clip_size = float(data['clip_stride']), float(0.5 * data['clip_size'])
        prm.fps, prm.duration = float(data['fps']), float(data['duration'])
        max_out = prm.max_num_segs if prm.max_num_segs > 0 else topk
        o_segs = self._cached(('o_segs', n, max_out), lambda: torch.zeros(n, max_out, 2, device=dev))
        o_scores = self._cached(('o_scores', n, max_out), lambda: torch.zeros(n, max_out, device=dev))
        o_cnt = self._cached(('o_cnt', n), lambda: torch.zeros(n, dtype=torch.int32, device=dev))
        ws = self._cached(('nms_ws', n, topk), lambda: torch.empty(int(cabi.nms_workspace_bytes(n, topk)), dtype=torch.uint8, device=dev))
        cabi.batched_nms(m_segs, m_scores, m_cnt, n, topk, prm, o_segs, o_scores, o_cnt, ws)
        h_segs, h_scores, h_cnt = o_segs.cpu(), o_scores.cpu(), o_cnt.cpu()
        results = [{'segments': h_segs[b, :int(h_cnt[b])].clone(), 'scores': h_scores[b, :int(h_cnt[b])].clone()} for b in range(n)]
        if return_candidates:
            return results, (m_segs, m_scores, m_idx, m_cnt)
        return results

"""Time sharding of hour-long videos across GPUs (BASELINE.json configs[2]).

The reference evaluates a video on one GPU (libs/worker_v2.py:922-923) and explicitly does not support
sliding windows (:932-933).  An hour-long MAD video is ~70k clips x 64 queries — activations of tens of GB — so
here the padded timeline is split along time into `world` contiguous shards (aligned to 2^(L-1) steps so that
every FPN level splits at an integer index); each rank runs the whole grounder on its shard plus a
recompute HALO on both sides that covers the receptive field of the network, so every point a rank owns sees
exactly the inputs it would see in the unsharded run.  Three things are global and need an exchange:

  1. the saliency top-k selection (libs/modeling/model.py:531-541) ranks blocks of the WHOLE valid length:
     every rank scores its own steps, the per-step scores are all-gathered (n x T floats), and every rank runs
     the (cheap, exact) selection kernel on the full row and keeps its window's columns;
  2. the absolute positional encoding is a function of the global index (video_net.py:143-152): each rank
     reads rows [w0, w0 + T_w) of the global table;
  3. the candidate top-k (libs/worker_v2.py:1169-1173) is over all points: each rank decodes the top-k of the
     points it OWNS (decaf_decode_window: global coordinates and global flat indices), the lists are
     all-gathered (n x topk x 4 words per rank) and merged by decaf_merge_candidates into exactly the order
     the unsharded decode produces; NMS then runs on the merged list.

`TimeShardedEvaluator(evaluator, rank, world, group)` is the multi-process form (one process per GPU, NCCL
all-gathers over NVLink); `emulate=S` runs S shards one after the other in a single process (tests, 1-GPU
parity checks against the unsharded path).
"""
import torch

from . import _cabi as cabi


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
    """One shard per process; NCCL (or gloo) all-gather of equally shaped tensors."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world = dist.get_world_size(group)

    def all_gather(self, per_shard):
        (t, ) = per_shard
        out = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t.contiguous(), group=self.group)
        return out


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
    def __init__(self, evaluator, rank=0, world=1, group=None, emulate=None, halo=None):
        self.ev = evaluator
        self.eng = evaluator.model.engine()
        self.opt = evaluator.opt
        if emulate is not None:
            self.world, self.local_ids, self.comm = int(emulate), list(range(int(emulate))), _LocalComm(int(emulate))
        else:
            self.world, self.local_ids = int(world), [int(rank)]
            self.comm = _DistComm(group) if world > 1 else _LocalComm(1)
        self.halo = receptive_halo(self.opt) if halo is None else int(halo)
        self._buf = {}

    # ------------------------------------------------------------------ staging
    def _stage_window(self, data, T, w0, w1):
        vid, shallow = data['vid'], data['shallow_vid']
        vid_len = vid.size(-1)
        Tw = w1 - w0
        hv = torch.zeros(vid.size(0), Tw).pin_memory()
        hs = torch.zeros(shallow.size(0), Tw).pin_memory()
        hi = min(w1, vid_len)
        if hi > w0:
            hv[:, :hi - w0] = vid[:, w0:hi]
            hs[:, :hi - w0] = shallow[:, w0:hi]
        return hv.cuda(non_blocking=True), hs.cuda(non_blocking=True)

    def _stage_text(self, data):
        ev = self.ev
        tokens = data['text']
        if not isinstance(tokens, tuple):
            tokens = (tokens, )
        n = len(tokens)
        Lmax = max(t.size(-1) for t in tokens)
        Lmax = (Lmax + ev.text_len_bucket - 1) // ev.text_len_bucket * ev.text_len_bucket
        tok = torch.zeros(n, Lmax, tokens[0].size(0))
        lens = torch.zeros(n, dtype=torch.int32)
        for i, t in enumerate(tokens):
            tok[i, :t.size(-1)] = t.t()
            lens[i] = t.size(-1)
        return tok.cuda(), lens.cuda(), data['text_cls'].float().cuda()

    # ------------------------------------------------------------------ predict
    @torch.no_grad()
    def predict_video(self, data, return_candidates=False):
        ev, eng, opt = self.ev, self.eng, self.opt
        vid_len = data['vid'].size(-1)
        T = ev.padded_len(vid_len)
        shards = plan_shards(T, self.world, eng.L, self.halo)
        mine = [shards[i] for i in self.local_ids]
        tok, lens, text_cls = self._stage_text(data)
        n = tok.size(0)
        text, kv_len, kv = eng.encode_text_batch(tok, lens)                   # replicated: n x <= 32 rows
        dev = text.device
        # (1) per-step saliency of the owned steps -> all-gather -> global selection on every rank
        wins, own_chunks = [], []
        for s in mine:
            w0, w1 = s['win']
            dv, ds = self._stage_window(data, T, w0, w1)
            corr = torch.empty(n, w1 - w0, device=dev)
            cabi.saliency(ds, text_cls, corr, ds.shape[0], w1 - w0, n, eng.norm)
            wins.append((dv, ds, corr))
            own_chunks.append(corr[:, s['own'][0] - w0:s['own'][1] - w0].contiguous())
        correl = assemble_rows(self.comm, shards, self.local_ids, own_chunks, T)
        vm = (torch.arange(T, device=dev) < vid_len).to(torch.uint8)
        sel = torch.empty(n, T, dtype=torch.uint8, device=dev)
        mask0 = torch.empty(n, T, dtype=torch.uint8, device=dev)
        max_blocks = (T + eng.sn - 1) // eng.sn
        pooled = torch.empty(n, max_blocks, device=dev)
        cabi.select(correl, vm, sel, mask0, pooled, max_blocks, T, n, eng.sn, eng.sratio, and_mask=not eng.msf)
        # (2) the grounder on every window, with the global selection / PE rows
        ev_opt = opt['eval']
        topk = int(ev_opt['pre_nms_topk'])
        cs, csc, cid, ccnt = [], [], [], []
        for s, (dv, ds, corr) in zip(mine, wins):
            w0, w1 = s['win']
            p = eng.plan(n, w1 - w0)
            p.correl.copy_(correl[:, w0:w1])
            p.sel.copy_(sel[:, w0:w1])
            p.mask0.copy_(mask0[:, w0:w1])
            eng.forward(dv, ds, vm[w0:w1], text, kv_len, text_cls, text_kv=kv, window=(T, w0))
            segs = torch.zeros(n, topk, 2, device=dev)
            scores = torch.zeros(n, topk, device=dev)
            idx = torch.zeros(n, topk, dtype=torch.int32, device=dev)
            cnt = torch.zeros(n, dtype=torch.int32, device=dev)
            cabi.decode_window(p.logits2, p.offsets, p.hmask, p.lv, n, True, float(ev_opt['pre_nms_thresh']), topk,
                               float(ev_opt['seg_len_thresh']), w0, s['own'][0] - w0, s['own'][1] - w0, T, segs, scores, idx, cnt)
            cs.append(segs); csc.append(scores); cid.append(idx); ccnt.append(cnt)
        # (3) all-gather the per-shard candidates, merge into the global top-k, NMS
        g_segs = torch.stack(self.comm.all_gather(cs)).contiguous()
        g_scores = torch.stack(self.comm.all_gather(csc)).contiguous()
        g_idx = torch.stack(self.comm.all_gather(cid)).contiguous()
        g_cnt = torch.stack(self.comm.all_gather(ccnt)).contiguous()
        m_segs = torch.zeros(n, topk, 2, device=dev)
        m_scores = torch.zeros(n, topk, device=dev)
        m_idx = torch.zeros(n, topk, dtype=torch.int32, device=dev)
        m_cnt = torch.zeros(n, dtype=torch.int32, device=dev)
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
        o_segs = torch.zeros(n, max_out, 2, device=dev)
        o_scores = torch.zeros(n, max_out, device=dev)
        o_cnt = torch.zeros(n, dtype=torch.int32, device=dev)
        ws = torch.empty(int(cabi.nms_workspace_bytes(n, topk)), dtype=torch.uint8, device=dev)
        cabi.batched_nms(m_segs, m_scores, m_cnt, n, topk, prm, o_segs, o_scores, o_cnt, ws)
        o_segs, o_scores, o_cnt = o_segs.cpu(), o_scores.cpu(), o_cnt.cpu()
        results = [{'segments': o_segs[b, :int(o_cnt[b])], 'scores': o_scores[b, :int(o_cnt[b])]} for b in range(n)]
        if return_candidates:
            return results, (m_segs, m_scores, m_idx, m_cnt)
        return results

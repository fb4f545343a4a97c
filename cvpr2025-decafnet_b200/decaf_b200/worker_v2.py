"""Mirror of the evaluation half of libs/worker_v2.py (Evaluator, create_model) on top of the
sm_100a kernels.  Same method names, attributes and result format:

    evaluator = Evaluator(opt, dataset=items)          # items: dicts with the schema of
    evaluator.run()                                    #   libs/data/dataset.py:977-994
    outputs, results, loss = evaluator.simple_predict(item)
    results[i] == {'segments': (k, 2) float32 CPU seconds, 'scores': (k,)}

What differs from the reference loop (and why it is faster): all queries of a video are encoded,
fused, decoded and NMS-ed in one batch instead of one at a time (libs/worker_v2.py:940-955,
1077-1124, libs/modeling/model.py:526-563); inputs are staged through pinned host buffers; there
is one device->host copy per video (the <= max_num_segs final segments) instead of one per
query before a CPU NMS.  The reference's feature-file datasets, logger, W&B and training loop
are out of scope.
"""
from collections import OrderedDict, defaultdict
import os
import time

import numpy as np
import torch

from . import _cabi as cabi
from .modeling import PtGenerator, PtTransformerEarlyFusionIterative
from .nms import batched_nms


def create_model(opt, act_dtype=torch.bfloat16, gemm_impl=0):
    """libs/worker_v2.py:182-211: only opt.model.name == 'iter' is live in the reference."""
    if opt.model.name == 'iter':
        return PtTransformerEarlyFusionIterative(opt, second_fusion=False, act_dtype=act_dtype, gemm_impl=gemm_impl)
    raise NotImplementedError(f"model name {opt.model.name!r}: the reference only builds 'iter'")


def iou(pred_segs, gt_segs):
    """libs/train_utils.py:81-96."""
    ps, pe = pred_segs[..., 0], pred_segs[..., 1]
    gs, ge = gt_segs[..., 0], gt_segs[..., 1]
    overlap = (torch.minimum(pe, ge) - torch.maximum(ps, gs)).clamp(min=0)
    union = (pe - ps) + (ge - gs) - overlap
    return overlap / union


def _carve(buf, specs):
    """Typed views into one uint8 buffer: specs = [(name, shape, dtype)], every view 256-byte aligned.  Returns
    ({name: view}, bytes used); call with buf=None to size the buffer."""
    out, off = {}, 0
    for name, shape, dtype in specs:
        nbytes = int(np.prod(shape)) * torch.empty(0, dtype=dtype).element_size()
        if buf is not None:
            out[name] = buf[off:off + nbytes].view(dtype).view(*shape)
        off = (off + nbytes + 255) // 256 * 256
    return out, off


class Evaluator:
    """libs/worker_v2.py:726-1227 (evaluation path)."""

    def __init__(self, opt, train_time=False, dataset=None, model=None, state_dict=None,
                 act_dtype=torch.bfloat16, gemm_impl=0, logger=None, use_graphs=True, text_len_bucket=4, n_lanes=8,
                 max_cached_shapes=8, gemm_sms=None):
        self.opt = opt
        if dataset is None:
            raise ValueError(
                'pass dataset=<iterable of items with the schema of libs/data/dataset.py:977-994>; the '
                "reference's feature-file loaders are outside the hot path (SURVEY.md section 2, row 17)")
        self.dataset = dataset
        self.dataloader = dataset
        self.num_itrs = len(dataset) if hasattr(dataset, '__len__') else None
        self.itr = self.text_cnt = 0

        if model is not None:
            self.model = model
        elif not train_time:
            self.model = create_model(opt, act_dtype=act_dtype, gemm_impl=gemm_impl).cuda()
            if state_dict is not None:
                self.model.load_state_dict(state_dict)
            else:
                self.load_model()
            self.model.eval().requires_grad_(False)
        else:
            self.model = None
        pt_gen = opt.pt_gen.clone()
        pt_gen.max_seq_len = opt.model.vid_net.max_seq_len * 10         # libs/worker_v2.py:752-754
        self.pt_gen = PtGenerator(**pt_gen).cuda()
        self.logger = logger

        self.max_vid_len = opt['model']['max_vid_len']
        self.vid_stride = opt['model'].get('vid_stride', 1)
        self.input_vid_len = self.max_vid_len * self.vid_stride
        num_fpn_levels = opt['model']['num_fpn_levels']
        mha_win_size = opt['model']['mha_win_size']
        min_chunk_size = 1
        for idx in range(num_fpn_levels):
            stride = 2 ** idx
            if mha_win_size > 0:
                stride *= (mha_win_size // 2) * 2
            min_chunk_size = max(min_chunk_size, stride)
        assert self.max_vid_len % min_chunk_size == 0, (
            f"max video length must be a multiple of {min_chunk_size}")
        self.min_chunk_size = min_chunk_size

        self.ranks = opt['eval'].get('ranks', (1, 5))
        self.topk = max(self.ranks)
        self.iou_threshs = np.array(opt['eval'].get('iou_threshs', (0.3, 0.5)))
        self.counts = np.zeros((len(self.ranks), len(self.iou_threshs)))
        self.window_size = opt['eval'].get('window_size')
        self.window_stride = opt['eval'].get('window_stride')
        self.batched_nms = lambda segs, scores: batched_nms(segs, scores, **opt['nms'])
        self.pre_nms_topk = opt['eval']['pre_nms_topk']
        self.pre_nms_thresh = opt['eval']['pre_nms_thresh']
        self.seg_len_thresh = opt['eval']['seg_len_thresh']
        self.time_dict = defaultdict(list)
        self._stage = {}
        # the ~150 launches of one video are captured once per (T, n_query, Lmax) into a CUDA graph and replayed:
        # the kernels of the bf16 path are short enough that per-launch host cost would otherwise dominate
        self.use_graphs = use_graphs
        self.text_len_bucket = max(1, int(text_len_bucket))    # Lmax is rounded up to this (padding keys are masked)
        self._graphs = {}
        # videos in flight (predict_videos / run): each lane owns a stream, pinned staging buffers, engine workspaces
        # and CUDA graphs, so the host staging + H2D of video i+1 and the latency-bound phases of video i (text
        # encoder, top FPN levels, TCN: a few CTAs each) overlap with the SM-filling GEMMs of its neighbour
        self.n_lanes = max(1, int(n_lanes))
        self._lanes = {}
        # Launch width of the persistent tensor-core kernels inside the lanes' graphs (decaf_set_gemm_sms).  A full-width
        # GEMM holds every SM (one 640-thread CTA with ~200 KB of shared memory each) from its first TMA issue to its last
        # store, although for ~5 us of that - first bytes in flight, last epilogue - the SMs have nothing to do, and no other
        # lane's kernel fits beside it.  With launches a third of the device wide, three lanes' GEMMs run side by side and
        # those fixed costs idle a third of the SMs: 13.7k -> 15.8k pairs/s at the NLQ shape (profiles/README.md).  One video
        # at a time (n_lanes == 1, eager passes, the time-sharded MAD path) keeps the full width.
        if gemm_sms is None:
            gemm_sms = int(os.environ.get('DECAF_LANE_GEMM_SMS', '48')) if self.n_lanes > 1 else 0
        self.gemm_sms = int(gemm_sms)
        # Per-shape state — pinned host slots, device input buffers, CUDA graphs (with their private pools) and the engine's
        # activation workspaces (~40 MB per query at T = 2304) — is kept for the `max_cached_shapes` most recently used
        # (T, n_query, Lmax-bucket) shapes only; real evaluation sets vary in all three, so an unbounded cache grows
        # monotonically and eventually exhausts the device.  Eviction synchronises the device (rare: once per new shape beyond
        # the cap) and frees everything the evicted shape owned.
        self.max_cached_shapes = max(1, int(max_cached_shapes))
        self._shape_lru = OrderedDict()
        self.evictions = 0

    def reset(self):
        self.counts = np.zeros((len(self.ranks), len(self.iou_threshs)))
        self.text_cnt = 0
        self.itr = 0

    def load_model(self):
        """libs/worker_v2.py:806-812: <root>/models/<ckpt>.pth, key 'model_ema'."""
        filename = os.path.join(self.opt['_root'], 'models', f"{self.opt['_ckpt']}.pth")
        ckpt = torch.load(filename, map_location='cpu')
        self.model.load_state_dict(ckpt['model_ema'])

    # ------------------------------------------------------------------ loop + metrics
    @torch.no_grad()
    def run(self, train_time_data=None):
        """libs/worker_v2.py:814-911: per-video predict, R@k x IoU accumulation."""
        if train_time_data is not None:
            self.model = train_time_data[0]
        start = time.time()
        items = []

        def feed():
            for data in self.dataloader:
                if isinstance(data, (list, tuple)):
                    data = data[0]
                items.append(data)
                yield data
                if self.opt.get('aux', {}).get('dryrun', False):
                    return
        for segs, scores, count in self.predict_videos(feed(), raw=True):
            data = items.pop(0)
            targets = data['segment']
            assert len(count) == len(targets)
            self._accumulate_raw(segs, scores, count, targets)
            self.itr += 1
        metrics = self.counts / max(self.text_cnt, 1)
        log_str = "\nFinal:"
        for i, rank in enumerate(self.ranks):
            log_str += "\n-----"
            for j, thresh in enumerate(self.iou_threshs):
                log_str += f"\nRank@{rank}, IoU@{thresh:.1f}: {(metrics[i, j] * 100):.2f}"
        log_str += f"\n-----\nEvaluation completed in {time.time() - start:.1f}s."
        if self.logger is not None:
            self.logger.write(log_str)
        return metrics

    def _accumulate_raw(self, segs, scores, count, targets):
        """R@k x IoU counts of one video for all its queries at once (libs/worker_v2.py:857-878, libs/train_utils.py:81-96)
        from the padded result arrays (B, max_out, 2) / (B, max_out) / (B,): the rows come out of the NMS finalize kernel
        sorted by score, so the top-k are the first min(count, k) rows.  Same fp32 arithmetic as _accumulate, without ~10 small
        tensor ops per query (0.8 ms per 16-query video: more than the device needs for the whole video)."""
        B, K = scores.shape
        tg = np.asarray(targets, dtype=np.float32).reshape(B, 2)
        ps, pe = segs[..., 0], segs[..., 1]
        gs, ge = tg[:, None, 0], tg[:, None, 1]
        overlap = np.clip(np.minimum(pe, ge) - np.maximum(ps, gs), 0, None)
        union = (pe - ps) + (ge - gs) - overlap
        with np.errstate(divide='ignore', invalid='ignore'):
            iou_bk = overlap / union
        pos = np.arange(K)[None, :]
        valid = pos < np.minimum(count, self.topk)[:, None]
        for i, r in enumerate(self.ranks):
            m = valid & (pos < r)
            best = np.where(m, iou_bk, -np.inf).max(axis=1)
            best = np.where(m.any(axis=1), best, 0.0)
            self.counts[i] += (best[:, None] >= self.iou_threshs[None]).sum(axis=0)
        self.text_cnt += B

    def _accumulate(self, results, targets):
        for result, target in zip(results, targets):
            segs, scores = result['segments'], result['scores']
            idx = scores.argsort(descending=True)
            segs, scores = segs[idx[:self.topk]], scores[idx[:self.topk]]
            target = torch.as_tensor(target, dtype=torch.float)
            target = target.expand(len(segs), -1)
            iou_topk = iou(segs, target)
            iou_n = []
            for i in self.ranks:
                tmp = iou_topk[:i]
                iou_n.append(tmp.max().item() if len(tmp) > 0 else 0)
            iou_n = np.array(iou_n)
            self.counts += (iou_n[:, None] >= self.iou_threshs[None])
        self.text_cnt += len(targets)

    # ------------------------------------------------------------------ predict
    def padded_len(self, vid_len):
        """libs/worker_v2.py:969-976."""
        input_vid_len = self.input_vid_len
        if vid_len > input_vid_len:
            stride = self.min_chunk_size * self.vid_stride
            input_vid_len = (vid_len + (stride - 1)) // stride * stride
        return input_vid_len

    def _lane(self, lane):
        L = self._lanes.get(lane)
        if L is None:
            L = dict(stream=torch.cuda.Stream(), side=torch.cuda.Stream(), done=torch.cuda.Event())
            self._lanes[lane] = L
        return L

    def _stage_inputs(self, data, lane=0):
        """Pinned host staging + one async H2D per tensor (on the current stream).  Returns the lane's device tensors
        (d_vid (Ce,T), d_sh (Cs,T), d_mask (T,), d_tok (n,Lmax,Ctok), d_len (n,), d_cls (n,Cs), d_meta) together with
        the pinned h_* buffers they were filled from."""
        return self._upload(self._stage_host(data, lane), lane)

    @staticmethod
    def _small_specs(T, n, Lmax, Cs, Ctok):
        return [('mask', (T, ), torch.uint8), ('tok', (n, Lmax, Ctok), torch.float32), ('len', (n, ), torch.int32),
                ('cls', (n, Cs), torch.float32), ('meta', (5, ), torch.float32)]

    def _stage_host(self, data, slot=0):
        """CPU half of the staging: pad / transpose the item into pinned host buffers of slot `slot` (predict_videos keeps
        n_lanes + 2 slots so that the next video is staged while every lane is still busy)."""
        assert self.window_size is None, "sliding-window evaluation is not supported"
        assert self.window_stride is None, "sliding-window evaluation is not supported"
        if data.get('ext_scores') is not None:
            raise NotImplementedError('external scores are not on the released eval path')
        tokens = data['text']
        if not isinstance(tokens, tuple):
            tokens = (tokens, )
        vid, shallow = data['vid'], data['shallow_vid']
        vid_index = data.get('vid_index')                 # compact ingest: vid is (Ce, K) = the clips listed in vid_index
        vid_len = shallow.size(-1)
        T = self.padded_len(vid_len)
        n = len(tokens)
        lens = [t.size(-1) for t in tokens]
        Lm = max(lens)
        Lmax = (Lm + self.text_len_bucket - 1) // self.text_len_bucket * self.text_len_bucket
        Ce, Cs, Ctok = vid.size(0), shallow.size(0), tokens[0].size(0)
        skey = (T, n, Lmax, Ce, Cs, Ctok)
        self._touch_shape(skey)
        hs = self._stage.get(('h', skey, slot))
        if hs is None:
            pin = lambda *s, dtype=torch.float32: torch.zeros(*s, dtype=dtype).pin_memory()
            # the small per-video tensors share ONE pinned buffer (and one device buffer per lane): one H2D copy instead of five
            small = self._small_specs(T, n, Lmax, Cs, Ctok)
            h_small = pin(_carve(None, small)[1], dtype=torch.uint8)
            hs = dict(h_vid=pin(Ce, T), h_sh=pin(Cs, T), h_small=h_small, h_idx=None, skey=skey, prev_len=0, free=None)
            hs.update({'h_' + k: v for k, v in _carve(h_small, small)[0].items()})
            hs['np_len'], hs['np_meta'] = hs['h_len'].numpy(), hs['h_meta'].numpy()      # views of the pinned buffers
            hs['np_tok'], hs['prev_rows'] = hs['h_tok'].numpy(), [0] * n
            hs['np_cls'] = hs['h_cls'].numpy()
            self._stage[('h', skey, slot)] = hs
        if hs['free'] is not None:                        # the previous upload out of this slot has been read by the copy engine
            hs['free'].synchronize()
        # the pinned buffers start zeroed and only the tail a shorter input leaves behind is re-zeroed
        prev, prev_m = hs['prev_len'], hs.get('prev_mask', 0)        # non-zero extents of h_vid / h_sh and of h_mask
        if vid_len < prev_m:
            hs['h_mask'][vid_len:prev_m] = 0
        hs['h_mask'][:vid_len] = 1
        hs['prev_mask'] = vid_len
        # features that already live in pinned host memory (vid and shallow_vid both) are uploaded straight from there:
        # no second host copy (the caller keeps them unchanged until the video's results are out)
        direct = vid_index is None and vid.is_pinned() and shallow.is_pinned() and vid.dtype == torch.float32 and \
            shallow.dtype == torch.float32 and vid.is_contiguous() and shallow.is_contiguous()
        hs['direct'] = (vid, shallow, vid_len) if direct else None
        hs['K'] = None
        if direct:
            pass                                          # h_vid / h_sh keep whatever an earlier video left (prev_len unchanged)
        elif vid_index is None:
            if vid_len < prev:
                hs['h_vid'][:, vid_len:prev] = 0
                hs['h_sh'][:, vid_len:prev] = 0
            hs['h_vid'][:, :vid_len] = vid
            hs['h_sh'][:, :vid_len] = shallow
            hs['prev_len'] = vid_len
        else:
            # compact expert-feature ingest: data['vid'] holds only the K clips listed in
            # data['vid_index'] (a superset of what select_clips() reports); only those columns cross PCIe and are scattered
            # into the dense device buffer, every other step is zero - which is what the merge makes of unselected steps
            if vid_len < prev:
                hs['h_sh'][:, vid_len:prev] = 0
            hs['h_sh'][:, :vid_len] = shallow
            index = torch.as_tensor(vid_index, dtype=torch.int32)
            K = int(index.numel())
            assert vid.size(-1) == K, 'vid must be (C_e, K) with K = len(vid_index)'
            if hs['h_idx'] is None:
                hs['h_idx'] = torch.zeros(T, dtype=torch.int32).pin_memory()
            hs['h_vid'][:, :K] = vid                      # columns [0, K) now hold compact data: a later dense video re-zeroes
            hs['h_idx'][:K] = index                       # its tail up to prev_len
            hs['prev_len'] = max(prev, K, vid_len)
            hs['K'] = K
        # token matrices (C_tok, L_i) -> rows [0, L_i) of the pinned (n, Lmax, C_tok) batch.  Host tensors go through numpy
        # views of the pinned buffer (one strided assignment per query, ~1.5 us each; pad_sequence over the transposed views
        # cost 170 us per 16-query video - a third of the host time per video at the Charades shape); only the tail a shorter
        # query leaves behind in its row is re-zeroed
        if all(t.device.type == 'cpu' and t.dtype == torch.float32 for t in tokens):
            np_tok, prev_rows = hs['np_tok'], hs['prev_rows']
            for i, t in enumerate(tokens):
                L = lens[i]
                np_tok[i, :L] = t.detach().numpy().T
                if L < prev_rows[i]:
                    np_tok[i, L:prev_rows[i]] = 0
                prev_rows[i] = L
        else:
            padded = torch.nn.utils.rnn.pad_sequence([t.t() for t in tokens], batch_first=True)  # zeros beyond each L_i
            hs['h_tok'][:, :Lm].copy_(padded)
            prev_m = max(hs['prev_rows'])
            if Lm < prev_m:
                hs['h_tok'][:, Lm:prev_m] = 0
            hs['prev_rows'] = [Lm] * n
        hs['np_len'][:] = lens
        tc = data['text_cls']
        if tc.device.type == 'cpu' and tc.dtype == torch.float32:
            hs['np_cls'][:] = tc.detach().numpy()
        else:
            hs['h_cls'].copy_(tc)
        # seconds conversion constants of libs/worker_v2.py:1113-1122, read on the device by the NMS kernel
        hs['np_meta'][:] = (float(self.vid_stride), float(data.get('clip_stride', 1)), float(0.5 * data.get('clip_size', 0)),
                            float(data.get('fps', 1)), float(data.get('duration', 0)))
        return hs

    def _touch_shape(self, skey):
        lru = self._shape_lru
        if skey in lru:
            lru.move_to_end(skey)
            return
        lru[skey] = True
        while len(lru) > self.max_cached_shapes:
            old, _ = lru.popitem(last=False)
            self._evict_shape(old)

    def _evict_shape(self, skey):
        """Free everything cached for one (T, n_query, Lmax, Ce, Cs, Ctok) shape: graphs first (they reference the buffers),
        then staging buffers, then the engine workspaces no remaining shape shares."""
        torch.cuda.synchronize()                          # nothing of this shape may still be running / replaying
        T, n, Lmax = skey[:3]
        for k in [k for k in self._graphs if k[0][:len(skey)] == skey]:
            del self._graphs[k]
        for k in [k for k in self._stage if k[1] == skey]:
            del self._stage[k]
        if self.model is not None and getattr(self.model, '_engine', None) is not None:
            keep_plans = {(k[1], k[0]) for k in self._shape_lru}           # (n, T) still cached
            keep_text = {(k[1], k[2]) for k in self._shape_lru}            # (n, Lmax) still cached
            self.model.engine().drop_workspaces(n, T, Lmax, drop_plan=(n, T) not in keep_plans,
                                                drop_text=(n, Lmax) not in keep_text)
        self.evictions += 1
        torch.cuda.empty_cache()

    def _upload(self, hs, lane=0):
        """GPU half of the staging: async H2D copies (on the current stream) from a filled host slot into the device
        buffers of `lane` (the inputs the lane's CUDA graph was captured on)."""
        skey = hs['skey']
        T, n, Lmax, Ce, Cs, Ctok = skey
        ds = self._stage.get(('d', skey, lane))
        if ds is None:
            dev = lambda *s, dtype=torch.float32: torch.zeros(*s, dtype=dtype, device='cuda')
            small = self._small_specs(T, n, Lmax, Cs, Ctok)
            d_small = dev(_carve(None, small)[1], dtype=torch.uint8)
            ds = dict(d_vid=dev(Ce, T), d_sh=dev(Cs, T), d_small=d_small, d_idx=None, d_vidc=None, key=skey + (lane, ), lane=lane)
            ds.update({'d_' + k: v for k, v in _carve(d_small, small)[0].items()})
            self._stage[('d', skey, lane)] = ds
        K = hs['K']
        direct = hs.get('direct')
        if direct is not None:
            vid, shallow, vid_len = direct
            dl = ds.get('dev_len', T)                        # columns of d_vid / d_sh that may be non-zero
            if vid_len < dl:
                ds['d_vid'][:, vid_len:dl].zero_()
                ds['d_sh'][:, vid_len:dl].zero_()
            ds['d_vid'][:, :vid_len].copy_(vid, non_blocking=True)
            ds['d_sh'][:, :vid_len].copy_(shallow, non_blocking=True)
            ds['dev_len'] = vid_len
        elif K is None:
            ds['dev_len'] = T
            ds['d_vid'].copy_(hs['h_vid'], non_blocking=True)
        else:
            if ds['d_idx'] is None:
                ds['d_idx'] = torch.zeros(T, dtype=torch.int32, device='cuda')
                ds['d_vidc'] = torch.zeros_like(ds['d_vid'])
            ds['d_vidc'][:, :K].copy_(hs['h_vid'][:, :K], non_blocking=True)
            ds['d_idx'][:K].copy_(hs['h_idx'][:K], non_blocking=True)
        if K is not None:
            ds['dev_len'] = T
        if direct is None:
            ds['d_sh'].copy_(hs['h_sh'], non_blocking=True)
        ds['d_small'].copy_(hs['h_small'], non_blocking=True)       # mask, tokens, lengths, text_cls, metadata
        if hs['free'] is None:
            hs['free'] = torch.cuda.Event()
        hs['free'].record()
        if K is not None:
            cabi.scatter_clips(ds['d_vidc'], T, ds['d_idx'], Ce, K, ds['d_vid'], T)
        st = dict(ds)
        st.update({k: v for k, v in hs.items() if k.startswith('h_') and k != 'h_small' and v is not None})
        return st

    @torch.no_grad()
    def select_clips(self, data):
        """Which clips need expert features: runs the saliency scorer + exact top-k selection (libs/modeling/model.py:
        500-541) on the sidekick features alone.  Returns (union (t,) bool, per_query (n, t) bool) on the host; pass
        data['vid'][:, union] with data['vid_index'] = union.nonzero() to predict_video(s) for the compact ingest."""
        eng = self.model.engine()
        shallow = data['shallow_vid']
        t = shallow.size(-1)
        T = self.padded_len(t)
        n = data['text_cls'].size(0)
        torch.cuda.synchronize()                          # lane 0's workspaces are used below: nothing may be in flight on them
        eng.lane = 0
        p = eng.plan(n, T)
        sh = torch.zeros(shallow.size(0), T, device='cuda')
        sh[:, :t] = shallow.cuda(non_blocking=True)
        mask = torch.zeros(T, dtype=torch.uint8, device='cuda')
        mask[:t] = 1
        cls = data['text_cls'].float().cuda()
        cabi.saliency(sh, cls, p.correl, sh.shape[0], T, n, eng.norm)
        cabi.select(p.correl, mask, p.sel, p.mask0, p.pooled, p.max_blocks, T, n, eng.sn, eng.sratio,
                    and_mask=not eng.msf, vid_len_out=p.vid_len)
        per_query = p.sel[:, :t].bool().cpu()
        return per_query.any(0), per_query

    @torch.no_grad()
    def predict_video(self, data, return_outputs=False):
        """Batched fast path of simple_predict: stage -> text encode -> grounder -> decode -> NMS
        -> one D2H.  Returns results (and the reference-format outputs when asked)."""
        t0 = time.perf_counter()
        st = self._stage_inputs(data)
        eng = self.model.engine()
        t1 = time.perf_counter()
        p = self.run_staged(st)
        t2 = t3 = time.perf_counter()
        p.out_host.copy_(p.out_buf, non_blocking=True)     # the one D2H of the video: <= max_num_segs rows per query
        torch.cuda.current_stream().synchronize()
        nb = p.B * p.max_out
        segs = p.out_host[:2 * nb].view(p.B, p.max_out, 2)
        scores = p.out_host[2 * nb:3 * nb].view(p.B, p.max_out)
        count = p.out_host[3 * nb:].view(torch.int32)
        t4 = time.perf_counter()
        results = []
        for b in range(p.B):
            k = int(count[b])
            results.append({'segments': segs[b, :k].clone(), 'scores': scores[b, :k].clone()})
        self.time_dict['prepare'].append(t1 - t0)
        self.time_dict['forward'].append(t2 - t1)
        self.time_dict['post_process'].append(t3 - t2)
        self.time_dict['nms'].append(t4 - t3)
        if return_outputs:
            logits, offsets, masks = eng.level_views(p)
            pts = self.pt_gen([m.size(-1) for m in masks[0]])
            self.outputs = [logits, offsets, pts, masks]
            return results, self.outputs
        return results

    def _device_pass(self, st):
        """Text encoder on a forked stream, concurrently with the video-only prologue of the grounder (saliency ->
        selection -> merge -> vid_map -> first pre-attention block and query projection); the streams join right
        before the first cross-attention.  The text kernels are a dozen CTAs each, so they fit beside the prologue."""
        eng = self.model.engine()
        eng.lane = st.get('lane', 0)
        main = torch.cuda.current_stream()
        side = self._lane(eng.lane)['side']
        side.wait_stream(main)
        with torch.cuda.stream(side):
            text, kv_len, kv = eng.encode_text_batch(st['d_tok'], st['d_len'])
        p = eng.forward(st['d_vid'], st['d_sh'], st['d_mask'], text, kv_len, st['d_cls'], text_kv=kv,
                        text_ready=lambda: main.wait_stream(side))
        main.wait_stream(side)                     # (no-op when text_ready already joined)
        eng.decode(p)
        eng.nms(p, meta=st['d_meta'])
        return p

    def run_staged(self, st):
        """Everything between the H2D copies and the D2H read for one staged video: text encoder -> grounder
        forward -> decode -> NMS, as one CUDA-graph replay (captured on first use of a staging shape)."""
        if not self.use_graphs:
            return self._device_pass(st)
        gkey = (st['key'], st['d_vid'].data_ptr())
        entry = self._graphs.get(gkey)
        if entry is None:
            prev = cabi.set_gemm_sms(self.gemm_sms)     # the grid sizes are baked into the graph
            try:
                self._device_pass(st)                   # eager warm-up: plans, PE tables, function attributes
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    p = self._device_pass(st)
            finally:
                cabi.set_gemm_sms(prev)
            entry = (graph, p)
            self._graphs[gkey] = entry
        entry[0].replay()
        return entry[1]

    def launch_staged(self, st):
        """run_staged on the stream of the staging's lane (ordered after the current stream).  Pair with
        join_lanes() before reading results or recording an end-of-work event on the current stream."""
        L = self._lane(st.get('lane', 0))
        L['stream'].wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(L['stream']):
            return self.run_staged(st)

    def join_lanes(self):
        cur = torch.cuda.current_stream()
        for L in self._lanes.values():
            cur.wait_stream(L['stream'])

    def _raw_from_host(self, p, host=None):
        nb = p.B * p.max_out
        h = (p.out_host if host is None else host).numpy().copy()                   # one copy: the pinned buffer is reused by the lane's next video
        return (h[:2 * nb].reshape(p.B, p.max_out, 2), h[2 * nb:3 * nb].reshape(p.B, p.max_out),
                h[3 * nb:].view(np.int32)[:p.B])

    def _results_from_host(self, p, host=None):
        # ONE copy out of the pinned buffer (the lane's next video reuses it), then views into that copy
        nb = p.B * p.max_out
        h = (p.out_host if host is None else host).numpy().copy()
        segs = torch.from_numpy(h[:2 * nb].reshape(p.B, p.max_out, 2)).unbind(0)
        scores = torch.from_numpy(h[2 * nb:3 * nb].reshape(p.B, p.max_out)).unbind(0)
        count = h[3 * nb:].view(np.int32)[:p.B].tolist()
        return [{'segments': segs[b][:count[b]], 'scores': scores[b][:count[b]]} for b in range(p.B)]

    @torch.no_grad()
    def predict_videos(self, videos, raw=False):
        """Pipelined predict_video over an iterable of items: yields the results list of every video, in order.
        Up to n_lanes videos are in flight, each on its own stream with its own staging buffers, workspaces and
        CUDA graph: host staging and the H2D copies of the next video, the D2H of the previous one and the
        latency-bound kernels of both overlap with the current video's GEMMs.  Same arithmetic, same kernels and
        bit-identical results as predict_video (only the scheduling differs).  raw=True yields the padded arrays
        (segments (n, max_num_segs, 2), scores (n, max_num_segs), count (n,)) instead of per-query dicts."""
        harvest = self._raw_from_host if raw else self._results_from_host
        # Two videos are queued per lane: video i + n_lanes is enqueued on its lane's stream BEFORE the host has seen video i
        # finish, so a lane never idles for the host's reaction time (event wake-up, result harvest, staging and upload of
        # its next video: ~0.3 ms per video, 4-6 % of a lane's cycle when it waited).  What the second video of a lane
        # overwrites is ordered by the stream (device inputs, workspaces) except the pinned result buffer the host reads:
        # each lane alternates between two of them.  Host slots: 2 n_lanes + 2, so a slot is refilled only after the upload
        # that read it has completed (the FIFO below has synchronised on a later video of the same lane by then).
        depth = 2
        pending = []
        n_slots = depth * self.n_lanes + 2
        for i, data in enumerate(videos):
            if isinstance(data, (list, tuple)):
                data = data[0]
            hs = self._stage_host(data, i % n_slots)     # CPU staging of the next video while every lane is still busy
            lane = i % self.n_lanes
            L = self._lane(lane)
            if len(pending) == depth * self.n_lanes:    # FIFO: the oldest video in flight owns the result slot reused below
                ev_done, pp, host = pending.pop(0)
                ev_done.synchronize()
                yield harvest(pp, host)
            slot = (i // self.n_lanes) % depth
            with torch.cuda.stream(L['stream']):
                st = self._upload(hs, lane)
                p = self.run_staged(st)
                hosts = getattr(p, 'out_hosts', None)
                if hosts is None:
                    hosts = p.out_hosts = [p.out_host] + [torch.zeros_like(p.out_host).pin_memory() for _ in range(depth - 1)]
                hosts[slot].copy_(p.out_buf, non_blocking=True)
                done = L.setdefault('done_ring', [torch.cuda.Event() for _ in range(depth)])[slot]
                done.record()
            pending.append((done, p, hosts[slot]))
        for ev_done, pp, host in pending:
            ev_done.synchronize()
            yield harvest(pp, host)

    def simple_predict(self, data):
        """libs/worker_v2.py:921-928: (outputs, results, loss) with loss = the eval-time statistics of _calc_loss when the
        item carries ground-truth `target` segments, else an empty dict."""
        results, outputs = self.predict_video(data, return_outputs=True)
        loss = self._calc_loss(data, outputs) if data.get('target') is not None else {}
        return outputs, results, loss

    def _reg_ranges(self):
        """PtGenerator's per-level regression range (libs/modeling/model.py:690-702) as a (L, 2) device tensor."""
        if getattr(self, '_rr', None) is None:
            self._rr = torch.tensor(self.pt_gen.regression_range, dtype=torch.float32).cuda().contiguous()
        return self._rr

    @torch.no_grad()
    def _calc_loss(self, data, outputs=None):
        """libs/worker_v2.py:1029-1061: focal classification loss / 1 - IoU regression loss of every query against its
        ground-truth segment, each normalised by the query's number of positive points, averaged over the queries (NaN
        skipped).  One reduction kernel over the logits / offsets the forward left on the device (decaf_eval_loss);
        `outputs` (the reference-format lists) is accepted for signature parity and repacked when given explicitly."""
        eng = self.model.engine()
        targets = torch.as_tensor(np.asarray(data['target']), dtype=torch.float32).reshape(-1, 2) / float(self.vid_stride)
        n = targets.size(0)
        if outputs is not None and outputs is not getattr(self, 'outputs', None):
            hl, ho, hm, lv = self._pack_levels(outputs[0], outputs[1], outputs[3])
        else:
            T = self.padded_len(data['shallow_vid'].size(-1))
            pl = eng.plan(n, T)
            hl, ho, hm, lv = pl.logits2, pl.offsets, pl.hmask, pl.lv
        tr = self.opt['train']
        out = torch.zeros(n, 3, device='cuda')
        cabi.eval_loss(hl, ho, hm, lv, n, targets.cuda(), self._reg_ranges(), tr.get('center_sampling', 'radius') == 'radius',
                       float(tr['center_sampling_radius']), 0.2, 0.5, out)
        o = out.cpu().numpy().astype(np.float64)
        norm = np.maximum(o[:, 2], 1.0)
        cls, reg = o[:, 0] / norm, o[:, 1] / norm
        return {'cls_loss': float(np.nanmean(cls)) if n else float('nan'), 'reg_loss': float(np.nanmean(reg)) if n else float('nan')}

    # ------------------------------------------------------------------ reference-format entry points
    @torch.no_grad()
    def _forward(self, data):
        """libs/worker_v2.py:930-1026: returns [fpn_logits_list, fpn_offsets_list, fpn_points,
        fpn_masks_list]."""
        st = self._stage_inputs(data)
        eng = self.model.engine()
        text, kv_len, kv = eng.encode_text_batch(st['d_tok'], st['d_len'])
        p = eng.forward(st['d_vid'], st['d_sh'], st['d_mask'], text, kv_len, st['d_cls'], text_kv=kv)
        logits, offsets, masks = eng.level_views(p)
        pts = self.pt_gen([m.size(-1) for m in masks[0]])
        self.outputs = [logits, offsets, pts, masks]
        return self.outputs

    def _pack_levels(self, logits_list, offsets_list, masks_list):
        """Per-query per-level lists -> the padded flat layout the decode kernel reads."""
        B = len(logits_list)
        lens = [int(x.size(-1)) for x in logits_list[0]]
        lv = cabi.make_levels(lens)
        dev = logits_list[0][0].device
        hl = torch.zeros(B, lv.Pp, device=dev)
        ho = torch.zeros(B, lv.Pp, 2, device=dev)
        hm = torch.zeros(B, lv.Pp, dtype=torch.uint8, device=dev)
        for b in range(B):
            for l, n in enumerate(lens):
                o = lv.off[l]
                hl[b, o:o + n] = logits_list[b][l].reshape(-1)
                ho[b, o:o + n] = offsets_list[b][l].reshape(-1, 2)
                hm[b, o:o + n] = masks_list[b][l].reshape(-1).to(torch.uint8)
        return hl, ho, hm, lv

    def _decode_lists(self, logits_list, offsets_list, masks_list):
        hl, ho, hm, lv = self._pack_levels(logits_list, offsets_list, masks_list)
        B, K = hl.size(0), int(self.pre_nms_topk)
        dev = hl.device
        segs = torch.zeros(B, K, 2, device=dev)
        scores = torch.zeros(B, K, device=dev)
        idx = torch.zeros(B, K, dtype=torch.int32, device=dev)
        cnt = torch.zeros(B, dtype=torch.int32, device=dev)
        cabi.decode(hl, ho, hm, lv, B, True, float(self.pre_nms_thresh), K, float(self.seg_len_thresh),
                    segs, scores, idx, cnt)
        return segs, scores, idx, cnt

    @torch.no_grad()
    def _collect_segments(self, fpn_points, fpn_logits, fpn_offsets, fpn_masks, ext_scores):
        """libs/worker_v2.py:1131-1187 for one query -> (segs (k,2), scores (k,)) on device."""
        if ext_scores is not None:
            raise NotImplementedError('external scores are not on the released eval path')
        segs, scores, _, cnt = self._decode_lists([fpn_logits], [fpn_offsets], [fpn_masks])
        k = int(cnt[0])
        return segs[0, :k], scores[0, :k]

    @torch.no_grad()
    def _generate_proposals(self, data, outputs, window_ext=None, window_offset=0, idx=None):
        """libs/worker_v2.py:1063-1129."""
        assert window_ext is None and window_offset == 0 and idx is None
        fpn_logits_list, fpn_offsets_list, fpn_points, fpn_masks_list = outputs
        segs, scores, _, cnt = self._decode_lists(fpn_logits_list, fpn_offsets_list, fpn_masks_list)
        B, K = scores.shape
        nm = self.opt['nms']
        prm = cabi.NmsParams()
        prm.mode = {None: 0, 'nms': 1, 'soft_nms': 2}[nm['mode']]
        prm.iou_thresh, prm.sigma, prm.min_score = float(nm['iou_thresh']), float(nm['sigma']), float(nm['min_score'])
        prm.max_num_segs, prm.voting_thresh = int(nm['max_num_segs']), float(nm['voting_thresh'])
        prm.to_seconds = 1
        prm.vid_stride = float(self.vid_stride)
        prm.clip_stride, prm.half_clip_size = float(data['clip_stride']), float(0.5 * data['clip_size'])
        prm.fps, prm.duration = float(data['fps']), float(data['duration'])
        max_out = prm.max_num_segs if prm.max_num_segs > 0 else K
        dev = segs.device
        out_segs = torch.zeros(B, max_out, 2, device=dev)
        out_scores = torch.zeros(B, max_out, device=dev)
        out_count = torch.zeros(B, dtype=torch.int32, device=dev)
        ws = torch.empty(int(cabi.nms_workspace_bytes(B, K)), dtype=torch.uint8, device=dev)
        cabi.batched_nms(segs, scores, cnt, B, K, prm, out_segs, out_scores, out_count, ws)
        out_segs, out_scores, out_count = out_segs.cpu(), out_scores.cpu(), out_count.cpu()
        return [{'segments': out_segs[b, :int(out_count[b])], 'scores': out_scores[b, :int(out_count[b])]}
                for b in range(B)]

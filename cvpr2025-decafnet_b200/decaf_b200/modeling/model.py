"""Mirror of libs/modeling/model.py: the live grounder PtTransformerEarlyFusionIterative and
PtGenerator, with the reference's constructor / call signatures and state-dict layout.

The forward is executed by decaf_b200.engine.GrounderEngine (hand-written sm_100a kernels behind
the C ABI); all queries of a video are batched where the reference loops over them with batch
size 1 (libs/modeling/model.py:526-563).
"""
import torch
import torch.nn as nn

from .blocks import MaskedConv1D, _ParamsOnly
from .fusion import make_fusion
from .head import make_head
from .tcn import TCN
from .text_net import make_text_net
from .video_net import make_video_net


class PtTransformerEarlyFusionIterative(nn.Module):
    """libs/modeling/model.py:397-565 (eval path).  `act_dtype`: torch.bfloat16 (tcgen05 GEMMs,
    fp32 accumulation / LayerNorm / softmax / residual stream) or torch.float32 (the reference's
    FP32 configuration, eval.py:40-41)."""

    def __init__(self, opt, second_fusion=True, act_dtype=torch.bfloat16, gemm_impl=0):
        super().__init__()
        if second_fusion:
            raise NotImplementedError('create_model only builds second_fusion=False (libs/worker_v2.py:191-193)')
        self.opt = opt
        self.act_dtype = act_dtype
        self.gemm_impl = gemm_impl
        self.text_net = make_text_net(opt.model['text_net'])
        in_dim = opt.model.vid_net.in_dim
        if opt.model.msf:
            in_dim *= 2
        if opt.model.scat:
            in_dim += 1
        self.vid_map = MaskedConv1D(in_dim, opt.model.vid_net.embd_dim, 1)
        _opt = opt.model.vid_net.clone()
        _opt.in_dim = _opt.embd_dim
        self.vid_net = make_video_net(_opt)
        self.fusion = make_fusion(opt.model['fusion'])
        self.cls_head = make_head(opt.model['cls_head'])
        n_levels = opt.model.vid_net.arch[-1]
        self.refine = TCN(n_levels, 32, 32, num_layers=n_levels, in_map=True)
        opt.model.cls_head.embd_dim += 32          # the reference mutates its opt the same way (:426-428)
        self.cls_head2 = make_head(opt.model['cls_head'])
        opt.model.reg_head.embd_dim += 32
        self.reg_head = make_head(opt.model['reg_head'])
        self.second_fusion = second_fusion
        self._engine = None
        self.register_load_state_dict_post_hook(lambda m, k: setattr(m, '_engine', None))

    # ------------------------------------------------------------------ engine plumbing
    def engine(self):
        if self._engine is None:
            from ..engine import GrounderEngine
            dev = next(self.parameters()).device
            if dev.type != 'cuda':
                raise RuntimeError('decaf_b200 runs on CUDA only: call model.cuda() first (no CPU fallback)')
            self._engine = GrounderEngine(self.opt, self.state_dict(), act_dtype=self.act_dtype, device=dev,
                                          gemm_impl=self.gemm_impl)
        return self._engine

    def refresh(self):
        """Re-pack weights after an in-place parameter change."""
        self._engine = None

    # ------------------------------------------------------------------ reference API
    @torch.no_grad()
    def encode_text(self, tokens, token_masks):
        """tokens (1, C_tok, L) fp32, token_masks (1, 1, L) bool -> text (1, C_t, L+1),
        masks (1, 1, L+1).  libs/modeling/model.py:434-436."""
        eng = self.engine()
        assert tokens.size(0) == 1, 'the reference eval loop encodes one query at a time'
        L = tokens.size(-1)
        tok = tokens[0].t().contiguous()[None].float()                  # (1, L, C_tok) channels-last
        lens = token_masks.reshape(1, -1).sum(dim=1).to(torch.int32)
        text, _, _ = eng.encode_text_batch(tok, lens)
        out = text[0].t().contiguous()[None]                            # (1, C_t, L+1)
        mask = torch.cat((token_masks[..., :1], token_masks), dim=-1)
        return out, mask

    @torch.no_grad()
    def forward(self, vid, shallow_vid, vid_masks, text, text_cls, text_masks, text_size=None, mv_data=None,
                eval=False):
        """libs/modeling/model.py:473-565.  vid (1, C_e, T), shallow_vid (1, C_s, T), vid_masks
        (1, T) bool, text: list of n (1, C_t, L_i+1), text_cls (n, C_s), text_masks: list of n
        (1, 1, L_i+1).  Returns (fpn_logits_list, fpn_offsets_list, fpn_masks_list)."""
        if not eval:
            raise NotImplementedError('only the released evaluation path (eval=True) is implemented')
        assert mv_data is None
        assert vid.size(0) == 1, vid.size()
        eng = self.engine()
        n = len(text)
        L1 = max(t.size(-1) for t in text)
        Ct = text[0].size(1)
        tx = torch.zeros(n, L1, Ct, device=vid.device)
        kv_len = torch.zeros(n, dtype=torch.int32, device=vid.device)
        for i, (t, m) in enumerate(zip(text, text_masks)):
            tx[i, :t.size(-1)] = t[0].t()
            kv_len[i] = int(m.sum())
        p = eng.forward(vid[0].contiguous().float(), shallow_vid[0].contiguous().float(),
                        vid_masks[0].contiguous(), tx, kv_len, text_cls.contiguous().float())
        self._last_plan = p
        return eng.level_views(p)


class PtGenerator(nn.Module):
    """libs/modeling/model.py:668-743: per-level (p, 4) = [coordinate, reg_lo, reg_hi, stride]
    tables.  Constant-table construction; the decode kernel derives (coordinate, stride) from
    (level, index) itself."""

    def __init__(self, max_seq_len, num_fpn_levels, regression_range=4, sigma=1, use_offset=False):
        super().__init__()
        self.num_fpn_levels = num_fpn_levels
        assert max_seq_len % 2 ** (num_fpn_levels - 1) == 0
        self.max_seq_len = max_seq_len
        self.regression_range = ((0, regression_range), )
        assert 0 < sigma <= 1
        for l in range(1, num_fpn_levels):
            assert regression_range <= max_seq_len
            v_min = regression_range * sigma
            v_max = regression_range * 2
            if l == num_fpn_levels - 1:
                v_max = max(v_max, max_seq_len + 1)
            self.regression_range += ((v_min, v_max), )
            regression_range = v_max
        self.use_offset = use_offset
        tics = torch.arange(0, max_seq_len, 1.0)
        for l in range(num_fpn_levels):
            stride = 2 ** l
            pts = tics[::stride][:, None]
            if use_offset:
                pts = pts + 0.5 * stride
            rr = torch.as_tensor(self.regression_range[l], dtype=torch.float32)[None].repeat(len(pts), 1)
            st = torch.full((len(pts), 1), float(stride))
            self.register_buffer(f'points_{l}', torch.cat((pts, rr, st), 1), persistent=False)

    def forward(self, fpn_n_points):
        assert len(fpn_n_points) == self.num_fpn_levels
        out = tuple()
        for l, n_pts in enumerate(fpn_n_points):
            pts = getattr(self, f'points_{l}')
            assert n_pts <= len(pts), (
                'number of requested points {:d} cannot exceed max number '
                'of buffered points {:d}'.format(n_pts, len(pts)))
            out += (pts[:n_pts], )
        return out

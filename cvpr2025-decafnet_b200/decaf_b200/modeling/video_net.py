"""Mirror of libs/modeling/video_net.py: registry + VideoTransformer weight container."""
from copy import deepcopy
import math

import torch.nn as nn

from .blocks import MaskedConv1D, LayerNorm, TransformerEncoder, _ParamsOnly

backbones = dict()


def register_video_net(name):
    def decorator(module):
        backbones[name] = module
        return module
    return decorator


@register_video_net('transformer')
class VideoTransformer(_ParamsOnly):
    """libs/modeling/video_net.py:21-164 (same ctor kwargs)."""
    def __init__(self, in_dim, embd_dim, max_seq_len, n_heads, mha_win_size, stride=1, arch=(2, 1, 6),
                 attn_pdrop=0.0, proj_pdrop=0.0, path_pdrop=0.0, use_abs_pe=False, pool_only=False, **kwargs):
        super().__init__()
        assert len(arch) == 3, '(embed convs, stem, branch)'
        assert stride & (stride - 1) == 0
        assert arch[0] >= int(math.log2(stride))
        if stride != 1 or pool_only:
            raise NotImplementedError('vid_net.stride > 1 / pool_only are not on the released eval path')
        self.max_seq_len, self.use_abs_pe = max_seq_len, use_abs_pe
        self.embd_fc = MaskedConv1D(in_dim, embd_dim, 1)
        self.embd_convs, self.embd_norms = nn.ModuleList(), nn.ModuleList()
        for _ in range(arch[0]):
            self.embd_convs.append(MaskedConv1D(embd_dim, embd_dim, 3, 1, 1, bias=False))
            self.embd_norms.append(LayerNorm(embd_dim))
        self.stem = nn.ModuleList([
            TransformerEncoder(embd_dim, stride=1, n_heads=n_heads, window_size=mha_win_size)
            for _ in range(arch[1])])
        self.branch = nn.ModuleList([
            TransformerEncoder(embd_dim, stride=2 if idx > 0 else 1, n_heads=n_heads, window_size=mha_win_size)
            for idx in range(arch[2])])
        self._kw = dict(in_dim=in_dim, embd_dim=embd_dim, max_seq_len=max_seq_len, n_heads=n_heads, mha_win_size=mha_win_size,
                        arch=tuple(arch), use_abs_pe=use_abs_pe)
        self.act_dtype = None                    # None: bf16 on sm_100, else fp32

    def forward(self, x, mask):
        """libs/modeling/video_net.py:123-164: (fpn, fpn_masks), tuples over the FPN levels."""
        from ..engine import run_video_net
        return run_video_net(self, self._kw, x, mask)


def make_video_net(opt):
    opt = deepcopy(opt)
    return backbones[opt.pop('name')](**opt)

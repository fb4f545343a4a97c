"""Mirror of libs/modeling/head.py: registry + ClsHead / RegHead.  The modules hold the reference's parameters (same
state-dict layout) and their forward runs the same sm_100a kernels the engine uses for its head sub-graph."""
from copy import deepcopy

import numpy as np
import torch.nn as nn

from .blocks import MaskedConv1D, LayerNorm, Scale, _ParamsOnly

heads = dict()


def register_head(name):
    def decorator(module):
        heads[name] = module
        return module
    return decorator


@register_head('cls')
class ClsHead(_ParamsOnly):
    """libs/modeling/head.py:18-64."""
    def __init__(self, embd_dim, n_layers=2, prior_prob=0.0):
        super().__init__()
        self.convs, self.norms = nn.ModuleList(), nn.ModuleList()
        for _ in range(n_layers):
            self.convs.append(MaskedConv1D(embd_dim, embd_dim, 3, 1, 1, bias=False))
            self.norms.append(LayerNorm(embd_dim))
        self.cls_head = MaskedConv1D(embd_dim, 1, 3, 1, 1)
        assert 0 <= prior_prob < 1
        if prior_prob > 0:
            nn.init.constant_(self.cls_head.conv.bias, -np.log((1 - prior_prob) / prior_prob))

    def forward(self, fpn, fpn_masks):
        """libs/modeling/head.py:53-64: tuple of (bs, T_l) logits, tuple of (bs, T_l) masks."""
        from ..engine import run_head
        return run_head(self, fpn, fpn_masks, 1, 'cls_head')


@register_head('reg')
class RegHead(_ParamsOnly):
    """libs/modeling/head.py:67-108."""
    def __init__(self, embd_dim, num_fpn_levels, n_layers=2):
        super().__init__()
        self.convs, self.norms = nn.ModuleList(), nn.ModuleList()
        for _ in range(n_layers):
            self.convs.append(MaskedConv1D(embd_dim, embd_dim, 3, 1, 1, bias=False))
            self.norms.append(LayerNorm(embd_dim))
        self.reg_head = MaskedConv1D(embd_dim, 2, 3, 1, 1)
        self.scales = nn.ModuleList([Scale() for _ in range(num_fpn_levels)])

    def forward(self, fpn, fpn_masks):
        """libs/modeling/head.py:95-108: tuple of (bs, T_l, 2) offsets = ReLU(scale_l * conv), tuple of (bs, T_l) masks."""
        from ..engine import run_head
        return run_head(self, fpn, fpn_masks, 2, 'reg_head', scales=[sc.scale for sc in self.scales])


def make_head(opt):
    opt = deepcopy(opt)
    return heads[opt.pop('name')](**opt)

"""Mirror of libs/modeling/fusion.py: registry + XAttNFusion weight container."""
from copy import deepcopy

import torch.nn as nn

from .blocks import LayerNorm, TransformerDecoder, _ParamsOnly

modules = dict()


def register_fusion(name):
    def decorator(module):
        modules[name] = module
        return module
    return decorator


@register_fusion('xattn')
class XAttNFusion(_ParamsOnly):
    """libs/modeling/fusion.py:16-78 (same ctor kwargs)."""
    def __init__(self, vid_dim, text_dim, n_layers=2, n_heads=4, attn_pdrop=0.0, proj_pdrop=0.0,
                 path_pdrop=0.0, xattn_mode='adaln'):
        super().__init__()
        self.layers = nn.ModuleList([
            TransformerDecoder(vid_dim, text_dim, n_heads=n_heads, xattn_mode=xattn_mode)
            for _ in range(n_layers)])
        self.ln_out = LayerNorm(vid_dim)
        self._kw = dict(vid_dim=vid_dim, text_dim=text_dim, n_layers=n_layers, n_heads=n_heads)
        self.act_dtype = None                    # None: bf16 on sm_100, else fp32

    def forward(self, vid, vid_masks, text, text_mask, text_size=None):
        """libs/modeling/fusion.py:56-78."""
        from ..engine import run_fusion
        return run_fusion(self, self._kw, vid, vid_masks, text, text_mask, text_size)


def make_fusion(opt):
    opt = deepcopy(opt)
    return modules[opt.pop('name')](**opt)

"""Mirror of libs/modeling/text_net.py: registry + TextTransformer weight container."""
from copy import deepcopy

import torch
import torch.nn as nn

from .blocks import MaskedConv1D, TransformerEncoder, _ParamsOnly

backbones = dict()


def register_text_net(name):
    def decorator(module):
        backbones[name] = module
        return module
    return decorator


@register_text_net('transformer')
class TextTransformer(_ParamsOnly):
    """libs/modeling/text_net.py:92-188 (same ctor kwargs)."""
    def __init__(self, in_dim, embd_dim, n_heads, max_seq_len, n_layers=5, attn_pdrop=0.0, proj_pdrop=0.0,
                 path_pdrop=0.0, use_abs_pe=True, use_bkgd_token=True):
        super().__init__()
        if not use_bkgd_token:
            raise NotImplementedError('use_bkgd_token=False is not on the released eval path')
        self.max_seq_len, self.use_abs_pe = max_seq_len, use_abs_pe
        self.embd_fc = MaskedConv1D(in_dim, embd_dim, 1)
        self.bkgd_token = nn.Parameter(torch.empty(embd_dim, 1))
        nn.init.trunc_normal_(self.bkgd_token, mean=0.0, std=0.02)
        self.transformer = nn.ModuleList([
            TransformerEncoder(embd_dim, stride=0, n_heads=n_heads) for _ in range(n_layers)])
        self._text_opt = dict(in_dim=in_dim, embd_dim=embd_dim, n_heads=n_heads, max_seq_len=max_seq_len, n_layers=n_layers,
                              use_abs_pe=use_abs_pe, use_bkgd_token=use_bkgd_token)
        self.act_dtype = torch.float32           # torch.bfloat16 selects the tensor-core launches of the bf16 configuration

    def forward(self, x, mask):
        """libs/modeling/text_net.py:158-188: x (bs, C_tok, L), mask (bs, 1, L) -> (x (bs, C_t, L + 1), mask (bs, 1, L + 1));
        every row of the batch is encoded with the PE / key mask of its own length."""
        from ..engine import run_text_net
        return run_text_net(self, self._text_opt, x, mask)


@register_text_net('identity')
class TextIdentity(_ParamsOnly):
    """libs/modeling/text_net.py:22-89 — registered for API parity; not built (the live model
    `create_model` instantiates, libs/worker_v2.py:182-211, uses 'transformer')."""
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError("text_net 'identity' is not on the released eval path")


def make_text_net(opt):
    opt = deepcopy(opt)
    return backbones[opt.pop('name')](**opt)

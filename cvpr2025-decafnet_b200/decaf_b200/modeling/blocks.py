"""Parameter containers mirroring libs/modeling/blocks.py of the reference.

These modules exist so that a reference checkpoint (`ckpt['model_ema']`, libs/worker_v2.py:
806-812) loads with `load_state_dict` unchanged: same module tree, same parameter names and
shapes.  They hold weights only — the arithmetic is done by libdecaf_b200.so through
decaf_b200.engine.GrounderEngine, never by these modules (calling one raises).
"""
import torch
import torch.nn as nn


class _ParamsOnly(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError(
            f'{type(self).__name__} is a weight container; run the model through '
            'decaf_b200 (GrounderEngine / PtTransformerEarlyFusionIterative), which executes the '
            'sm_100a kernels. No PyTorch fallback exists.')


class _Conv(_ParamsOnly):
    """Stands in for nn.Conv1d (weight (Cout, Cin/groups, k), optional bias)."""
    def __init__(self, cin, cout, k, groups=1, bias=True):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin // groups, k))
        nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)
        if bias:
            self.bias = nn.Parameter(torch.zeros(cout))
        else:
            self.register_parameter('bias', None)


class MaskedConv1D(_ParamsOnly):
    """libs/modeling/blocks.py:63-106."""
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, groups=1, bias=True):
        super().__init__()
        self.stride = stride
        self.conv = _Conv(in_channels, out_channels, kernel_size, groups, bias)


class LayerNorm(_ParamsOnly):
    """libs/modeling/blocks.py:109-131."""
    def __init__(self, n_channels, affine=True, eps=1e-5):
        super().__init__()
        self.n_channels, self.eps = n_channels, eps
        if affine:
            self.weight = nn.Parameter(torch.ones(n_channels, 1))
            self.bias = nn.Parameter(torch.zeros(n_channels, 1))
        else:
            self.weight = self.bias = None


class MaskedMHA(_ParamsOnly):
    """libs/modeling/blocks.py:145-393."""
    def __init__(self, embd_dim, q_dim=None, kv_dim=None, out_dim=None, n_heads=4, window_size=0,
                 attn_pdrop=0.0, proj_pdrop=0.0):
        super().__init__()
        assert embd_dim % n_heads == 0
        q_dim = q_dim or embd_dim
        kv_dim = kv_dim or embd_dim
        out_dim = out_dim or q_dim
        self.n_heads, self.window_size = n_heads, window_size
        assert window_size == 0 or window_size % 2 == 1
        self.query = _Conv(q_dim, embd_dim, 1)
        self.key = _Conv(kv_dim, embd_dim, 1)
        self.value = _Conv(kv_dim, embd_dim, 1)
        self.proj = _Conv(embd_dim, out_dim, 1)


class ConvAttNLayer(_ParamsOnly):
    """libs/modeling/blocks.py:414-473."""
    def __init__(self, embd_dim, out_dim=None, stride=1, n_heads=4, window_size=0, attn_pdrop=0.0, proj_pdrop=0.0):
        super().__init__()
        self.use_conv = stride > 0
        if self.use_conv:
            assert stride == 1 or stride % 2 == 0
            for n in 'qkv':
                setattr(self, f'{n}_conv', MaskedConv1D(embd_dim, embd_dim, 3, stride, 1, groups=embd_dim, bias=False))
            for n in 'qkv':
                setattr(self, f'{n}_norm', LayerNorm(embd_dim))
        self.attn = MaskedMHA(embd_dim, out_dim=out_dim or embd_dim, n_heads=n_heads, window_size=window_size)


class ConvXAttNLayer(_ParamsOnly):
    """libs/modeling/blocks.py:476-520."""
    def __init__(self, embd_dim, kv_dim, out_dim=None, stride=1, n_heads=4, attn_pdrop=0.0, proj_pdrop=0.0):
        super().__init__()
        self.use_conv = stride > 0
        if self.use_conv:
            self.q_conv = MaskedConv1D(embd_dim, embd_dim, 3, stride, 1, groups=embd_dim, bias=False)
            self.q_norm = LayerNorm(embd_dim)
        self.xattn = MaskedMHA(embd_dim, kv_dim=kv_dim, out_dim=out_dim or embd_dim, n_heads=n_heads)


class FFN(_ParamsOnly):
    """libs/modeling/blocks.py:523-538."""
    def __init__(self, channels, expansion=4, pdrop=0.0):
        super().__init__()
        self.fc = _Conv(channels, channels * expansion, 1)
        self.proj = _Conv(channels * expansion, channels, 1)


class LayerScale(_ParamsOnly):
    """libs/modeling/blocks.py:670-682."""
    def __init__(self, n_channels, pdrop=0.0, init_scale=1e-4):
        super().__init__()
        self.scale = nn.Parameter(init_scale * torch.ones((1, n_channels, 1)))


class Scale(_ParamsOnly):
    """libs/modeling/blocks.py:653-667."""
    def __init__(self, init=1.0):
        super().__init__()
        self.scale = nn.Parameter(torch.as_tensor(init, dtype=torch.float))


class TransformerEncoder(_ParamsOnly):
    """libs/modeling/blocks.py:541-591."""
    def __init__(self, embd_dim, stride=1, n_heads=4, window_size=0, expansion=4, attn_pdrop=0.0,
                 proj_pdrop=0.0, path_pdrop=0.0):
        super().__init__()
        self.stride = stride
        self.attn = ConvAttNLayer(embd_dim, stride=stride, n_heads=n_heads, window_size=window_size)
        self.ln_attn = LayerNorm(embd_dim)
        self.drop_path_attn = LayerScale(embd_dim, path_pdrop)
        self.ffn = FFN(embd_dim, expansion, proj_pdrop)
        self.ln_ffn = LayerNorm(embd_dim)
        self.drop_path_ffn = LayerScale(embd_dim, path_pdrop)


class TransformerDecoder(_ParamsOnly):
    """libs/modeling/blocks.py:594-650."""
    def __init__(self, embd_dim, kv_dim, n_heads=4, expansion=4, attn_pdrop=0.0, proj_pdrop=0.0,
                 path_pdrop=0.0, xattn_mode='adaln'):
        super().__init__()
        assert xattn_mode in ('affine', 'adaln')
        if xattn_mode != 'adaln':
            raise NotImplementedError("only xattn_mode='adaln' is on the released eval path")
        self.xattn = ConvXAttNLayer(embd_dim, kv_dim, embd_dim * 2, stride=1, n_heads=n_heads)
        self.ln_xattn_q = LayerNorm(embd_dim)
        self.ln_xattn_kv = LayerNorm(kv_dim)
        self.adaln = LayerNorm(embd_dim, affine=False)
        self.ffn = FFN(embd_dim, expansion, proj_pdrop)
        self.ln_ffn = LayerNorm(embd_dim)
        self.drop_path_ffn = LayerScale(embd_dim, path_pdrop)

"""Mirror of libs/modeling/tcn.py (weight containers)."""
import torch.nn as nn

from .blocks import _Conv, _ParamsOnly


class DilatedResidualLayer(_ParamsOnly):
    """libs/modeling/tcn.py:4-38."""
    def __init__(self, dilation, nchannels, dropout=0.5, layernorm=True, layernorm_eps=1e-5, ngroup=1):
        super().__init__()
        assert layernorm and ngroup == 1
        self.dilation = dilation
        self.conv_dilated = _Conv(nchannels, nchannels, 3)
        self.conv_1x1 = _Conv(nchannels, nchannels, 1)
        self.norm = nn.LayerNorm(nchannels, eps=layernorm_eps)


class TCN(_ParamsOnly):
    """libs/modeling/tcn.py:40-84."""
    def __init__(self, in_dim, hid_dim, out_dim, num_layers, dropout=0.5, dilation_factor=2, ln=True,
                 ngroup=1, in_map=False):
        super().__init__()
        assert in_map and ln and dilation_factor == 2
        self.conv_1x1 = _Conv(in_dim, hid_dim, 1)
        self.layers = nn.ModuleList([DilatedResidualLayer(dilation_factor ** i, hid_dim, dropout)
                                     for i in range(num_layers)])
        self.conv_out = _Conv(hid_dim, out_dim, 1)

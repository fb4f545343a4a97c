"""TEST / BENCH INFRASTRUCTURE (oracle side): parameter names and shapes of the reference model's state dict,
derived from the option tree alone.

`bench.py --impl reference` and the oracle need weights in the reference's `state_dict` layout without importing the
product package (whose modules load libdecaf_b200.so).  This restates the constructors' parameter shapes:
PtTransformerEarlyFusionIterative.__init__ (libs/modeling/model.py:401-432), TextTransformer (text_net.py:102-156),
VideoTransformer (video_net.py:32-121), XAttNFusion / TransformerDecoder (fusion.py:21-54, blocks.py:594-650),
ConvAttNLayer / ConvXAttNLayer / MaskedMHA / FFN / TransformerEncoder / LayerScale (blocks.py:145-220, 396-591,
653-682), ClsHead / RegHead (head.py:23-108), TCN / DilatedResidualLayer (tcn.py:9-84).
tests/test_cabi_cpu.py checks it key by key against the product mirror's state dict (itself checked against the
reference's state-dict layout through the golden fixtures' `state_keys` / `state_shapes`).
"""

R_REFINE = 32            # libs/modeling/model.py:424


def _mha(out, pre, embd, q_dim=None, kv_dim=None, out_dim=None):
    q_dim, kv_dim = q_dim or embd, kv_dim or embd
    out_dim = out_dim or q_dim
    for n, cin, cout in (('query', q_dim, embd), ('key', kv_dim, embd), ('value', kv_dim, embd), ('proj', embd, out_dim)):
        out[f'{pre}{n}.weight'] = (cout, cin, 1)
        out[f'{pre}{n}.bias'] = (cout,)


def _ln(out, pre, c):
    out[f'{pre}weight'] = (c, 1)
    out[f'{pre}bias'] = (c, 1)


def _ffn(out, pre, c, expansion=4):
    out[f'{pre}ffn.fc.weight'] = (expansion * c, c, 1)
    out[f'{pre}ffn.fc.bias'] = (expansion * c,)
    out[f'{pre}ffn.proj.weight'] = (c, expansion * c, 1)
    out[f'{pre}ffn.proj.bias'] = (c,)
    _ln(out, f'{pre}ln_ffn.', c)
    out[f'{pre}drop_path_ffn.scale'] = (1, c, 1)


def _encoder(out, pre, c, conv):
    """TransformerEncoder (blocks.py:541-591); conv=False is the text encoder's stride 0."""
    if conv:
        for n in 'qkv':
            out[f'{pre}attn.{n}_conv.conv.weight'] = (c, 1, 3)
        for n in 'qkv':
            _ln(out, f'{pre}attn.{n}_norm.', c)
    _mha(out, f'{pre}attn.attn.', c)
    _ln(out, f'{pre}ln_attn.', c)
    out[f'{pre}drop_path_attn.scale'] = (1, c, 1)
    _ffn(out, pre, c)


def _head_tower(out, pre, c, n_layers, fin, n_out):
    for i in range(n_layers):
        out[f'{pre}convs.{i}.conv.weight'] = (c, c, 3)           # bias=False: followed by LayerNorm (head.py:33-44)
        _ln(out, f'{pre}norms.{i}.', c)
    out[f'{pre}{fin}.conv.weight'] = (n_out, c, 3)
    out[f'{pre}{fin}.conv.bias'] = (n_out,)


def state_dict_shapes(opt):
    m = opt['model']
    vn, tn, fu = m['vid_net'], m['text_net'], m['fusion']
    C, Ct, Ctok, Cin = vn['embd_dim'], tn['embd_dim'], tn['in_dim'], vn['in_dim']
    L = vn['arch'][2]
    out = {}
    # text net
    if tn.get('use_bkgd_token', True):
        out['text_net.bkgd_token'] = (Ct, 1)
    out['text_net.embd_fc.conv.weight'] = (Ct, Ctok, 1)
    out['text_net.embd_fc.conv.bias'] = (Ct,)
    for i in range(tn.get('n_layers', 5)):
        _encoder(out, f'text_net.transformer.{i}.', Ct, conv=False)
    # vid_map input width (model.py:410-416)
    if not m['msf']:
        cin_map = Cin
    elif m['sfonly']:
        cin_map = Cin
    else:
        cin_map = 2 * Cin
    if m['scat']:
        cin_map += 1
    out['vid_map.conv.weight'] = (C, cin_map, 1)
    out['vid_map.conv.bias'] = (C,)
    # video net
    out['vid_net.embd_fc.conv.weight'] = (C, C, 1)
    out['vid_net.embd_fc.conv.bias'] = (C,)
    for i in range(vn['arch'][0]):
        out[f'vid_net.embd_convs.{i}.conv.weight'] = (C, C, 3)
        _ln(out, f'vid_net.embd_norms.{i}.', C)
    for i in range(vn['arch'][1]):
        _encoder(out, f'vid_net.stem.{i}.', C, conv=True)
    for i in range(L):
        _encoder(out, f'vid_net.branch.{i}.', C, conv=True)
    # fusion (adaln: the cross-attention projects to 2C = scale | shift)
    for i in range(fu['n_layers']):
        p = f'fusion.layers.{i}.'
        out[f'{p}xattn.q_conv.conv.weight'] = (C, 1, 3)
        _ln(out, f'{p}xattn.q_norm.', C)
        _mha(out, f'{p}xattn.xattn.', C, q_dim=C, kv_dim=Ct, out_dim=2 * C)
        _ln(out, f'{p}ln_xattn_q.', C)
        _ln(out, f'{p}ln_xattn_kv.', Ct)
        _ffn(out, p, C)
    _ln(out, 'fusion.ln_out.', C)
    # heads + refinement TCN (model.py:418-432: the second heads see C + 32 channels)
    _head_tower(out, 'cls_head.', C, m['cls_head']['n_layers'], 'cls_head', 1)
    out['refine.conv_1x1.weight'] = (R_REFINE, L, 1)
    out['refine.conv_1x1.bias'] = (R_REFINE,)
    for i in range(L):
        p = f'refine.layers.{i}.'
        out[f'{p}conv_dilated.weight'] = (R_REFINE, R_REFINE, 3)
        out[f'{p}conv_dilated.bias'] = (R_REFINE,)
        out[f'{p}conv_1x1.weight'] = (R_REFINE, R_REFINE, 1)
        out[f'{p}conv_1x1.bias'] = (R_REFINE,)
        out[f'{p}norm.weight'] = (R_REFINE,)
        out[f'{p}norm.bias'] = (R_REFINE,)
    out['refine.conv_out.weight'] = (R_REFINE, R_REFINE, 1)
    out['refine.conv_out.bias'] = (R_REFINE,)
    C2 = C + R_REFINE
    _head_tower(out, 'cls_head2.', C2, m['cls_head']['n_layers'], 'cls_head', 1)
    _head_tower(out, 'reg_head.', C2, m['reg_head']['n_layers'], 'reg_head', 2)
    for l in range(L):
        out[f'reg_head.scales.{l}.scale'] = ()
    return out

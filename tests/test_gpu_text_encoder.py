"""GPU: the one-launch text encoder (decaf_text_encoder: a cluster of 8 CTAs per query, activations in
distributed shared memory) against (a) the same computation composed from decaf_gemm / decaf_layernorm /
decaf_xattn launches and (b) the oracle's restatement of TextTransformer.forward
(libs/modeling/text_net.py:158-188) + the fusion key/value projections (libs/modeling/blocks.py:640-641)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double(), b.double()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-12)


def _engine(opt):
    from decaf_b200 import synth
    from decaf_b200.worker_v2 import Evaluator, create_model
    shapes = {k: tuple(v.shape) for k, v in create_model(opt.clone()).state_dict().items()}
    sd = synth.fill_state_dict(shapes, 11)
    ev = Evaluator(opt.clone(), dataset=[], state_dict=sd, act_dtype=torch.float32, gemm_impl=1)
    eng = ev.model.engine()
    eng.fused_text = True
    return eng, sd


CONFIGS = [
    dict(kind='nlq', n=16, Lmax=24),            # canonical: Ct 128, 4 heads, 5 layers, tok 768, C 256
    dict(kind='nlq', n=3, Lmax=31),             # fills all 32 rows
    dict(kind='nlq_pe', n=5, Lmax=12),          # absolute PE on the words
    dict(kind='tiny', n=4, Lmax=8),             # Ct 32, head dim 8, tok 48, C 64
    dict(kind='smoke', n=7, Lmax=20),           # Ct 64, head dim 16, C 128
]


@pytest.mark.parametrize('cfg', CONFIGS, ids=lambda c: f"{c['kind']}-n{c['n']}-L{c['Lmax']}")
def test_fused_text_encoder_matches_composed_and_oracle(cfg):
    from decaf_b200 import _cabi as cabi, synth
    from oracle import grounder_oracle as go
    kind = cfg['kind']
    if kind == 'nlq':
        opt = synth.nlq_opt(n_levels=4, win=9, max_seq_len=256)
    elif kind == 'nlq_pe':
        opt = synth.nlq_opt(n_levels=4, win=9, max_seq_len=256, text_abs_pe=True, text_max_len=32)
    elif kind == 'tiny':
        opt = synth.tiny_opt()
    else:
        opt = synth.tiny_opt(embd_dim=128, n_levels=5, win=9, max_seq_len=256, sn=12, vid_in_dim=64, text_dim=64)
    eng, sd = _engine(opt)
    n, Lmax = cfg['n'], cfg['Lmax']
    tn = opt.model.text_net
    assert cabi.text_encoder_supported(Lmax, tn.embd_dim, tn.in_dim, tn.n_heads, tn.n_layers, opt.model.vid_net.embd_dim,
                                       opt.model.fusion.n_layers)
    g = torch.Generator().manual_seed(5)
    lens = torch.randint(1, Lmax + 1, (n,), generator=g)
    lens[0] = Lmax
    tok = torch.zeros(n, Lmax, tn.in_dim)
    for i in range(n):
        tok[i, :lens[i]] = torch.randn(int(lens[i]), tn.in_dim, generator=g)
    d_tok, d_len = tok.cuda(), lens.to(torch.int32).cuda()
    xt_f, kvl_f, kv_f = eng._encode_text_fused(d_tok, d_len)
    xt_f, kvl_f, kv_f = xt_f.clone(), kvl_f.clone(), kv_f.clone()
    xt_c, kvl_c = eng._encode_text_composed(d_tok, d_len)
    kv_c = eng.text_kv(xt_c, n, Lmax + 1)
    torch.cuda.synchronize()
    assert torch.equal(kvl_f.cpu(), (lens + 1).to(torch.int32))
    assert torch.equal(kvl_c.cpu(), kvl_f.cpu())
    L1 = Lmax + 1
    valid = (torch.arange(L1)[None] < (lens + 1)[:, None])
    assert _rel(xt_f.cpu() * valid[..., None], xt_c.cpu() * valid[..., None]) < 2e-5
    vk = valid.reshape(-1)
    for f in range(opt.model.fusion.n_layers):
        for j in range(2):
            assert _rel(kv_f[f, j].cpu()[vk], kv_c[f, j].cpu()[vk]) < 2e-5
    # oracle: one query at a time, exactly like the reference loop (libs/worker_v2.py:940-955)
    for i in range(n):
        L = int(lens[i])
        t_ref, _ = go.text_net_forward(sd, opt, tok[i, :L].t()[None].contiguous(), torch.ones(1, 1, L, dtype=torch.bool))
        got = xt_f[i, :L + 1].cpu().t()[None]
        assert _rel(got, t_ref) < 1e-4, i


@pytest.mark.parametrize('variant', ['fused', 'composed', 'tc'])
def test_abs_pe_is_per_query_when_lengths_straddle_max_seq_len(variant):
    """use_abs_pe=True with a batch whose query lengths straddle text_net.max_seq_len: the reference encodes every query
    alone (libs/worker_v2.py:945-955) and interpolates the PE table to THAT query's length only when it is longer than
    max_seq_len (libs/modeling/text_net.py:163-172); short queries of the same video keep the raw rows.  Also covers a
    length bucket that pushes Lmax past a max_seq_len which is not a multiple of 4."""
    from decaf_b200 import synth
    from decaf_b200.worker_v2 import Evaluator, create_model
    from oracle import grounder_oracle as go
    opt = synth.nlq_opt(n_levels=4, win=9, max_seq_len=256, text_abs_pe=True, text_max_len=14)
    shapes = {k: tuple(v.shape) for k, v in create_model(opt.clone()).state_dict().items()}
    sd = synth.fill_state_dict(shapes, 11)
    bf16 = variant == 'tc'
    ev = Evaluator(opt.clone(), dataset=[], state_dict=sd, act_dtype=torch.bfloat16 if bf16 else torch.float32,
                   gemm_impl=0 if bf16 else 1)
    eng = ev.model.engine()
    tn = opt.model.text_net
    lens = torch.tensor([5, 14, 15, 23, 9, 28, 13, 2])
    Lmax = 28
    g = torch.Generator().manual_seed(9)
    tok = torch.zeros(len(lens), Lmax, tn.in_dim)
    for i, L in enumerate(lens.tolist()):
        tok[i, :L] = torch.randn(L, tn.in_dim, generator=g)
    d_tok, d_len = tok.cuda(), lens.to(torch.int32).cuda()
    if variant == 'fused':
        xt, _, _ = eng._encode_text_fused(d_tok, d_len)
    elif variant == 'composed':
        xt, _ = eng._encode_text_composed(d_tok, d_len)
    else:
        assert eng.text_tc
        xt, _ = eng._encode_text_tc(d_tok, d_len)
    torch.cuda.synchronize()
    tol = 3e-2 if bf16 else 1e-4
    for i, L in enumerate(lens.tolist()):
        t_ref, _ = go.text_net_forward(sd, opt, tok[i, :L].t()[None].contiguous(), torch.ones(1, 1, L, dtype=torch.bool))
        assert _rel(xt[i, :L + 1].cpu().t()[None], t_ref) < tol, (variant, i, L)

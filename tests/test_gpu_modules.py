"""GPU: the modules returned by the builder registries (make_video_net / make_text_net / make_fusion / make_head) are callable
with the reference's forward signatures (libs/modeling/video_net.py:123-164, text_net.py:158-188, fusion.py:56-78,
head.py:53-64, 95-108) and run the same sm_100a kernels as the engine's corresponding sub-graph; checked against the oracle's
restatement of each forward on the module's own (synthetic) weights, fp32 configuration <= 1e-3 and bf16 within the stated
tolerance of the small-width tier."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-12)


def _fill(module, seed, prefix='m.'):
    """Synthetic weights for a stand-alone module (the rules of synth.fill_state_dict key on the reference's dotted names)."""
    from decaf_b200 import synth
    shapes = {prefix + k: tuple(v.shape) for k, v in module.state_dict().items()}
    sd = {k[len(prefix):]: v for k, v in synth.fill_state_dict(shapes, seed).items()}
    module.load_state_dict(sd)
    return sd


def _opt():
    from decaf_b200 import synth
    return synth.tiny_opt(embd_dim=128, n_levels=5, win=9, max_seq_len=256, sn=12, vid_in_dim=64, text_dim=64)


@pytest.mark.parametrize('act_dtype,tol', [(torch.float32, 1e-3), (torch.bfloat16, 3e-2)])
def test_make_head_modules_are_callable(act_dtype, tol):
    from decaf_b200.modeling import make_head
    from oracle import grounder_oracle as go
    opt = _opt()
    g = torch.Generator().manual_seed(1)
    C, B = 128, 3
    lens = [256, 128, 64, 32, 16]
    fpn = tuple(torch.randn(B, C, n, generator=g) for n in lens)
    valid = torch.tensor([256, 200, 77])
    masks = tuple((torch.arange(n)[None] < ((valid + 2 ** l - 1) // 2 ** l)[:, None])[:, None] for l, n in enumerate(lens))
    fpn = tuple(f * m for f, m in zip(fpn, masks))
    for name, key in (('cls_head', 'cls'), ('reg_head', 'reg')):
        o = dict(opt.model[name])
        m = make_head(o).cuda()
        m.act_dtype = act_dtype
        sd = _fill(m, 31, prefix=f'{name}.')
        out, out_masks = m(tuple(f.cuda() for f in fpn), tuple(x.cuda() for x in masks))
        sdp = {f'h.{k}': v for k, v in sd.items()}
        ref = (go.cls_head_forward if key == 'cls' else go.reg_head_forward)(sdp, 'h.', list(fpn), list(masks), o['n_layers'])
        assert len(out) == len(lens)
        for l in range(len(lens)):
            assert out[l].shape == ref[l].shape and torch.equal(out_masks[l].cpu(), masks[l].squeeze(1))
            mm = masks[l].squeeze(1)
            mm = mm if key == 'cls' else mm[..., None].expand_as(ref[l])
            assert _rel(out[l].cpu()[mm], ref[l][mm]) < tol, (name, l)


def test_make_text_net_module_is_callable():
    from decaf_b200.modeling import make_text_net
    from oracle import grounder_oracle as go
    opt = _opt()
    m = make_text_net(opt.model.text_net).cuda()
    sd = _fill(m, 32)
    sdp = {f'text_net.{k}': v for k, v in sd.items()}
    g = torch.Generator().manual_seed(2)
    lens = [9, 4, 12]
    L = max(lens)
    x = torch.randn(3, opt.model.text_net.in_dim, L, generator=g)
    mask = (torch.arange(L)[None] < torch.tensor(lens)[:, None])[:, None]
    out, out_mask = m(x.cuda(), mask.cuda())
    assert out.shape == (3, opt.model.text_net.embd_dim, L + 1) and out_mask.shape == (3, 1, L + 1)
    for i, n in enumerate(lens):
        ref, rm = go.text_net_forward(sdp, opt, x[i:i + 1, :, :n], torch.ones(1, 1, n, dtype=torch.bool))
        assert _rel(out[i:i + 1, :, :n + 1], ref) < 1e-4 and bool(out_mask[i, 0, :n + 1].all()) and not bool(out_mask[i, 0, n + 1:].any())


@pytest.mark.parametrize('act_dtype,tol', [(torch.float32, 1e-3), (torch.bfloat16, 3e-2)])
def test_make_video_net_module_is_callable(act_dtype, tol):
    from decaf_b200.modeling import make_video_net
    from oracle import grounder_oracle as go
    opt = _opt()
    vo = opt.model.vid_net.clone()
    vo.in_dim = vo.embd_dim
    m = make_video_net(vo).cuda()
    m.act_dtype = act_dtype
    sd = _fill(m, 33)
    sdp = {f'vid_net.{k}': v for k, v in sd.items()}
    oo = opt.clone()
    oo.model.vid_net.in_dim = vo.embd_dim
    g = torch.Generator().manual_seed(3)
    B, T = 2, 256
    x = torch.randn(B, vo.embd_dim, T, generator=g)
    mask = torch.arange(T)[None] < torch.tensor([256, 190])[:, None]
    fpn, fpn_masks = m(x.cuda(), mask.cuda())
    ref, ref_masks = go.video_net_forward(sdp, oo, x, mask)
    assert len(fpn) == vo.arch[2]
    for l in range(len(fpn)):
        assert fpn[l].shape == ref[l].shape and torch.equal(fpn_masks[l].cpu(), ref_masks[l])
        mm = ref_masks[l].expand_as(ref[l])
        assert _rel(fpn[l].cpu()[mm], ref[l][mm]) < tol, l


@pytest.mark.parametrize('act_dtype,tol', [(torch.float32, 1e-3), (torch.bfloat16, 3e-2)])
def test_make_fusion_module_is_callable(act_dtype, tol):
    from decaf_b200.modeling import make_fusion
    from oracle import grounder_oracle as go
    opt = _opt()
    m = make_fusion(opt.model.fusion).cuda()
    m.act_dtype = act_dtype
    sd = _fill(m, 34)
    sdp = {f'fusion.{k}': v for k, v in sd.items()}
    g = torch.Generator().manual_seed(4)
    B, T, L = 2, 192, 10
    C, Ct = opt.model.fusion.vid_dim, opt.model.fusion.text_dim
    qmask = (torch.arange(T)[None] < torch.tensor([192, 150])[:, None])[:, None]
    q = torch.randn(B, C, T, generator=g) * qmask
    kmask = (torch.arange(L)[None] < torch.tensor([10, 6])[:, None])[:, None]
    kv = torch.randn(B, Ct, L, generator=g) * kmask
    out, om = m(q.cuda(), qmask.cuda(), kv.cuda(), kmask.cuda())
    for b in range(B):
        n = int(kmask[b].sum())
        ref, _ = go.fusion_forward(sdp, opt, q[b:b + 1], qmask[b:b + 1], kv[b:b + 1, :, :n], kmask[b:b + 1, :, :n])
        mm = qmask[b:b + 1].expand_as(ref)
        assert _rel(out[b:b + 1].cpu()[mm], (ref * qmask[b:b + 1])[mm]) < tol, b

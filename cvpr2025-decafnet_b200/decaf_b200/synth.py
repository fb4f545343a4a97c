"""Deterministic synthetic weights and inputs for the DeCaf-Grounder hot path.

The reference ships neither a checkpoint nor data (reference README.md:42-44), so every
parity test, golden fixture and benchmark runs on tensors generated here.  Generation is
keyed by *name* (crc32 of the state-dict key / input name mixed with a seed), never by
module construction order, so the same values are produced whether the consumer is the
reference model (tests/golden/make_golden.py), the oracle (oracle/grounder_oracle.py) or
the CUDA path.

Weight scales are chosen so that every sub-block is numerically visible in the outputs
(the reference initialises LayerScale to 1e-4, libs/modeling/blocks.py:675-678, which would
hide attention/FFN errors) and so that decoded segments survive ``seg_len_thresh``
(reg-head bias ~ +1).
"""
import copy
import zlib

import numpy as np
import torch


class AttrDict(dict):
    """Minimal stand-in for yacs.config.CfgNode (attribute + key access, ``clone()``).

    The reference reads ``opt.model.vid_net`` both ways and calls ``.clone()`` on it
    (libs/modeling/model.py:404-428)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        return copy.deepcopy(self)

    def get(self, k, d=None):
        return self[k] if k in self else d


def make_opt(
    embd_dim=256, text_dim=128, tok_dim=768, vid_in_dim=256, n_levels=8, win=19,
    max_seq_len=2304, n_heads=4, msf=True, scat=False, sfonly=False, norm=True,
    sn=60, sratio=0.3, text_max_len=24, text_layers=5, text_abs_pe=False,
    vid_abs_pe=True, fusion_layers=2, head_layers=2, n_embd_convs=2,
    pre_nms_thresh=0.001, pre_nms_topk=2000, seg_len_thresh=0.1,
    nms_mode='soft_nms', iou_thresh=0.1, min_score=0.001, max_num_segs=5,
    sigma=0.9, voting_thresh=0.95,
):
    """Build the option tree the reference model/evaluator read (field names follow
    libs/core/opt.py:77-194 plus the derived fields of ``_update_opt`` :458-488)."""
    model = AttrDict(
        name='iter',
        text_net=AttrDict(name='transformer', in_dim=tok_dim, embd_dim=text_dim,
                          max_seq_len=text_max_len, n_heads=n_heads, n_layers=text_layers,
                          use_abs_pe=text_abs_pe, use_bkgd_token=True),
        vid_net=AttrDict(name='transformer', in_dim=vid_in_dim, embd_dim=embd_dim,
                         n_heads=n_heads, max_seq_len=max_seq_len, stride=1,
                         arch=(n_embd_convs, 0, n_levels), mha_win_size=win,
                         attn_pdrop=0.0, proj_pdrop=0.1, path_pdrop=0.1,
                         use_abs_pe=vid_abs_pe, fuse='cat', pool_only=False, cdrop=0.0),
        fusion=AttrDict(name='xattn', n_layers=fusion_layers, n_heads=n_heads,
                        attn_pdrop=0.0, proj_pdrop=0.1, path_pdrop=0.1,
                        xattn_mode='adaln', text_dim=text_dim, vid_dim=embd_dim),
        cls_head=AttrDict(name='cls', n_layers=head_layers, prior_prob=0.0,
                          embd_dim=embd_dim),
        reg_head=AttrDict(name='reg', n_layers=head_layers, embd_dim=embd_dim,
                          num_fpn_levels=n_levels),
        sn=sn, sratio=sratio, msf=msf, scat=scat, sfonly=sfonly, norm=norm,
        max_vid_len=max_seq_len, max_text_len=text_max_len, vid_stride=1,
        num_fpn_levels=n_levels, mha_win_size=win,
    )
    opt = AttrDict(
        model=model,
        pt_gen=AttrDict(regression_range=4, sigma=0.5, num_fpn_levels=n_levels,
                        max_seq_len=max_seq_len * 4),
        eval=AttrDict(ranks=(1, 5), iou_threshs=(0.3, 0.5),
                      pre_nms_thresh=pre_nms_thresh, pre_nms_topk=pre_nms_topk,
                      seg_len_thresh=seg_len_thresh, window_size=None,
                      window_stride=None),
        nms=AttrDict(mode=nms_mode, iou_thresh=iou_thresh, min_score=min_score,
                     max_num_segs=max_num_segs, sigma=sigma,
                     voting_thresh=voting_thresh),
        train=AttrDict(num_workers=0, center_sampling='radius',
                       center_sampling_radius=1.5),
        aux=AttrDict(dryrun=False),
    )
    return opt


# canonical shapes ("Canonical shapes")
def nlq_opt(**kw):
    return make_opt(**kw)


def charades_opt(**kw):
    d = dict(embd_dim=128, text_dim=128, vid_in_dim=128, n_levels=6, win=5,
             max_seq_len=256, sn=8)
    d.update(kw)
    return make_opt(**d)


def tiny_opt(**kw):
    """Small config for golden fixtures and fast CPU tests."""
    d = dict(embd_dim=64, text_dim=32, tok_dim=48, vid_in_dim=40, n_levels=4, win=5,
             max_seq_len=64, sn=6, sratio=0.3, text_layers=2, pre_nms_topk=50)
    d.update(kw)
    return make_opt(**d)


def _rng(seed, name):
    return np.random.Generator(np.random.PCG64([seed, zlib.crc32(name.encode())]))


def synth_tensor(seed, name, shape, kind='normal', scale=1.0, shift=0.0, lo=0.0, hi=1.0):
    g = _rng(seed, name)
    if kind == 'normal':
        a = g.standard_normal(shape, dtype=np.float64) * scale + shift
    elif kind == 'uniform':
        a = g.uniform(lo, hi, shape)
    else:
        raise ValueError(kind)
    return torch.from_numpy(np.asarray(a, dtype=np.float32).reshape(shape))


def fill_state_dict(shapes, seed=2022):
    """Deterministic values for a reference-layout state dict.

    ``shapes``: mapping key -> shape (e.g. ``{k: tuple(v.shape) for k, v in
    model.state_dict().items()}``).  Returns key -> float32 tensor."""
    out = {}
    for k in sorted(shapes):
        shp = tuple(shapes[k])
        leaf = k.split('.')[-1]
        if k.endswith('drop_path_attn.scale') or k.endswith('drop_path_ffn.scale'):
            v = synth_tensor(seed, k, shp, 'uniform', lo=0.5, hi=1.0)
        elif '.scales.' in k:                       # reg_head.scales.i.scale ()
            v = synth_tensor(seed, k, shp, 'uniform', lo=0.8, hi=1.2)
        elif k.endswith('bkgd_token'):
            v = synth_tensor(seed, k, shp, 'normal', 0.5)
        elif leaf == 'weight' and len(shp) == 3:    # conv weights (Cout, Cin/groups, k)
            fan_in = shp[1] * shp[2]
            v = synth_tensor(seed, k, shp, 'normal', 1.0 / np.sqrt(fan_in))
        elif leaf == 'weight':                      # LayerNorm weights (C,1) / (C,)
            v = synth_tensor(seed, k, shp, 'normal', 0.1, 1.0)
        elif leaf == 'bias' and k.endswith('reg_head.conv.bias'):
            v = synth_tensor(seed, k, shp, 'normal', 0.1, 1.0)
        elif leaf == 'bias' and k.endswith('cls_head.conv.bias'):
            v = synth_tensor(seed, k, shp, 'normal', 0.3, -1.0)
        elif leaf == 'bias':
            v = synth_tensor(seed, k, shp, 'normal', 0.1)
        else:
            raise KeyError(f'no synthetic rule for {k} {shp}')
        out[k] = v
    return out


def synth_video(opt, vid_len, n_query, seed=2022, tag='v0', text_len_range=(6, 24),
                n_events=0):
    """One item with the dict schema of libs/data/dataset.py:977-994.

    vid / shallow_vid: (C, vid_len) float32 N(0,1); text: tuple of (C_tok, L_i);
    text_cls: (n, C_s); fps 30, clip_size 32, clip_stride 16.
    ``n_events`` > 0 plants that many (query-correlated) events in the sidekick
    stream so the saliency scores are not flat."""
    m = opt.model
    ce = m.vid_net.in_dim
    cs = ce
    ctok = m.text_net.in_dim
    vid = synth_tensor(seed, f'{tag}.vid', (ce, vid_len))
    shallow = synth_tensor(seed, f'{tag}.shallow', (cs, vid_len))
    text_cls = synth_tensor(seed, f'{tag}.text_cls', (n_query, cs))
    g = _rng(seed, f'{tag}.lens')
    lens = g.integers(text_len_range[0], text_len_range[1] + 1, size=n_query)
    lens = np.minimum(lens, m.text_net.max_seq_len)
    text = tuple(synth_tensor(seed, f'{tag}.text{i}', (ctok, int(l)))
                 for i, l in enumerate(lens))
    segs = []
    for i in range(n_query):
        c = g.uniform(0.1, 0.9) * vid_len
        w = g.uniform(0.01, 0.1) * vid_len
        lo, hi = max(0.0, c - w), min(float(vid_len), c + w)
        segs.append((lo, hi))
        if n_events:
            a, b = int(lo), max(int(lo) + 1, int(hi))
            shallow[:, a:b] += 0.5 * text_cls[i][:, None]
    clip_stride, clip_size, fps = 16, 32, 30.0
    duration = vid_len * clip_stride / fps
    segs = np.asarray(segs, dtype=np.float32)
    seg_sec = (segs * clip_stride + 0.5 * clip_size) / fps
    return {
        'fps': fps, 'num_frames': vid_len * clip_stride, 'duration': duration,
        'segment': seg_sec, 'clip_size': clip_size, 'clip_stride': clip_stride,
        'target': torch.from_numpy(segs), 'clip_id': tag, 'text_id': list(range(n_query)),
        'vid': vid, 'shallow_vid': shallow, 'text': text, 'text_cls': text_cls,
        'ext_scores': None,
    }

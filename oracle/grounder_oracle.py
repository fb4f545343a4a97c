"""ORACLE (test infrastructure, NOT product code): CPU restatement of the DeCaf-Grounder
inference path of ZijiaLewisLu/CVPR2025-DeCafNet, written from the reference's semantics.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.  The product path (cvpr2025-decafnet_b200/) never does.

Parity status: PINNED.  tests/golden/*.npz hold outputs of the *unmodified reference*
(imported from /root/reference in the build container by tests/golden/make_golden.py);
tests/test_oracle_golden.py checks this restatement against them.  The reference itself has
no tests or golden vectors.

Everything is functional torch-CPU code over a reference-layout ``state_dict`` (fp32, or
fp64 when ``dtype=torch.float64``).  Each function cites the reference lines it restates
(paths relative to the reference repo root).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- blocks
def layer_norm_c(x, w=None, b=None, eps=1e-5):
    """Channel LayerNorm on (B, C, T): two-pass, biased variance, eps inside sqrt.
    libs/modeling/blocks.py:125-131."""
    x = x - x.mean(dim=1, keepdim=True)
    sigma = (x ** 2).mean(dim=1, keepdim=True)
    x = x / torch.sqrt(sigma + eps)
    if w is not None:
        x = x * w + b
    return x


def masked_conv1d(x, mask, w, b=None, stride=1, padding=0, groups=1):
    """conv(x * mask); stride-2 mask = nearest down-sample = mask[..., ::2].
    libs/modeling/blocks.py:87-106."""
    x = F.conv1d(x * mask.to(x.dtype), w, b, stride=stride, padding=padding, groups=groups)
    if stride > 1:
        mask = F.interpolate(mask.to(x.dtype), size=x.size(-1), mode='nearest').bool()
    return x, mask


def masked_max_pool1d(x, mask, kernel_size=3, stride=2):
    """libs/modeling/blocks.py:31-47 — equals "max over valid taps, 0 if none valid"
    (the global-amin fill never wins a window that holds a valid element)."""
    neg = torch.finfo(x.dtype).min
    xf = torch.where(mask, x, torch.full_like(x, neg))
    pad = (kernel_size - 1) // 2
    xp = F.max_pool1d(xf, kernel_size, stride, pad)
    mp = F.max_pool1d(mask.to(x.dtype), kernel_size, stride, pad)
    return torch.where(mp > 0, xp, torch.zeros_like(xp)), mp.bool()


def sinusoid_encoding(seq_len, n_freqs):
    """libs/modeling/blocks.py:134-142."""
    tics = torch.arange(seq_len, dtype=torch.float)
    freqs = 10000 ** torch.linspace(0, 1, n_freqs + 1)[:n_freqs]
    x = tics[None, :] / freqs[:, None]
    return torch.cat((torch.sin(x), torch.cos(x)))


def abs_pe(max_seq_len, embd_dim, t, dtype):
    """PE buffer (C, max_seq_len) / sqrt(C); linearly interpolated (align_corners) when the
    eval sequence is longer.  libs/modeling/video_net.py:74-79,143-152."""
    pe = sinusoid_encoding(max_seq_len, embd_dim // 2) / embd_dim ** 0.5
    pe = pe.to(dtype)
    if t > max_seq_len:
        pe = F.interpolate(pe[None], size=t, mode='linear', align_corners=True)[0]
    return pe[..., :t]


def mha_global(sd, pre, q, k, v, kv_mask, n_heads):
    """MaskedMHA, window_size == 0 (global / cross attention).
    libs/modeling/blocks.py:348-351,374-393: scale d^-1/4 on q and k, -inf on masked
    keys, no query masking."""
    q = F.conv1d(q, sd[pre + 'query.weight'], sd[pre + 'query.bias'])
    k = F.conv1d(k, sd[pre + 'key.weight'], sd[pre + 'key.bias'])
    v = F.conv1d(v, sd[pre + 'value.weight'], sd[pre + 'value.bias'])
    bs, c, _ = q.shape
    h = n_heads
    d = c // h
    scale = 1.0 / math.sqrt(math.sqrt(d))
    q = q.view(bs, h, d, -1).transpose(2, 3)
    k = k.view(bs, h, d, -1)
    v = v.view(bs, h, d, -1).transpose(2, 3)
    attn = (q * scale) @ (k * scale)
    attn = attn.masked_fill(torch.logical_not(kv_mask[:, :, None, :]), float('-inf'))
    attn = F.softmax(attn, dim=-1)
    o = attn @ v
    o = o.transpose(2, 3).reshape(bs, c, -1)
    return F.conv1d(o, sd[pre + 'proj.weight'], sd[pre + 'proj.bias'])


def mha_local(sd, pre, q, k, v, mask, n_heads, win):
    """MaskedMHA, local window (libs/modeling/blocks.py:357-373 with the chunked helpers
    :224-325).  Restated as a direct band: key j = t - s + i, i in [0, w); -inf outside
    [0, T); additive -1e4 on masked keys (:277-285); masked query rows zeroed (:293)."""
    q = F.conv1d(q, sd[pre + 'query.weight'], sd[pre + 'query.bias'])
    k = F.conv1d(k, sd[pre + 'key.weight'], sd[pre + 'key.bias'])
    v = F.conv1d(v, sd[pre + 'value.weight'], sd[pre + 'value.bias'])
    bs, c, t = q.shape
    h = n_heads
    d = c // h
    s = win // 2
    scale = 1.0 / math.sqrt(math.sqrt(d))
    q = (q * scale).view(bs, h, d, t)
    k = (k * scale).view(bs, h, d, t)
    v = v.view(bs, h, d, t)
    kp = F.pad(k, (s, s)).unfold(3, win, 1)             # (bs,h,d,t,w)
    vp = F.pad(v, (s, s)).unfold(3, win, 1)
    attn = torch.einsum('bhdt,bhdtw->bhtw', q, kp)
    pos = torch.arange(t)[:, None] - s + torch.arange(win)[None, :]   # (t,w)
    oob = (pos < 0) | (pos >= t)
    key_valid = F.pad(mask.to(q.dtype), (s, s)).unfold(2, win, 1)      # (bs,1,t,w)
    add = torch.where(key_valid > 0, 0.0, -1e4).to(q.dtype)
    attn = attn + add
    attn = attn.masked_fill(oob[None, None], float('-inf'))
    attn = F.softmax(attn, dim=-1)
    attn = attn.masked_fill(torch.logical_not(mask)[:, :, :, None], 0.0)
    o = torch.einsum('bhtw,bhdtw->bhdt', attn, vp).reshape(bs, c, t)
    return F.conv1d(o, sd[pre + 'proj.weight'], sd[pre + 'proj.bias'])


def ffn(sd, pre, x):
    """libs/modeling/blocks.py:535-538 (erf GELU, dropout off)."""
    x = F.gelu(F.conv1d(x, sd[pre + 'fc.weight'], sd[pre + 'fc.bias']))
    return F.conv1d(x, sd[pre + 'proj.weight'], sd[pre + 'proj.bias'])


def transformer_encoder(sd, pre, x, mask, stride, n_heads, win):
    """libs/modeling/blocks.py:578-591 (+ ConvAttNLayer :462-473).
    stride == 0: no convs (text net); win == 0: global attention."""
    mf = mask.to(x.dtype)
    x = x * mf
    if stride > 1:
        skip = masked_max_pool1d(x, mask, 3, stride)[0]
    else:
        skip = x
    ln = layer_norm_c(x, sd[pre + 'ln_attn.weight'], sd[pre + 'ln_attn.bias'])
    if stride > 0:
        a = pre + 'attn.'
        k, _ = masked_conv1d(ln, mask, sd[a + 'k_conv.conv.weight'], None, stride, 1, ln.size(1))
        v, _ = masked_conv1d(ln, mask, sd[a + 'v_conv.conv.weight'], None, stride, 1, ln.size(1))
        q, mask = masked_conv1d(ln, mask, sd[a + 'q_conv.conv.weight'], None, stride, 1, ln.size(1))
        q = layer_norm_c(q, sd[a + 'q_norm.weight'], sd[a + 'q_norm.bias'])
        k = layer_norm_c(k, sd[a + 'k_norm.weight'], sd[a + 'k_norm.bias'])
        v = layer_norm_c(v, sd[a + 'v_norm.weight'], sd[a + 'v_norm.bias'])
    else:
        q = k = v = ln
    if win > 0:
        h = mha_local(sd, pre + 'attn.attn.', q, k, v, mask, n_heads, win)
    else:
        h = mha_global(sd, pre + 'attn.attn.', q, k, v, mask, n_heads)
    mf = mask.to(x.dtype)
    x = skip * mf + sd[pre + 'drop_path_attn.scale'] * h
    h = ffn(sd, pre + 'ffn.', layer_norm_c(x, sd[pre + 'ln_ffn.weight'], sd[pre + 'ln_ffn.bias'])) * mf
    x = x + sd[pre + 'drop_path_ffn.scale'] * h
    return x, mask


def transformer_decoder(sd, pre, q, q_mask, kv, kv_mask, n_heads):
    """libs/modeling/blocks.py:632-650 (+ ConvXAttNLayer :513-520), xattn_mode 'adaln'."""
    mf = q_mask.to(q.dtype)
    q = q * mf
    lq = layer_norm_c(q, sd[pre + 'ln_xattn_q.weight'], sd[pre + 'ln_xattn_q.bias'])
    lkv = layer_norm_c(kv, sd[pre + 'ln_xattn_kv.weight'], sd[pre + 'ln_xattn_kv.bias'])
    a = pre + 'xattn.'
    qc, _ = masked_conv1d(lq, q_mask, sd[a + 'q_conv.conv.weight'], None, 1, 1, lq.size(1))
    qc = layer_norm_c(qc, sd[a + 'q_norm.weight'], sd[a + 'q_norm.bias'])
    h = mha_global(sd, a + 'xattn.', qc, lkv, lkv, kv_mask, n_heads)
    q = layer_norm_c(q * mf)                              # adaln: affine=False
    scale, shift = h.chunk(2, dim=1)
    q = q * scale + shift
    h = ffn(sd, pre + 'ffn.', layer_norm_c(q, sd[pre + 'ln_ffn.weight'], sd[pre + 'ln_ffn.bias'])) * mf
    q = q + sd[pre + 'drop_path_ffn.scale'] * h
    return q, q_mask


# ----------------------------------------------------------------------------- nets
def text_net_forward(sd, opt, tokens, mask):
    """TextTransformer.forward, libs/modeling/text_net.py:158-188.
    tokens (1, C_tok, L), mask (1, 1, L) bool -> (1, C_t, L+1), (1, 1, L+1)."""
    tn = opt['model']['text_net']
    assert tn['name'] == 'transformer'
    x, _ = masked_conv1d(tokens, mask, sd['text_net.embd_fc.conv.weight'],
                         sd['text_net.embd_fc.conv.bias'])
    t = x.size(-1)
    if tn.get('use_abs_pe', True):
        pe = abs_pe(tn['max_seq_len'], tn['embd_dim'], t, x.dtype)
        x = x + pe * mask.to(x.dtype)
    if tn.get('use_bkgd_token', True):
        x = torch.cat((sd['text_net.bkgd_token'][None].to(x.dtype).repeat(x.size(0), 1, 1), x), dim=-1)
        mask = torch.cat((mask[..., :1], mask), dim=-1)
    n_layers = tn.get('n_layers', 5)
    for i in range(n_layers):
        x, _ = transformer_encoder(sd, f'text_net.transformer.{i}.', x, mask, 0, tn['n_heads'], 0)
    return x, mask


def fusion_forward(sd, opt, q, q_mask, kv, kv_mask):
    """XAttNFusion._forward, libs/modeling/fusion.py:56-66."""
    fo = opt['model']['fusion']
    for i in range(fo['n_layers']):
        q, q_mask = transformer_decoder(sd, f'fusion.layers.{i}.', q, q_mask, kv, kv_mask, fo['n_heads'])
    q = layer_norm_c(q, sd['fusion.ln_out.weight'], sd['fusion.ln_out.bias'])
    return q, q_mask


def video_net_forward(sd, opt, x, mask):
    """VideoTransformer.forward, libs/modeling/video_net.py:123-164 (stride 1, no stem
    restriction: arch = (n_convs, n_stem, n_branch))."""
    vn = opt['model']['vid_net']
    assert vn['stride'] == 1
    arch = vn['arch']
    if mask.ndim == 2:
        mask = mask.unsqueeze(1)
    x, _ = masked_conv1d(x, mask, sd['vid_net.embd_fc.conv.weight'], sd['vid_net.embd_fc.conv.bias'])
    for i in range(arch[0]):
        x, mask = masked_conv1d(x, mask, sd[f'vid_net.embd_convs.{i}.conv.weight'], None, 1, 1)
        x = F.relu(layer_norm_c(x, sd[f'vid_net.embd_norms.{i}.weight'], sd[f'vid_net.embd_norms.{i}.bias']))
    t = x.size(-1)
    if vn['use_abs_pe']:
        pe = abs_pe(vn['max_seq_len'], vn['embd_dim'], t, x.dtype)
        x = x + pe * mask.to(x.dtype)
    for i in range(arch[1]):
        x, mask = transformer_encoder(sd, f'vid_net.stem.{i}.', x, mask, 1, vn['n_heads'], vn['mha_win_size'])
    fpn, fpn_masks = [], []
    for i in range(arch[2]):
        x, mask = transformer_encoder(sd, f'vid_net.branch.{i}.', x, mask, 2 if i > 0 else 1,
                                      vn['n_heads'], vn['mha_win_size'])
        fpn.append(x)
        fpn_masks.append(mask)
    return fpn, fpn_masks


def cls_head_forward(sd, pre, fpn, fpn_masks, n_layers):
    """ClsHead.forward, libs/modeling/head.py:53-64."""
    out = []
    for x, mask in zip(fpn, fpn_masks):
        for i in range(n_layers):
            x, _ = masked_conv1d(x, mask, sd[f'{pre}convs.{i}.conv.weight'], None, 1, 1)
            x = F.relu(layer_norm_c(x, sd[f'{pre}norms.{i}.weight'], sd[f'{pre}norms.{i}.bias']))
        logits, _ = masked_conv1d(x, mask, sd[pre + 'cls_head.conv.weight'], sd[pre + 'cls_head.conv.bias'], 1, 1)
        out.append(logits.squeeze(1))
    return out


def reg_head_forward(sd, pre, fpn, fpn_masks, n_layers):
    """RegHead.forward, libs/modeling/head.py:95-108."""
    out = []
    for l, (x, mask) in enumerate(zip(fpn, fpn_masks)):
        for i in range(n_layers):
            x, _ = masked_conv1d(x, mask, sd[f'{pre}convs.{i}.conv.weight'], None, 1, 1)
            x = F.relu(layer_norm_c(x, sd[f'{pre}norms.{i}.weight'], sd[f'{pre}norms.{i}.bias']))
        off, _ = masked_conv1d(x, mask, sd[pre + 'reg_head.conv.weight'], sd[pre + 'reg_head.conv.bias'], 1, 1)
        off = F.relu(off * sd[f'{pre}scales.{l}.scale'].to(off.dtype))
        out.append(off.transpose(1, 2))
    return out


def tcn_forward(sd, pre, x, mask, n_layers):
    """TCN.forward, libs/modeling/tcn.py:66-84 with DilatedResidualLayer :21-38
    (dilation 2^i, nn.LayerNorm over channels, dropout off)."""
    out = F.conv1d(x, sd[pre + 'conv_1x1.weight'], sd[pre + 'conv_1x1.bias'])
    mf = mask[:, 0:1, :].to(x.dtype)
    for i in range(n_layers):
        p = f'{pre}layers.{i}.'
        dil = 2 ** i
        o = F.relu(F.conv1d(out, sd[p + 'conv_dilated.weight'], sd[p + 'conv_dilated.bias'],
                            padding=dil, dilation=dil))
        o = F.conv1d(o, sd[p + 'conv_1x1.weight'], sd[p + 'conv_1x1.bias'])
        out = (out + o) * mf
        out = F.layer_norm(out.permute(0, 2, 1), (out.size(1),), sd[p + 'norm.weight'],
                           sd[p + 'norm.bias'], 1e-5).permute(0, 2, 1)
    out = F.conv1d(out, sd[pre + 'conv_out.weight'], sd[pre + 'conv_out.bias'])
    return out * mf


def fuse_and_predict(sd, opt, fpn, fpn_masks):
    """PtTransformerEarlyFusionIterative.fuse_and_predict with second_fusion=False,
    libs/modeling/model.py:442-471."""
    nl = opt['model']['cls_head']['n_layers']
    logits1 = cls_head_forward(sd, 'cls_head.', fpn, fpn_masks, nl)
    ref_len = logits1[0].shape[1]
    expand = [logits1[0]]
    for l in logits1[1:]:
        e = F.interpolate(l.unsqueeze(1), size=ref_len, mode='nearest')[:, 0]
        expand.append(e * fpn_masks[0][:, 0].to(e.dtype))
    expand = torch.stack(expand, dim=1)
    n_levels = len(fpn)
    refine = tcn_forward(sd, 'refine.', expand, fpn_masks[0], n_levels)
    new_fpn = []
    for i, f in enumerate(fpn):
        if i != 0:
            refine = masked_max_pool1d(refine, fpn_masks[i - 1])[0]
        new_fpn.append(torch.cat([f, refine], dim=1))
    logits2 = cls_head_forward(sd, 'cls_head2.', new_fpn, fpn_masks, nl)
    offsets = reg_head_forward(sd, 'reg_head.', new_fpn, fpn_masks, opt['model']['reg_head']['n_layers'])
    return logits1, logits2, offsets, [m.squeeze(1) for m in fpn_masks]


# ----------------------------------------------------------------------------- saliency
def saliency_scores(shallow_vid, text_cls, norm):
    """libs/modeling/model.py:500-505.  shallow (1, Cs, T), text_cls (n, Cs) -> (n, T)."""
    if norm:
        v = shallow_vid / (shallow_vid.norm(dim=1, keepdim=True) + 1e-4)
        t = text_cls / (text_cls.norm(dim=1, keepdim=True) + 1e-4)
    else:
        v, t = shallow_vid, text_cls
    return torch.einsum('bht,bh->bt', v.expand(t.size(0), -1, -1), t)


def select_clips(correl_row, vid_len, sn, sratio):
    """Top-k block selection for one query, libs/modeling/model.py:531-541.

    correl_row (T,) float; returns (pooled (M,), selected block ids (sorted), weight (vid_len,)
    bool).  Restated without torch ops so every quirk is explicit:
      * block mean over the *valid* count of the last partial block (ceil_mode avg_pool1d);
      * k = int(sratio * M) in Python double; k == 0 selects ALL blocks (slice [-0:]);
      * ties: the reference's argsort is unstable; this oracle uses a stable ascending sort
        (equal scores: higher block index ranks higher) — only tie-free inputs are compared
        against the reference;
      * nearest up-sampling index = min(int(floorf(i * (float)M / (float)len)), M - 1) in fp32.
    """
    x = correl_row[:vid_len].detach().to(torch.float32).numpy()
    m = (vid_len + sn - 1) // sn
    pooled = np.zeros(m, dtype=np.float32)
    for j in range(m):
        seg = x[j * sn:min((j + 1) * sn, vid_len)]
        acc = np.float32(0)
        for val in seg:
            acc = np.float32(acc + val)
        pooled[j] = np.float32(acc / np.float32(len(seg)))
    k = int(sratio * m)
    order = np.argsort(pooled, kind='stable')
    sel = order if k == 0 else order[-k:]
    w = np.zeros(m, dtype=bool)
    w[sel] = True
    scale = np.float32(np.float32(m) / np.float32(vid_len))
    idx = np.minimum(np.floor(np.arange(vid_len, dtype=np.float32) * scale).astype(np.int64), m - 1)
    return pooled, np.sort(sel), torch.from_numpy(w[idx])


# ----------------------------------------------------------------------------- model
def grounder_forward(sd, opt, vid, shallow_vid, vid_masks, text_list, text_cls,
                     text_mask_list, return_aux=False):
    """PtTransformerEarlyFusionIterative._drop_forward_eval, libs/modeling/model.py:480-565.

    vid (1, Ce, T), shallow_vid (1, Cs, T), vid_masks (1, T) bool, text_list: n x (1, Ct, Lq+1)
    (already encoded), text_cls (n, Cs).  Returns per-query lists of per-level
    logits (1, T_l), offsets (1, T_l, 2), masks (1, T_l)."""
    m = opt['model']
    assert vid.size(0) == 1
    correl = saliency_scores(shallow_vid, text_cls, m['norm'])
    vid_len = int(vid_masks.sum())
    out_logits, out_offsets, out_masks, aux = [], [], [], []
    for b, (text, text_masks) in enumerate(zip(text_list, text_mask_list)):
        pooled, sel, weight = select_clips(correl[b], vid_len, m['sn'], m['sratio'])
        all_weight = torch.zeros_like(vid_masks)
        all_weight[0, :vid_len] = weight
        v = vid * all_weight.unsqueeze(1).to(vid.dtype)
        masks = vid_masks
        if not m['msf']:
            masks = torch.logical_and(all_weight, vid_masks)
        elif m['sfonly']:
            v = shallow_vid
        else:
            v = torch.cat([v, shallow_vid], dim=1)
        if m['scat']:
            v = torch.cat([v, correl[b][None, None, :]], dim=1)
        masks = masks.unsqueeze(1)
        xm, masks = masked_conv1d(v, masks, sd['vid_map.conv.weight'], sd['vid_map.conv.bias'])
        x, masks = fusion_forward(sd, opt, xm, masks, text, text_masks)
        fpn, fpn_masks = video_net_forward(sd, opt, x, masks)
        l1, l2, o, mk = fuse_and_predict(sd, opt, fpn, fpn_masks)
        out_logits.append(l2)
        out_offsets.append(o)
        out_masks.append(mk)
        if return_aux:
            aux.append(dict(correl=correl[b], pooled=pooled, sel=sel, weight=all_weight[0],
                            vid_map=xm, fusion=x, fpn=fpn, logits1=l1))
    if return_aux:
        return out_logits, out_offsets, out_masks, aux
    return out_logits, out_offsets, out_masks


def fpn_points(n_points_per_level):
    """PtGenerator (libs/modeling/model.py:703-743): eval uses only columns 0 (coordinate
    t * 2^l) and 3 (stride 2^l)."""
    pts = []
    for l, n in enumerate(n_points_per_level):
        stride = float(2 ** l)
        c = torch.arange(n, dtype=torch.float32) * stride
        pts.append(torch.stack([c, torch.zeros(n), torch.zeros(n), torch.full((n,), stride)], dim=1))
    return pts


# ----------------------------------------------------------------------------- decode / NMS
def collect_segments(logits, offsets, masks, pre_nms_thresh, pre_nms_topk, seg_len_thresh,
                     scores_override=None):
    """Evaluator._collect_segments, libs/worker_v2.py:1131-1187 (ext_scores None).

    Per-level lists for ONE query.  Candidate order: level-major concatenation, then
    descending score; this oracle breaks score ties by ascending flat index (the reference's
    argsort is unstable).  Returns segs (k,2), scores (k,), flat point idx (k,).
    """
    pts, sc, off = [], [], []
    base = 0
    for l, (lg, of, mk) in enumerate(zip(logits, offsets, masks)):
        lg, of, mk = lg[0], of[0], mk[0]
        s = torch.sigmoid(lg) if scores_override is None else scores_override[l]
        s = s * mk.float()
        idx = s > pre_nms_thresh
        n = lg.numel()
        coords = torch.arange(n, dtype=torch.float32) * float(2 ** l)
        flat = torch.arange(n) + base
        base += n
        pts.append(torch.stack([coords[idx], torch.full((int(idx.sum()),), float(2 ** l)), flat[idx].float()], 1))
        sc.append(s[idx])
        off.append(of[idx])
    pts, sc, off = torch.cat(pts), torch.cat(sc), torch.cat(off)
    n_topk = min(len(pts), pre_nms_topk)
    order = torch.sort(sc, descending=True, stable=True)[1][:n_topk]
    pts, sc, off = pts[order], sc[order], off[order]
    left = pts[:, 0] - off[:, 0] * pts[:, 1]
    right = pts[:, 0] + off[:, 1] * pts[:, 1]
    keep = (right - left) > seg_len_thresh
    segs = torch.stack((left, right), dim=-1)[keep]
    return segs, sc[keep], pts[keep, 2].long()


def soft_nms_np(segs, scores, iou_thresh, sigma, min_score, method, max_iters=None):
    """softnms_1d_cpu, libs/nms/src/nms_cpu.cpp:72-172, statement by statement in numpy
    float32 (pure-Python loops: small cases only; oracle/nms_oracle.c is the fast twin).
    Returns dets (n_out, 3) and original indices (n_out,) where n_out = number of outer
    steps executed (bounded by ``max_iters`` when given)."""
    f = np.float32
    x1 = segs[:, 0].astype(f).copy()
    x2 = segs[:, 1].astype(f).copy()
    sc = scores.astype(f).copy()
    areas = (x2 - x1 + f(1e-6)).astype(f)
    n = len(sc)
    inds = np.arange(n)
    dets = np.zeros((n, 3), dtype=f)
    i = 0
    while i < n and (max_iters is None or i < max_iters):
        mp = i
        ms = sc[i]
        for pos in range(i + 1, n):
            if ms < sc[pos]:
                ms, mp = sc[pos], pos
        ix1, ix2, isc, iar, iin = x1[mp], x2[mp], sc[mp], areas[mp], inds[mp]
        dets[i] = (ix1, ix2, isc)
        x1[mp], x2[mp], sc[mp], areas[mp], inds[mp] = x1[i], x2[i], sc[i], areas[i], inds[i]
        x1[i], x2[i], sc[i], areas[i], inds[i] = ix1, ix2, isc, iar, iin
        pos = i + 1
        while pos < n:
            xx1 = max(ix1, x1[pos])
            xx2 = min(ix2, x2[pos])
            inter = max(f(0), f(xx2 - xx1))
            ovr = f(inter / f(f(iar + areas[pos]) - inter))
            w = f(1)
            if method == 0:
                if ovr >= f(iou_thresh):
                    w = f(0)
            elif method == 1:
                if ovr >= f(iou_thresh):
                    w = f(f(1) - ovr)
            elif method == 2:
                w = f(np.exp(f(-f(ovr * ovr) / f(sigma))))
            sc[pos] = f(sc[pos] * w)
            if sc[pos] < f(min_score):
                x1[pos], x2[pos], sc[pos], areas[pos], inds[pos] = \
                    x1[n - 1], x2[n - 1], sc[n - 1], areas[n - 1], inds[n - 1]
                n -= 1
                pos -= 1
            pos += 1
        i += 1
    return dets[:i], inds[:i]


def hard_nms_np(segs, scores, iou_thresh):
    """nms_1d_cpu, libs/nms/src/nms_cpu.cpp:20-63; stable descending sort."""
    f = np.float32
    x1, x2 = segs[:, 0].astype(f), segs[:, 1].astype(f)
    areas = (x2 - x1 + f(1e-6)).astype(f)
    order = np.argsort(-scores.astype(f), kind='stable')
    n = len(order)
    sel = np.ones(n, dtype=bool)
    for _i in range(n):
        if not sel[_i]:
            continue
        i = order[_i]
        for _j in range(_i + 1, n):
            if not sel[_j]:
                continue
            j = order[_j]
            inter = max(f(0), f(min(x2[i], x2[j]) - max(x1[i], x1[j])))
            ovr = f(inter / f(f(areas[i] + areas[j]) - inter))
            if ovr >= f(iou_thresh):
                sel[_j] = False
    return order[sel]


def segment_voting(nms_segs, all_segs, all_scores, iou_thresh):
    """libs/nms/nms.py:64-103."""
    a = nms_segs[:, None]
    b = all_segs[None, :]
    left = torch.maximum(a[..., 0], b[..., 0])
    right = torch.minimum(a[..., 1], b[..., 1])
    overlap = (right - left).clamp(min=0)
    union = (a[..., 1] - a[..., 0]) + (b[..., 1] - b[..., 0]) - overlap
    iou = overlap / union
    w = (iou >= iou_thresh).float() * all_scores[None]
    w = w / torch.sum(w, dim=1, keepdim=True)
    return w @ all_segs


def batched_nms(segs, scores, iou_thresh, min_score, max_num_segs, mode='soft_nms',
                sigma=0.5, voting_thresh=0.75, softnms_fn=None, nms_fn=None):
    """libs/nms/nms.py:106-148 (+ NMSop :6-31, SoftNMSop :34-61).  ``softnms_fn`` /
    ``nms_fn`` let the caller plug the compiled C twin (oracle/nms_oracle.c) or the compiled
    reference extension (oracle/_ref) in place of the numpy loops."""
    if len(segs) == 0:
        return torch.zeros(0, 2), torch.zeros(0)
    if mode is not None:
        if mode == 'nms':
            s, c = segs, scores
            if min_score > 0:
                keep = c > min_score
                s, c = s[keep], c[keep]
            if nms_fn is not None:
                idx = nms_fn(s.numpy(), c.numpy(), float(iou_thresh))
            else:
                idx = hard_nms_np(s.numpy(), c.numpy(), float(iou_thresh))
            if max_num_segs > 0:
                idx = idx[:min(max_num_segs, len(idx))]
            idx = torch.as_tensor(np.asarray(idx), dtype=torch.long)
            nms_segs, nms_scores = s[idx].contiguous(), c[idx].contiguous()
        elif mode == 'soft_nms':
            iters = max_num_segs if max_num_segs > 0 else None
            if softnms_fn is not None:
                dets, _ = softnms_fn(segs.numpy(), scores.numpy(), float(iou_thresh), float(sigma),
                                     float(min_score), 2, iters)
            else:
                dets, _ = soft_nms_np(segs.numpy(), scores.numpy(), float(iou_thresh), float(sigma),
                                      float(min_score), 2, iters)
            dets = torch.from_numpy(np.ascontiguousarray(dets))
            nms_segs, nms_scores = dets[:, :2].contiguous(), dets[:, 2].contiguous()
        else:
            raise NotImplementedError('invalid NMS mode')
        if voting_thresh > 0:
            nms_segs = segment_voting(nms_segs, segs, scores, voting_thresh)
    else:
        nms_segs, nms_scores = segs, scores
    idx = torch.sort(nms_scores, descending=True, stable=True)[1]
    k = min(max_num_segs, len(nms_segs))
    return nms_segs[idx[:k]], nms_scores[idx[:k]]


def generate_proposals(opt, data, logits_list, offsets_list, masks_list, softnms_fn=None, nms_fn=None):
    """Evaluator._generate_proposals, libs/worker_v2.py:1063-1129 (window_offset 0)."""
    ev = opt['eval']
    vid_stride = opt['model'].get('vid_stride', 1)
    results, cands = [], []
    for lg, of, mk in zip(logits_list, offsets_list, masks_list):
        segs, scores, idx = collect_segments(lg, of, mk, ev['pre_nms_thresh'], ev['pre_nms_topk'],
                                             ev['seg_len_thresh'])
        cands.append((segs, scores, idx))
        s, c = batched_nms(segs, scores, softnms_fn=softnms_fn, nms_fn=nms_fn, **opt['nms'])
        if len(s) > 0:
            s = s * vid_stride
            s = (s * data['clip_stride'] + 0.5 * data['clip_size']) / data['fps']
            s = torch.clamp(s, min=0, max=data['duration'])
        results.append({'segments': s, 'scores': c})
    return results, cands


def predict(sd, opt, data, dtype=torch.float32, softnms_fn=None, nms_fn=None, return_aux=False):
    """Evaluator.simple_predict minus the loss statistics: _forward (libs/worker_v2.py:
    930-1026: per-query text encode, pad to input_vid_len, model call) + _generate_proposals."""
    m = opt['model']
    sd = {k: v.to(dtype) if v.is_floating_point() else v for k, v in sd.items()}
    text_list, text_mask_list = [], []
    for text in data['text']:
        text = text[None].to(dtype)
        tm = torch.ones((1, 1, text.size(-1)), dtype=torch.bool)
        t, tmk = text_net_forward(sd, opt, text, tm)
        text_list.append(t)
        text_mask_list.append(tmk)
    vid, shallow = data['vid'], data['shallow_vid']
    vid_len = vid.size(-1)
    input_vid_len = padded_len(opt, vid_len)
    window = F.pad(vid, (0, input_vid_len - vid_len))[None].to(dtype)
    shallow_window = F.pad(shallow, (0, input_vid_len - vid_len))[None].to(dtype)
    window_mask = torch.arange(input_vid_len).view(1, -1) < vid_len
    out = grounder_forward(sd, opt, window, shallow_window, window_mask, text_list,
                           data['text_cls'].to(dtype), text_mask_list, return_aux=return_aux)
    logits, offsets, masks = out[:3]
    logits = [[x.float() for x in q] for q in logits]
    offsets = [[x.float() for x in q] for q in offsets]
    results, cands = generate_proposals(opt, data, logits, offsets, masks, softnms_fn, nms_fn)
    ret = dict(logits=logits, offsets=offsets, masks=masks, results=results, cands=cands,
               text=text_list)
    if return_aux:
        ret['aux'] = out[3]
    return ret


def eval_loss(opt, data, logits_list, offsets_list, masks_list, regression_range):
    """Evaluator._calc_loss, libs/worker_v2.py:1029-1061, with annotate_points_per_video (:93-133), calc_focal_loss /
    calc_iou_loss (:85-91: smoothing 0.2, alpha 0.5, reg_loss 'iou' -> ctr_giou_loss) and sigmoid_focal_loss / ctr_giou_loss
    (libs/modeling/loss.py:6-108).  regression_range: PtGenerator.regression_range (one (lo, hi) per level)."""
    tr = opt['train']
    cs, rad = tr.get('center_sampling', 'radius'), tr['center_sampling_radius']
    sizes = [int(x.size(-1)) for x in logits_list[0]]
    pts = torch.cat([torch.stack((torch.arange(n, dtype=torch.float32) * 2 ** l, torch.full((n,), float(regression_range[l][0])),
                                  torch.full((n,), float(regression_range[l][1])), torch.full((n,), float(2 ** l))), 1)
                     for l, n in enumerate(sizes)])
    targets = torch.as_tensor(np.asarray(data['target']), dtype=torch.float32).reshape(-1, 2) / opt['model'].get('vid_stride', 1)
    cls_l, reg_l = [], []
    for i, target in enumerate(targets):
        pt2start, pt2end = pts[:, 0] - target[0], target[1] - pts[:, 0]
        gt_off = torch.stack((pt2start, pt2end), -1) / pts[:, 3:]
        if cs == 'radius':
            ctr = 0.5 * (target[0] + target[1])
            radius = pts[:, 3] * rad
            t_min, t_max = (ctr - radius).clamp(min=target[0]), (ctr + radius).clamp(max=target[1])
            inside = torch.logical_and(pts[:, 0] - t_min > 0, t_max - pts[:, 0] > 0)
        else:
            inside = torch.logical_and(pt2start > 0, pt2end > 0)
        md = torch.maximum(pt2start, pt2end)
        labels = torch.logical_and(inside, torch.logical_and(md >= pts[:, 1], md < pts[:, 2]))[None]
        logits = torch.cat(list(logits_list[i]), dim=1).float()
        offsets = torch.cat(list(offsets_list[i]), dim=1).float()
        masks = torch.cat(list(masks_list[i]), dim=1)
        pos = torch.logical_and(labels, masks)
        norm = max(int(pos.sum()), 1)
        x, t = logits[masks], labels[masks].float() * 0.8 + 0.1
        p = torch.sigmoid(x)
        p_t = p * t + (1 - p) * (1 - t)
        ce = F.binary_cross_entropy_with_logits(x, t, reduction='none')
        cls = (0.5 * ce * (1 - p_t) ** 2).sum() / norm
        po, go = offsets[pos], gt_off[None][pos]
        inter = torch.min(po[:, 0], go[:, 0]) + torch.min(po[:, 1], go[:, 1])
        union = (po[:, 0] + po[:, 1]) + (go[:, 0] + go[:, 1]) - inter
        reg = (1.0 - inter / union.clamp(min=1e-8)).sum() / norm
        cls_l.append(float(cls))
        reg_l.append(float(reg))
    return {'cls_loss': float(np.nanmean(cls_l)), 'reg_loss': float(np.nanmean(reg_l))}


def min_chunk_size(opt):
    """libs/worker_v2.py:769-781."""
    m = opt['model']
    mcs = 1
    for l in range(m['num_fpn_levels']):
        s = 2 ** l
        if m['mha_win_size'] > 0:
            s *= (m['mha_win_size'] // 2) * 2
        mcs = max(mcs, s)
    return mcs


def padded_len(opt, vid_len):
    """libs/worker_v2.py:969-976."""
    m = opt['model']
    vs = m.get('vid_stride', 1)
    input_vid_len = m['max_vid_len'] * vs
    if vid_len > input_vid_len:
        stride = min_chunk_size(opt) * vs
        input_vid_len = (vid_len + (stride - 1)) // stride * stride
    return input_vid_len

"""Generate golden fixtures by running the UNMODIFIED reference in the build container.

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

The reference (ZijiaLewisLu/CVPR2025-DeCafNet) is imported in place from /root/reference
(read-only; nothing is copied).  Import-time stubs cover the three packages that are absent
here and untouched on the eval path (yacs, torchtext, decord — SURVEY.md Appendix C); the
reference's own C++ NMS extension is compiled as-is by oracle/build_ref.py and imported as
``nms_1d_cpu_vg``.  Weights and inputs come from decaf_b200.synth (deterministic by name), so
a fixture only stores the config, the seeds and the reference's OUTPUTS.

Calling convention: per query ``model.encode_text``; ``model(...,
eval=True)``; ``PtGenerator``; ``Evaluator._generate_proposals`` on an Evaluator built with
``__new__`` (its __init__ needs dataset files).
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(ROOT, 'cvpr2025-decafnet_b200'))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.nn.functional as F

from decaf_b200 import synth
from oracle import build_ref

REF = '/root/reference'


def import_reference():
    """Returns (model_mod, worker_mod) of the reference with stubs in place."""
    ref_nms = build_ref.build_reference_nms()
    sys.path.insert(0, os.path.dirname(ref_nms))
    sys.path.insert(0, REF)
    tt = types.ModuleType('torchtext')
    tt.data = types.ModuleType('torchtext.data')
    tt.data.get_tokenizer = lambda *a, **k: None
    tt.vocab = types.ModuleType('torchtext.vocab')
    sys.modules.update({'torchtext': tt, 'torchtext.data': tt.data, 'torchtext.vocab': tt.vocab})
    dc = types.ModuleType('decord')
    dc.bridge = types.SimpleNamespace(set_bridge=lambda *a, **k: None)
    sys.modules['decord'] = dc
    yc = types.ModuleType('yacs')
    yc.config = types.ModuleType('yacs.config')
    yc.config.CfgNode = synth.AttrDict
    sys.modules.update({'yacs': yc, 'yacs.config': yc.config})
    for name in ('wandb', 'cv2'):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    import libs.modeling.model as model_mod
    import libs.worker_v2 as worker_mod
    return model_mod, worker_mod


def build_reference_model(model_mod, opt, seed):
    ropt = opt.clone()          # the ctor mutates cls_head/reg_head.embd_dim (model.py:426-428)
    model = model_mod.PtTransformerEarlyFusionIterative(ropt, second_fusion=False).eval()
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd = synth.fill_state_dict(shapes, seed)
    model.load_state_dict(sd)
    model.requires_grad_(False)
    return model, shapes


def padded_len(opt, vid_len):
    m = opt.model
    mcs = 1
    for l in range(m.num_fpn_levels):
        s = 2 ** l
        if m.mha_win_size > 0:
            s *= (m.mha_win_size // 2) * 2
        mcs = max(mcs, s)
    ivl = m.max_vid_len
    if vid_len > ivl:
        ivl = (vid_len + mcs - 1) // mcs * mcs
    return ivl


@torch.no_grad()
def run_reference(model_mod, worker_mod, opt, data, seed):
    model, shapes = build_reference_model(model_mod, opt, seed)
    text_list, mask_list = [], []
    for text in data['text']:
        text = text[None]
        tm = text.new_full((1, 1, text.size(-1)), 1, dtype=torch.bool)
        t, m = model.encode_text(text, tm)
        text_list.append(t)
        mask_list.append(m)
    vid_len = data['vid'].size(-1)
    T = padded_len(opt, vid_len)
    window = F.pad(data['vid'], (0, T - vid_len))[None]
    shallow = F.pad(data['shallow_vid'], (0, T - vid_len))[None]
    mask = torch.arange(T).view(1, -1) < vid_len
    hooks, cap = [], {}
    hooks.append(model.vid_map.register_forward_hook(lambda m, i, o: cap.setdefault('vid_map', []).append(o[0])))
    hooks.append(model.fusion.register_forward_hook(lambda m, i, o: cap.setdefault('fusion', []).append(o[0])))
    hooks.append(model.cls_head.register_forward_hook(lambda m, i, o: cap.setdefault('logits1', []).append(o[0])))
    hooks.append(model.refine.register_forward_hook(lambda m, i, o: cap.setdefault('refine', []).append(o)))
    logits, offsets, masks = model(window, shallow, mask, text_list, data['text_cls'], mask_list, eval=True)
    for h in hooks:
        h.remove()
    n_levels = opt.model.num_fpn_levels
    pt_gen = model_mod.PtGenerator(max_seq_len=max(T, opt.model.max_vid_len) * 2 ** 0,
                                   num_fpn_levels=n_levels, regression_range=4, sigma=0.5)
    pts = pt_gen([m.size(-1) for m in masks[0]])
    ev = worker_mod.Evaluator.__new__(worker_mod.Evaluator)
    ev.pre_nms_thresh = opt.eval.pre_nms_thresh
    ev.pre_nms_topk = opt.eval.pre_nms_topk
    ev.seg_len_thresh = opt.eval.seg_len_thresh
    ev.vid_stride = 1
    from collections import defaultdict
    ev.time_dict = defaultdict(list)
    ev.batched_nms = lambda segs, scores: worker_mod.batched_nms(segs, scores, **opt.nms)
    cands = [ev._collect_segments(pts, lg, of, mk, None) for lg, of, mk in zip(logits, offsets, masks)]
    results = ev._generate_proposals(
        {k: data[k] for k in ('clip_stride', 'clip_size', 'fps', 'duration')},
        [logits, offsets, pts, masks])
    # eval-time loss statistics of the reference (worker_v2.py:1029-1061); its `.cuda()` calls are no-ops on this CPU-only box
    ev.opt = opt
    _cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        loss = ev._calc_loss({'target': data['target']}, [logits, offsets, pts, masks])
    finally:
        torch.Tensor.cuda = _cuda
    # saliency internals recomputed with the reference's own five lines (model.py:500-541)
    m = opt.model
    if m.norm:
        v = shallow / (shallow.norm(dim=1, keepdim=True) + 1e-4)
        t = data['text_cls'] / (data['text_cls'].norm(dim=1, keepdim=True) + 1e-4)
        correl = torch.einsum('bht,bh->bt', v, t)
    else:
        correl = torch.einsum('bht,bh->bt', shallow, data['text_cls'])
    weights = []
    for b in range(len(text_list)):
        cb = F.avg_pool1d(correl[b, None, :vid_len], kernel_size=m.sn, stride=m.sn, ceil_mode=True)[0]
        ranked = cb.argsort()
        topk = ranked[-int(m.sratio * cb.shape[0]):]
        w = torch.zeros_like(cb)
        w[topk] = 1
        w = F.interpolate(w[None, None, :], size=vid_len, mode='nearest')[0, 0]
        aw = torch.zeros(T)
        aw[:vid_len] = w
        weights.append(aw)
    out = {'T': np.int64(T), 'vid_len': np.int64(vid_len), 'n_query': np.int64(len(text_list)),
           'correl': correl.numpy(), 'weight': torch.stack(weights).numpy().astype(np.uint8),
           'loss_cls': np.float64(loss['cls_loss']), 'loss_reg': np.float64(loss['reg_loss']),
           'regression_range': np.asarray(pt_gen.regression_range, dtype=np.float64)}
    for b in range(len(text_list)):
        out[f'text{b}'] = text_list[b][0].numpy()
        out[f'logits{b}'] = torch.cat([x[0] for x in logits[b]]).numpy()
        out[f'offsets{b}'] = torch.cat([x[0] for x in offsets[b]]).numpy()
        out[f'masks{b}'] = torch.cat([x.reshape(-1) for x in masks[b]]).numpy().astype(np.uint8)
        out[f'logits1_{b}'] = torch.cat([x[0] for x in cap['logits1'][b]]).numpy()
        out[f'vid_map{b}'] = cap['vid_map'][b][0].numpy()
        out[f'fusion{b}'] = cap['fusion'][b][0].numpy()
        out[f'refine{b}'] = cap['refine'][b][0].numpy()
        out[f'cand_segs{b}'] = cands[b][0].numpy()
        out[f'cand_scores{b}'] = cands[b][1].numpy()
        out[f'res_segs{b}'] = results[b]['segments'].numpy()
        out[f'res_scores{b}'] = results[b]['scores'].numpy()
    return out, shapes


# name -> (opt kwargs on top of synth.tiny_opt, vid_len, n_query, seed)
CASES = {
    'tiny_msf':      (dict(), 50, 3, 11),
    'tiny_nomsf':    (dict(msf=False, sratio=0.5), 61, 2, 12),
    'tiny_scat':     (dict(scat=True, norm=False, sratio=0.2), 64, 2, 13),
    'tiny_long':     (dict(sratio=0.3, text_abs_pe=True), 150, 2, 14),   # > max_seq_len: padded to 160, PE interpolated
    'tiny_hardnms':  (dict(nms_mode='nms', iou_thresh=0.5, voting_thresh=0.0, sratio=0.0), 40, 2, 15),
    'small_w9':      (dict(embd_dim=96, n_levels=5, win=9, max_seq_len=128, sn=10, n_heads=4,
                           text_dim=64, vid_in_dim=56, pre_nms_topk=200), 117, 3, 16),
}


def case_opt(name):
    kw = CASES[name][0]
    return synth.tiny_opt(**kw)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    model_mod, worker_mod = import_reference()
    for name, (kw, vid_len, nq, seed) in CASES.items():
        opt = case_opt(name)
        data = synth.synth_video(opt, vid_len, nq, seed=seed, tag=name, text_len_range=(3, 12), n_events=1)
        out, shapes = run_reference(model_mod, worker_mod, opt, data, seed)
        out['state_keys'] = np.array(sorted(shapes))
        out['state_shapes'] = np.array([','.join(map(str, shapes[k])) for k in sorted(shapes)])
        path = os.path.join(HERE, f'{name}.npz')
        np.savez_compressed(path, **out)
        print(name, 'T', int(out['T']), 'cands', [len(out[f'cand_scores{b}']) for b in range(nq)],
              'res', [out[f'res_scores{b}'].round(3).tolist() for b in range(nq)],
              os.path.getsize(path) // 1024, 'KiB')


if __name__ == '__main__':
    main()

"""Developer diagnostic (GPU box): per-stage error of the CUDA path vs the oracle, both dtypes.
    python tools/diag_parity.py [case ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'cvpr2025-decafnet_b200'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import numpy as np
import torch

from decaf_b200 import synth
from decaf_b200.worker_v2 import Evaluator, create_model
from oracle import grounder_oracle as go


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-12)


def rms(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.sqrt(((a - b) ** 2).mean()) / max(np.sqrt((b ** 2).mean()), 1e-12)


def run(name, opt, sd, data, dtypes=(torch.float32, torch.bfloat16)):
    ref = go.predict(sd, opt, data, return_aux=True)
    nq = len(ref['logits'])
    for dt in dtypes:
        ev = Evaluator(opt.clone(), dataset=[], state_dict=sd, act_dtype=dt, gemm_impl=0)
        eng = ev.model.engine()
        eng.capture = {}
        outputs, results, _ = ev.simple_predict(data)
        logits, offsets, pts, masks = outputs
        cap = eng.capture
        p = eng.plan(nq, cap['mask0'].shape[1])
        print(f'== {name} {dt} T={p.T} lens={p.lens}')
        for b in range(nq):
            aux = ref['aux'][b]
            v0 = cap['mask0'][b].cpu().numpy() > 0
            line = [f'q{b}']
            line.append('sel_diff %d' % int((cap['sel'][b].cpu().numpy() != aux['weight'].numpy()).sum()))
            vm = cap['vid_map'][b].cpu().numpy().T
            line.append('vid_map %.1e' % rel(vm[:, v0], aux['vid_map'][0].numpy()[:, v0]))
            fu = cap['fusion'][b].cpu().numpy().T
            line.append('fusion %.1e/%.1e' % (rel(fu[:, v0], aux['fusion'][0].numpy()[:, v0]), rms(fu[:, v0], aux['fusion'][0].numpy()[:, v0])))
            for l in range(len(p.lens)):
                m = ref['masks'][b][l][0].numpy()
                f = cap[f'fpn{l}'][b].cpu().numpy().T
                r = aux['fpn'][l][0].numpy()
                line.append('fpn%d %.1e/%.1e' % (l, rel(f[:, m], r[:, m]), rms(f[:, m], r[:, m])))
            l1 = cap['logits1'][b].cpu().numpy()
            for l in range(len(p.lens)):
                a = l1[p.off[l]:p.off[l] + p.lens[l]]
                line.append('l1_%d %.1e' % (l, rel(a, aux['logits1'][l][0].numpy())))
            for l in range(len(p.lens)):
                m = ref['masks'][b][l][0].numpy()
                line.append('lg%d %.1e/%.1e of%d %.1e' % (
                    l, rel(logits[b][l][0].cpu().numpy()[m], ref['logits'][b][l][0].numpy()[m]),
                    rms(logits[b][l][0].cpu().numpy()[m], ref['logits'][b][l][0].numpy()[m]),
                    l, rel(offsets[b][l][0].cpu().numpy()[m], ref['offsets'][b][l][0].numpy()[m])))
            print('  ' + ' | '.join(line))
            rs, os_ = results[b]['segments'].numpy(), ref['results'][b]['segments'].numpy()
            print('    segs', rs.shape, os_.shape, (np.abs(rs - os_).max() if rs.shape == os_.shape else 'shape!'))


if __name__ == '__main__':
    from golden_util import CASES, load_case
    names = sys.argv[1:] or ['tiny_msf', 'small_w9', 'fresh']
    for name in names:
        if name == 'fresh':
            opt = synth.tiny_opt(embd_dim=128, n_levels=5, win=9, max_seq_len=256, sn=12, vid_in_dim=64, text_dim=64)
            shapes = {k: tuple(v.shape) for k, v in create_model(opt.clone()).state_dict().items()}
            sd = synth.fill_state_dict(shapes, 5)
            data = synth.synth_video(opt, 256, 4, seed=256, tag='fresh', n_events=1)
        elif name == 'nlq':
            opt = synth.nlq_opt()
            shapes = {k: tuple(v.shape) for k, v in create_model(opt.clone()).state_dict().items()}
            sd = synth.fill_state_dict(shapes, 2022)
            data = synth.synth_video(opt, 2000, 2, seed=2022, tag='nlq', n_events=1)
        else:
            opt, sd, data, g = load_case(name)
        run(name, opt, sd, data)

"""Helpers shared by the golden-fixture tests (CPU oracle tests and GPU parity tests)."""
import os

import numpy as np
import torch

from decaf_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

# must mirror tests/golden/make_golden.py:CASES (kept literal so the tests do not import the
# generator, which needs /root/reference)
CASES = {
    'tiny_msf':      (dict(), 50, 3, 11),
    'tiny_nomsf':    (dict(msf=False, sratio=0.5), 61, 2, 12),
    'tiny_scat':     (dict(scat=True, norm=False, sratio=0.2), 64, 2, 13),
    'tiny_long':     (dict(sratio=0.3, text_abs_pe=True), 150, 2, 14),
    'tiny_hardnms':  (dict(nms_mode='nms', iou_thresh=0.5, voting_thresh=0.0, sratio=0.0), 40, 2, 15),
    'small_w9':      (dict(embd_dim=96, n_levels=5, win=9, max_seq_len=128, sn=10, n_heads=4,
                           text_dim=64, vid_in_dim=56, pre_nms_topk=200), 117, 3, 16),
}


def load_case(name):
    kw, vid_len, nq, seed = CASES[name]
    opt = synth.tiny_opt(**kw)
    g = np.load(os.path.join(GOLDEN_DIR, f'{name}.npz'))
    shapes = {k: tuple(int(x) for x in s.split(',') if x) for k, s in zip(g['state_keys'], g['state_shapes'])}
    sd = synth.fill_state_dict(shapes, seed)
    data = synth.synth_video(opt, vid_len, nq, seed=seed, tag=name, text_len_range=(3, 12), n_events=1)
    return opt, sd, data, g


def level_sizes(T, n_levels):
    return [T // 2 ** l for l in range(n_levels)]

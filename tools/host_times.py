#!/usr/bin/env python
"""Where the host thread spends its time in Evaluator.predict_videos (NLQ shape): staging copies, launch, result harvest."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, 'cvpr2025-decafnet_b200')):
    if _p not in sys.path:
        sys.path.insert(0, _p)
import torch
from decaf_b200 import synth
from decaf_b200.worker_v2 import Evaluator, create_model

opt = synth.nlq_opt()
shapes = {k: tuple(v.shape) for k, v in create_model(opt.clone()).state_dict().items()}
sd = synth.fill_state_dict(shapes, 2022)
videos = [synth.synth_video(opt, 2000, 16, seed=2022 + i, tag=f'v{i}', n_events=1) for i in range(8)]
ev = Evaluator(opt.clone(), dataset=videos, state_dict=sd)
for _ in ev.predict_videos(videos * 2):
    pass
torch.cuda.synchronize()
acc = {'stage': 0.0, 'upload': 0.0, 'launch': 0.0, 'harvest': 0.0}
orig_stage, orig_up, orig_run, orig_res = ev._stage_host, ev._upload, ev.run_staged, ev._results_from_host


def timed(name, fn):
    def w(*a, **k):
        t0 = time.perf_counter()
        r = fn(*a, **k)
        acc[name] += time.perf_counter() - t0
        return r
    return w


ev._stage_host = timed('stage', orig_stage)
ev._upload = timed('upload', orig_up)
ev.run_staged = timed('launch', orig_run)
ev._results_from_host = timed('harvest', orig_res)
n = 64
t0 = time.perf_counter()
for _ in ev.predict_videos(videos[i % 8] for i in range(n)):
    pass
torch.cuda.synchronize()
tot = time.perf_counter() - t0
print(f'per video: total {tot / n * 1e3:.3f} ms; ' + ', '.join(f'{k} {v / n * 1e3:.3f} ms' for k, v in acc.items()))

#!/usr/bin/env python
"""One eager NLQ step between cudaProfilerStart/Stop (for `ncu --profile-from-start off`): 3 warm-up steps, then
exactly one step = 1 video x 16 queries through text encoder -> grounder -> decode -> NMS."""
import os
import sys

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
videos = [synth.synth_video(opt, 2000, 16, seed=2022 + i, tag=f'v{i}', n_events=1) for i in range(2)]
ev = Evaluator(opt.clone(), dataset=videos, state_dict=sd, use_graphs=False)
sts = []
for v in videos:
    st = ev._stage_inputs(v)
    torch.cuda.synchronize()
    sts.append({k: (st[k].clone() if isinstance(st[k], torch.Tensor) else st[k]) for k in st})
for i in range(3):
    ev._device_pass(sts[i % 2])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
ev._device_pass(sts[1])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()

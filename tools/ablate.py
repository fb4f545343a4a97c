#!/usr/bin/env python
"""What each stage costs on the critical path of the CUDA-graph replay of one NLQ step: replays the step with one
C-ABI entry point (or the text encoder) stubbed out and reports the difference to the full step.  Buffers keep
their last contents, so the stubbed graphs still run on finite data; only timing is meaningful here."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, 'cvpr2025-decafnet_b200')):
    if _p not in sys.path:
        sys.path.insert(0, _p)
import torch
from decaf_b200 import _cabi as cabi, synth
from decaf_b200.worker_v2 import Evaluator, create_model

opt = synth.nlq_opt()
shapes = {k: tuple(v.shape) for k, v in create_model(opt.clone()).state_dict().items()}
sd = synth.fill_state_dict(shapes, 2022)
videos = [synth.synth_video(opt, 2000, 16, seed=2022 + i, tag=f'v{i}', n_events=1) for i in range(4)]


def measure(stub=None, fused_text=False):
    ev = Evaluator(opt.clone(), dataset=videos, state_dict=sd)
    eng = ev.model.engine()
    eng.fused_text = fused_text
    sts = []
    for v in videos:
        st = ev._stage_inputs(v)
        torch.cuda.synchronize()
        sts.append({k: (st[k].clone() if isinstance(st[k], torch.Tensor) else st[k]) for k in st})
    ev.use_graphs = False
    ev._device_pass(sts[0])                 # real pass: every buffer holds real data
    torch.cuda.synchronize()
    saved = {}
    if stub == 'text':
        cached = eng.encode_text_batch(sts[0]['d_tok'], sts[0]['d_len'])
        saved['enc'] = eng.encode_text_batch
        eng.encode_text_batch = lambda tok, lens: cached
    elif stub is not None:
        for name in stub.split(','):
            saved[name] = getattr(cabi, name)
            setattr(cabi, name, lambda *a, **k: None)
    ev.use_graphs = True
    for i in range(8):
        ev.run_staged(sts[i % 4])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 40
    for i in range(n):
        ev.run_staged(sts[i % 4])
    e1.record()
    torch.cuda.synchronize()
    for name, fn in saved.items():
        if name == 'enc':
            eng.encode_text_batch = fn
        else:
            setattr(cabi, name, fn)
    return e0.elapsed_time(e1) / n * 1e3


full = measure()
print(f'full step                {full:8.1f} us')
print(f'full step (fused text)   {measure(fused_text=True):8.1f} us')
for stub in ['text'] if '--quick' in sys.argv else ['text', 'preattn', 'local_attn', 'xattn', 'tcn_in,tcn_layer,tcn_out,refine_pool', 'head_out', 'decode', 'batched_nms',
             'layernorm', 'adaln', 'merge', 'saliency,select,build_masks', 'gemm']:
    t = measure(stub)
    print(f'without {stub:40s} {t:8.1f} us   (stage cost {full - t:7.1f} us)')

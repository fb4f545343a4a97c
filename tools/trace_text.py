"""Debug: per-stage clock64 timeline of cluster 0 of the text-encoder kernel (decaf_debug_text_trace)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'cvpr2025-decafnet_b200')):
    sys.path.insert(0, p)
import torch
from decaf_b200 import _cabi as cabi, synth
from decaf_b200.worker_v2 import Evaluator, create_model

opt = synth.nlq_opt()
shapes = {k: tuple(v.shape) for k, v in create_model(opt.clone()).state_dict().items()}
sd = synth.fill_state_dict(shapes, 2022)
ev = Evaluator(opt.clone(), dataset=[], state_dict=sd)
eng = ev.model.engine()
n, Lmax = 16, 24
tok = torch.randn(n, Lmax, 768, device='cuda')
lens = torch.randint(6, 25, (n,), device='cuda', dtype=torch.int32)
print('max active clusters:', cabi.debug_text_max_clusters())
for _ in range(3):
    eng._encode_text_fused(tok, lens)
torch.cuda.synchronize()
buf = torch.zeros(256, dtype=torch.int64, device='cuda')
cabi.debug_text_trace(buf.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); eng._encode_text_fused(tok, lens); e1.record()
torch.cuda.synchronize()
cabi.debug_text_trace(None)
print(f'event time {e0.elapsed_time(e1) * 1e3:.1f} us')
v = [int(x) for x in buf.cpu() if x > 0]
print('stage deltas (cycles):', [b - a for a, b in zip(v[:-1], v[1:])])
print('total cycles', v[-1] - v[0])

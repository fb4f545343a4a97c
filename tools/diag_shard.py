"""Debug: where (level, row) do the owned points of time shards differ from the unsharded run?
    python tools/diag_shard.py [seed] [n_query] [S] [halo_mode]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'cvpr2025-decafnet_b200')):
    sys.path.insert(0, p)
import torch
from decaf_b200 import synth
from decaf_b200.time_shard import TimeShardedEvaluator, plan_shards
from decaf_b200.worker_v2 import Evaluator, create_model

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 2023
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 2
S = int(sys.argv[3]) if len(sys.argv) > 3 else 4
mode = sys.argv[4] if len(sys.argv) > 4 else 'exchange'
opt = synth.nlq_opt()
shapes = {k: tuple(v.shape) for k, v in create_model(opt.clone()).state_dict().items()}
sd = synth.fill_state_dict(shapes, 2022)
data = synth.synth_video(opt, 70001, nq, seed=seed, tag='mad', n_events=2)
ev = Evaluator(opt.clone(), dataset=[data], state_dict=sd, act_dtype=torch.bfloat16, use_graphs=False)
eng = ev.model.engine()
ref = ev.predict_video(data)
T = ev.padded_len(70001)
p = eng.plan(nq, T)
L = eng.L
g_log = p.logits2.view(nq, p.Pp).clone()
g_off = p.offsets.view(nq, p.Pp, 2).clone()
g_l1 = p.logits1.view(nq, p.Pp).clone()
g_offs, g_lens = list(p.off), list(p.lens)
tse = TimeShardedEvaluator(ev, emulate=S, halo_mode=mode)
res = tse.predict_video(data)
print('final equal:', all(torch.equal(a['segments'], b['segments']) and torch.equal(a['scores'], b['scores']) for a, b in zip(res, ref)))
shards = plan_shards(T, S, L, tse.halo)
for i, s in enumerate(shards):
    (a, e), (w0, w1) = s['own'], s['win']
    eng.lane = 100 + i
    pp = eng.plan(nq, w1 - w0)
    for name, G, X in (('logits1', g_l1, pp.logits1.view(nq, pp.Pp)), ('logits2', g_log, pp.logits2.view(nq, pp.Pp)),
                       ('offsets', g_off, pp.offsets.view(nq, pp.Pp, 2))):
        for l in range(L):
            ga, ge = a >> l, e >> l
            la, le = (a - w0) >> l, (e - w0) >> l
            gg = G[:, g_offs[l] + ga:g_offs[l] + ge]
            xx = X[:, pp.off[l] + la:pp.off[l] + le]
            d = (gg != xx)
            if d.dim() == 3:
                d = d.any(-1)
            if bool(d.any()):
                rows = d.any(0).nonzero().flatten()
                print(f'shard {i} own {a}-{e} {name} level {l}: {int(d.sum())} differing points, rows (own coords) {rows[:8].tolist()} .. {rows[-4:].tolist()} of {ge - ga}; '
                      f'max abs {float((gg.float() - xx.float()).abs().max()):.3e}')

# ---- second pass: per-level FPN features (captured) of every shard against the unsharded run
class LaneDict(dict):
    def __setitem__(self, k, v):
        super().__setitem__((eng.lane, k), v)


eng.lane = 0
eng.capture = {}
ev.predict_video(data)
gcap = eng.capture
eng.capture = LaneDict()
tse.predict_video(data)
scap = eng.capture
eng.capture = None
for i, s in enumerate(shards):
    (a, e), (w0, w1) = s['own'], s['win']
    for l in range(L):
        G = gcap[f'fpn{l}'][:, a >> l:e >> l]
        X = scap[(100 + i, f'fpn{l}')][:, (a - w0) >> l:(e - w0) >> l]
        d = (G != X).any(-1)
        if bool(d.any()):
            rows = d.any(0).nonzero().flatten()
            print(f'shard {i} fpn{l}: rows (own coords) {rows[:6].tolist()} .. {rows[-3:].tolist()} of {(e - a) >> l}; window rows +{(a - w0) >> l}; max abs {float((G - X).abs().max()):.3e}')

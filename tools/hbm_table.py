#!/usr/bin/env python
"""Achieved HBM bandwidth of the byte-bound kernels of the path at the MAD size (BASELINE.json configs[2]: t = 70,001 clips ->
T = 71,424 steps, P = 142,290 points per query, 64 queries, NLQ network), one unsharded pass on one GPU.

Per call: device time (CUDA events around the C-ABI call, warm, median over the passes) and ALGORITHMIC bytes (every input
read once, every output written once — the formulas are next to each kernel below), against the measured HBM copy peak
(MEASURED_PEAKS.json, else the 6538 GB/s of profiles/README.md).

    python tools/hbm_table.py [--queries 64] [--clips 70001] [--passes 3]  > profiles/r02_hbm_bound_kernels.txt
"""
import argparse
import collections
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, 'cvpr2025-decafnet_b200')):
    sys.path.insert(0, _p)
import torch
from decaf_b200 import _cabi as cabi, synth
from decaf_b200.worker_v2 import Evaluator, create_model


def nbytes(*ts):
    return sum(t.numel() * t.element_size() for t in ts if isinstance(t, torch.Tensor))


def algorithmic_bytes(name, args, kw):
    """-> (bytes, formula)"""
    if name == 'saliency':                    # shallow (Cs, T) fp32 + class embeddings (n, Cs) -> correl (n, T) fp32
        sh, cls, cor, Cs, T, n = args[:6]
        return (Cs * T + n * Cs + n * T) * 4, 'Cs*T*4 + n*Cs*4 + n*T*4'
    if name == 'select':                      # correl (n, T) fp32 + video mask (T) -> selection + merged mask (n, T) u8 each + pooled blocks
        cor, vm, sel, om, pooled, blocks, T, n = args[:8]
        return n * T * 4 + T + 2 * n * T + n * blocks * 4, 'n*T*4 + T + 2*n*T + n*blocks*4'
    if name == 'map_combine':                 # E, S (T, C) fp32 + correl/sel/mask (n, T) -> X (n, T, C) fp32
        E, S, bias, cor, wc, sel, mask, X, T, C, n = args[:11]
        return 2 * T * C * 4 + n * T * (4 + 1 + 1) + n * T * C * X.element_size(), '2*T*C*4 + n*T*6 + n*T*C*4'
    if name == 'build_masks':
        m0, stride, hmask, lv, n = args[:5]
        return nbytes(m0) + nbytes(hmask), 'n*T + n*P'
    if name == 'adaln':                       # q (rows, C) fp32 in/out + scale|shift (rows, 2C) + mask -> q, act copy
        q, rows, C, ss, mask = args[:5]
        oa = args[8]
        return rows * C * (4 + 4) + rows * 2 * C * ss.element_size() + rows + rows * C * oa.element_size(), 'rows*C*8 + rows*2C*2 + rows + rows*C*2'
    if name == 'layernorm':
        x, C, n_seq, rps = args[:4]
        rows = n_seq * rps
        b = rows * C * x.element_size()
        if kw.get('out_act') is not None:
            b += rows * C * kw['out_act'].element_size()
        if kw.get('out_f32') is not None:
            b += rows * C * 4
        return b, 'rows*C*(4 in + 2 and/or 4 out)'
    if name == 'decode':                      # logits (n, P) + offsets (n, P, 2) fp32 + masks -> top-k candidates
        logits, offsets, hmask, lv, n = args[:5]
        topk = args[7]
        P = logits.numel() // n
        return n * P * (4 + 8 + 1) + n * topk * 16, 'n*P*13 + n*topk*16'
    if name == 'batched_nms':                 # candidates (n, topk): segments + scores -> <= max_num_segs results
        segs, scores, cnt, n, stride = args[:5]
        return n * stride * 12 * 2, 'n*topk*12 in, scores/segments rewritten once (soft-NMS)'
    if name == 'refine_pyramid':
        return None, ''
    return None, ''


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--queries', type=int, default=64)
    ap.add_argument('--clips', type=int, default=70001)
    ap.add_argument('--passes', type=int, default=3)
    a = ap.parse_args()
    peak, src = 6538.0, 'profiles/README.md (measured copy bandwidth of this pool)'
    mp = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(mp):
        try:
            j = json.load(open(mp))
            for k in ('hbm_gbs_sustained', 'hbm_gbs', 'hbm_copy_gbs'):
                if k in j:
                    peak, src = float(j[k]), f'MEASURED_PEAKS.json:{k}'
                    break
        except Exception:
            pass
    opt = synth.nlq_opt()
    shapes = {k: tuple(v.shape) for k, v in create_model(opt.clone()).state_dict().items()}
    sd = synth.fill_state_dict(shapes, 2022)
    data = synth.synth_video(opt, a.clips, a.queries, seed=2022, tag='mad', n_events=2)
    ev = Evaluator(opt.clone(), dataset=[data], state_dict=sd, act_dtype=torch.bfloat16, use_graphs=False, n_lanes=1)
    ev.predict_video(data)
    torch.cuda.synchronize()
    names = ['saliency', 'select', 'map_combine', 'build_masks', 'adaln', 'layernorm', 'decode', 'batched_nms']
    records = []
    orig = {}

    def wrap(name, fn):
        def w(*args, **kw):
            b, f = algorithmic_bytes(name, args, kw)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*args, **kw)
            e1.record()
            tag = name
            if name == 'layernorm':
                tag = f'layernorm rows={args[2] * args[3]}'
            records.append((tag, b, f, e0, e1))
            return r
        return w
    for n in names:
        orig[n] = getattr(cabi, n)
        setattr(cabi, n, wrap(n, orig[n]))
    agg = collections.OrderedDict()
    for _ in range(a.passes):
        records.clear()
        ev.predict_video(data)
        torch.cuda.synchronize()
        seen = collections.Counter()
        for tag, b, f, e0, e1 in records:
            seen[tag] += 1
            agg.setdefault((tag, seen[tag]), []).append((e0.elapsed_time(e1) * 1e3, b, f))
    T = ev.padded_len(a.clips)
    print(f'# byte-bound kernels at the MAD size: t = {a.clips} (T = {T}), {a.queries} queries, NLQ network, one GPU, unsharded, bf16 configuration')
    print(f'# device time = CUDA events around the C-ABI call (includes ~2-4 us of launch gap), median of {a.passes} warm passes')
    print(f'# peak = {peak:.0f} GB/s ({src})')
    print(f'{"kernel":<28} {"us":>9} {"alg. MB":>10} {"GB/s":>8} {"of peak":>8}  bytes')
    rows = collections.OrderedDict()
    for (tag, k), v in agg.items():
        us = statistics.median(x[0] for x in v)
        r = rows.setdefault(tag, [0.0, 0, v[0][2], 0])
        r[0] += us
        r[1] += v[0][1] or 0
        r[3] += 1
    for tag, (us, b, f, cnt) in sorted(rows.items(), key=lambda kv: -kv[1][0]):
        gbs = b / us / 1e3 if b else float('nan')
        print(f'{tag + (" x%d" % cnt if cnt > 1 else ""):<28} {us:9.1f} {b / 1e6:10.1f} {gbs:8.0f} {gbs / peak:8.2f}  {f}')


if __name__ == '__main__':
    main()

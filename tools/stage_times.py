#!/usr/bin/env python
"""Per-launch device times of one NLQ step under WARM caches (CUDA events around every C-ABI call of an eager
pass, CPU running ahead of the GPU behind a spin kernel).  ncu's launch list (profiles/*.csv) is cold-cache and
serialised; this is the complementary view used to rank optimisation targets.

    python tools/stage_times.py [--iters 5] [--dtype bf16] [--out gpurun_out/stage_times.txt]
"""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, 'cvpr2025-decafnet_b200')):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=5)
    ap.add_argument('--dtype', default='bf16')
    ap.add_argument('--out', default=None)
    ap.add_argument('--nq', type=int, default=16)
    ap.add_argument('--vid-len', type=int, default=2000)
    a = ap.parse_args()
    from decaf_b200 import _cabi as cabi, synth
    from decaf_b200.worker_v2 import Evaluator, create_model
    opt = synth.nlq_opt()
    shapes = {k: tuple(v.shape) for k, v in create_model(opt.clone()).state_dict().items()}
    sd = synth.fill_state_dict(shapes, 2022)
    videos = [synth.synth_video(opt, a.vid_len, a.nq, seed=2022 + i, tag=f'v{i}', n_events=1) for i in range(2)]
    act = torch.bfloat16 if a.dtype == 'bf16' else torch.float32
    ev = Evaluator(opt.clone(), dataset=videos, state_dict=sd, act_dtype=act, use_graphs=False)
    sts = []
    for v in videos:
        st = ev._stage_inputs(v)
        torch.cuda.synchronize()
        r = {k: (st[k].clone() if isinstance(st[k], torch.Tensor) else st[k]) for k in st}
        sts.append(r)
    for st in sts:
        ev._device_pass(st)
    torch.cuda.synchronize()

    records = []            # (tag, e0, e1)
    names = ['gemm', 'layernorm', 'preattn', 'adaln', 'local_attn', 'xattn', 'saliency', 'select', 'merge', 'build_masks',
             'head_out', 'tcn_in', 'tcn_layer', 'tcn_out', 'refine_pool', 'text_prep', 'decode', 'batched_nms',
             'text_encoder', 'tcn_fused', 'refine_pyramid', 'map_combine', 'ffn', 'upload_2d', 'merge_candidates', 'decode_window',
             'xattn_packed', 'xattn_pack_kv']
    orig = {}

    def wrap(name, fn):
        def w(*args, **kw):
            tag = name
            if name == 'gemm':
                A, W, N, K, n_seq, rps = args[:6]
                tag = (f"gemm M={n_seq * rps} N={N} K={K} t={kw.get('taps', 1)} g={kw.get('n_group', 1)} "
                       f"{'f32' if A.dtype == torch.float32 else 'bf16'}"
                       f"{' ln' if kw.get('ln') else ''}{' act%d' % kw['act'] if kw.get('act') else ''}"
                       f"{' res' if kw.get('resid') is not None else ''}")
            elif name == 'ffn':
                tag = f'ffn (fc + GELU + proj fused) M={args[6] * args[7]} C={args[5]}'
            elif name in ('preattn', 'local_attn'):
                tag = f'{name} rows={args[1] * args[2] if name == "preattn" else args[4] * args[5]}'
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*args, **kw)
            e1.record()
            records.append((tag, e0, e1))
            return r
        return w

    for n in names:
        if hasattr(cabi, n):
            orig[n] = getattr(cabi, n)
            setattr(cabi, n, wrap(n, orig[n]))
    agg = collections.OrderedDict()
    total = 0.0
    for it in range(a.iters):
        records.clear()
        torch.cuda.synchronize()
        torch.cuda._sleep(int(40e6))          # ~20 ms: the CPU queues the whole step behind it
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        ev._device_pass(sts[it % 2])
        s1.record()
        torch.cuda.synchronize()
        total += s0.elapsed_time(s1)
        for tag, e0, e1 in records:
            d = agg.setdefault(tag, [0, 0.0])
            d[0] += 1
            d[1] += e0.elapsed_time(e1)
    lines = [f'# warm-cache per-call device time (CUDA events, includes ~2-4 us event/launch gap), avg over {a.iters} eager steps',
             f'# step total {total / a.iters * 1e3:.1f} us, launches/step {sum(d[0] for d in agg.values()) // a.iters}']
    rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
    for tag, (n, ms) in rows:
        lines.append(f'{ms / a.iters * 1e3:9.1f} us  n={n // a.iters:3d}  avg {ms / n * 1e3:7.1f} us  {tag}')
    txt = '\n'.join(lines)
    print(txt)
    if a.out:
        os.makedirs(os.path.dirname(a.out), exist_ok=True)
        with open(a.out, 'w') as f:
            f.write(txt + '\n')


if __name__ == '__main__':
    main()

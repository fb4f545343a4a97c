#!/usr/bin/env python
"""What each kernel family costs on the THROUGHPUT path (4 videos in flight): the step is re-captured with one C-ABI entry
point stubbed out (outputs stale, timing only) and the device-resident step time compared with the full step.  With lanes
a kernel's latency is hidden; what remains is the SM time it takes away from the other lanes.

    python tools/ablate_lanes.py [--lanes 4]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, 'cvpr2025-decafnet_b200')):
    if _p not in sys.path:
        sys.path.insert(0, _p)
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--lanes', type=int, default=8)
    a = ap.parse_args()
    from decaf_b200 import _cabi as cabi, synth
    from decaf_b200.worker_v2 import Evaluator, create_model
    opt = synth.nlq_opt()
    shapes = {k: tuple(v.shape) for k, v in create_model(opt.clone()).state_dict().items()}
    sd = synth.fill_state_dict(shapes, 2022)
    videos = [synth.synth_video(opt, 2000, 16, seed=2022 + i, tag=f'v{i}', n_events=1) for i in range(8)]

    def measure(stub):
        orig = {}
        for name in stub:
            orig[name] = getattr(cabi, name)
            setattr(cabi, name, lambda *x, **k: None)
        try:
            ev = Evaluator(opt.clone(), dataset=videos, state_dict=sd, n_lanes=a.lanes)
            res = []
            for i, v in enumerate(videos):
                st = ev._stage_inputs(v, i % a.lanes)
                torch.cuda.synchronize()
                r = {k: st[k].clone() for k in ('d_vid', 'd_sh', 'd_mask', 'd_tok', 'd_len', 'd_cls', 'd_meta')}
                r['key'], r['lane'] = st['key'], st['lane']
                res.append(r)
            for r in res:
                ev.launch_staged(r)
            ev.join_lanes()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(48):
                ev.launch_staged(res[i % len(res)])
            ev.join_lanes()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / 48 * 1e3
        finally:
            for name, fn in orig.items():
                setattr(cabi, name, fn)

    full = measure([])
    print(f'# exposed cost per kernel family with {a.lanes} videos in flight (full step {full:.0f} us)')
    for label, stub in (('pre-attention (preattn)', ['preattn']), ('local attention', ['local_attn']), ('cross attention (video + text)', ['xattn']),
                        ('LayerNorm launches', ['layernorm']), ('AdaLN', ['adaln']), ('fused TCN + pyramid', ['tcn_fused', 'refine_pyramid']),
                        ('head output convs', ['head_out']), ('decode + NMS', ['decode', 'batched_nms']),
                        ('saliency + select + map', ['saliency', 'select', 'merge', 'map_combine', 'build_masks']),
                        ('all GEMM launches', ['gemm'])):
        t = measure(stub)
        print(f'{full - t:8.0f} us  {label}')


if __name__ == '__main__':
    main()

#!/usr/bin/env python
"""Kernel timeline of the THROUGHPUT path (CUDA graphs, several videos in flight) from CUPTI activity records
(torch.profiler; nsys is not in the image): which kernels overlap, how much wall time each family has the GPU to itself,
where the lanes leave it idle.

    python tools/timeline.py [--lanes 4] [--videos 16] [--out gpurun_out/timeline.json]

Prints, for the profiled window: wall time per video, the union of busy time, and per kernel family
  sum   = sum of its kernels' durations (per video)
  solo  = wall time during which ONLY kernels of this family were running (per video): what the family costs when
          nothing hides it
  n     = launches per video.
"""
import argparse
import collections
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, 'cvpr2025-decafnet_b200')):
    if _p not in sys.path:
        sys.path.insert(0, _p)
import torch


def family(name):
    name = re.sub(r'^void\s+', '', name).replace('decaf::', '')
    base = re.sub(r'[<(].*$', '', name)
    if base == 'gemm_tc_kernel':
        m = re.match(r'gemm_tc_kernel<\(int\)(-?\d+), \(bool\)(\d), \(int\)(\d+)>', name) or re.match(r'gemm_tc_kernel<(-?\d+), (\d), (\d+)>', name)
        return f'gemm_tc<{m.group(1)},{m.group(2)},{m.group(3)}>' if m else 'gemm_tc'
    return base


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--lanes', type=int, default=8)
    ap.add_argument('--videos', type=int, default=16)
    ap.add_argument('--out', default=None)
    a = ap.parse_args()
    from decaf_b200 import synth
    from decaf_b200.worker_v2 import Evaluator, create_model
    opt = synth.nlq_opt()
    shapes = {k: tuple(v.shape) for k, v in create_model(opt.clone()).state_dict().items()}
    sd = synth.fill_state_dict(shapes, 2022)
    videos = [synth.synth_video(opt, 2000, 16, seed=2022 + i, tag=f'v{i}', n_events=1) for i in range(8)]
    ev = Evaluator(opt.clone(), dataset=videos, state_dict=sd, n_lanes=a.lanes)
    res = []
    for i, v in enumerate(videos):
        st = ev._stage_inputs(v, i % a.lanes)
        torch.cuda.synchronize()
        r = {k: st[k].clone() for k in ('d_vid', 'd_sh', 'd_mask', 'd_tok', 'd_len', 'd_cls', 'd_meta')}
        r['key'], r['lane'] = st['key'], st['lane']
        res.append(r)
    for _ in range(2):
        for r in res:
            ev.launch_staged(r)
    ev.join_lanes()
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(a.videos):
            ev.launch_staged(res[i % len(res)])
        ev.join_lanes()
        torch.cuda.synchronize()
    trace = (a.out or '/tmp/decaf_timeline') + '.chrome.json'
    os.makedirs(os.path.dirname(trace) or '.', exist_ok=True)
    prof.export_chrome_trace(trace)
    evs = []
    for e in json.load(open(trace))['traceEvents']:
        if e.get('cat') == 'kernel' and e.get('dur', 0) > 0:
            g = e.get('args', {}).get('grid', [1, 1, 1])
            evs.append((float(e['ts']), float(e['ts']) + float(e['dur']), family(e['name']), int(g[0]) * int(g[1]) * int(g[2])))
    os.remove(trace)
    evs.sort()
    if not evs:
        raise SystemExit('no kernel records (CUPTI unavailable?)')
    # drop the ramp: keep the middle half of the videos by time
    t0, t1 = evs[0][0], max(e[1] for e in evs)
    lo, hi = t0 + (t1 - t0) * 0.25, t0 + (t1 - t0) * 0.75
    win = [(max(s, lo), min(e, hi), f, g) for s, e, f, g in evs if e > lo and s < hi]
    n_vid = a.videos * 0.5
    # sweep line
    pts = []
    for s, e, f, g in win:
        pts.append((s, 1, f, g))
        pts.append((e, -1, f, g))
    pts.sort(key=lambda p: (p[0], p[1]))
    active = collections.Counter()
    solo = collections.Counter()
    busy = 0.0
    conc_hist = collections.Counter()
    ctas = 0                                   # CTAs of the running kernels (an SM holds one GEMM CTA, several small ones)
    fill_hist = collections.Counter()          # wall time by (sum of grid sizes / 148), capped at 1: how full the GPU can be
    low_fill = collections.Counter()           # families running while that ratio is below 0.5
    last = lo
    for t, d, f, g in pts:
        dt = t - last
        if dt > 0:
            fams = [k for k, v in active.items() if v > 0]
            n_act = sum(active.values())
            conc_hist[min(n_act, 6)] += dt
            fill = min(ctas / 148.0, 1.0)
            fill_hist[min(int(fill * 4), 3)] += dt
            if fams:
                busy += dt
                if len(fams) == 1:
                    solo[fams[0]] += dt
                if fill < 0.5:
                    low_fill['+'.join(sorted(fams))] += dt
        active[f] += d
        ctas += d * g
        last = t
    tot = collections.Counter()
    cnt = collections.Counter()
    for s, e, f, g in win:
        tot[f] += e - s
        cnt[f] += 1
    wall = hi - lo
    print(f'# {a.lanes} videos in flight, window = middle half of {a.videos} graph replays: wall {wall / n_vid:.1f} us per video, '
          f'GPU busy (>= 1 kernel running) {busy / n_vid:.1f} us, idle {(wall - busy) / n_vid:.1f} us')
    print('# concurrency histogram (us per video with k kernels running): ' +
          ', '.join(f'{k}{"+" if k == 6 else ""}: {v / n_vid:.0f}' for k, v in sorted(conc_hist.items())))
    print('# wall time per video by CTAs in flight / 148 SMs: ' + ', '.join(f'{25 * k}-{25 * k + 25}%: {v / n_vid:.0f} us' for k, v in sorted(fill_hist.items())))
    print('# what runs while fewer than 74 CTAs are in flight (us per video):')
    for k, v in sorted(low_fill.items(), key=lambda kv: -kv[1])[:12]:
        print(f'#   {v / n_vid:7.1f}  {k}')
    print(f'{"sum us":>9} {"solo us":>9} {"n":>6}  family')
    for f, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f'{v / n_vid:9.1f} {solo[f] / n_vid:9.1f} {cnt[f] / n_vid:6.1f}  {f}')
    if a.out:
        os.makedirs(os.path.dirname(a.out) or '.', exist_ok=True)
        json.dump([(s - t0, e - t0, f, g) for s, e, f, g in evs], open(a.out, 'w'))


if __name__ == '__main__':
    main()

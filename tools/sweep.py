#!/usr/bin/env python
"""BASELINE.json configs 4 and 5 on one B200: JSON lines, one per measurement.

    python tools/sweep.py charades  [--videos 512 --queries 16 --lanes 8]    # config 4: short videos, large query batch
    python tools/sweep.py lengths   [--queries 16]                           # config 5: sratio {0.1,0.3,0.5} x t {1k..100k}
    python tools/sweep.py nms       [--batch 16]                             # config 5: batched 1D-NMS microbenchmark

`charades` and `lengths` time Evaluator.predict_videos (host inputs, H2D and D2H inside the timed region, CUDA-graph
replay, `--lanes` videos in flight) AND the device-resident replay; `nms` times decaf_batched_nms alone with CUDA events
and — only with --cpu-reference, a baseline leg of the same kind as bench.py's cpu_baseline, never on the measured path —
up to --cpu-max candidates the reference's own CPU extension (oracle/_ref) or its C twin on the host.
Multi-GPU (config 4 at 8 GPUs): launch under torchrun — videos are dealt round-robin to ranks, no collective on the
data path (the same sharding as bench.py).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, 'cvpr2025-decafnet_b200')):
    if _p not in sys.path:
        sys.path.insert(0, _p)
import numpy as np
import torch


def _env():
    return int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))


def _model(opt, seed=2022):
    from decaf_b200 import synth
    from decaf_b200.worker_v2 import create_model
    shapes = {k: tuple(v.shape) for k, v in create_model(opt.clone()).state_dict().items()}
    return synth.fill_state_dict(shapes, seed)


def _time_videos(ev, videos, steps, warmup):
    """(e2e pairs/s through predict_videos, device-resident pairs/s through launch_staged)."""
    n_pairs = sum(len(v['text']) for v in videos)
    for _ in range(warmup):
        for _ in ev.predict_videos(videos):
            pass
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        for _ in ev.predict_videos(videos):
            pass
    torch.cuda.synchronize()
    e2e = n_pairs * steps / (time.perf_counter() - t0)
    # device-resident: clone the staged inputs of (up to 16) videos, replay their graphs lane by lane
    res = []
    for i, v in enumerate(videos[:16]):
        st = ev._stage_inputs(v, i % ev.n_lanes)
        torch.cuda.synchronize()
        r = {k: st[k].clone() for k in ('d_vid', 'd_sh', 'd_mask', 'd_tok', 'd_len', 'd_cls', 'd_meta')}
        r['key'], r['lane'] = st['key'], st['lane']
        res.append(r)
    for r in res:
        ev.launch_staged(r)
    ev.join_lanes()
    torch.cuda.synchronize()
    n = max(len(videos) * steps, len(res))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        ev.launch_staged(res[i % len(res)])
    ev.join_lanes()
    e1.record()
    torch.cuda.synchronize()
    pairs_dev = sum(len(videos[i % len(res) % len(videos)]['text']) for i in range(n))
    return e2e, pairs_dev / (e0.elapsed_time(e1) * 1e-3)


def run_charades(a):
    from decaf_b200 import synth
    from decaf_b200.worker_v2 import Evaluator
    rank, world, local = _env()
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    opt = synth.charades_opt()
    sd = _model(opt)
    mine = list(range(rank, a.videos, world))                      # videos (with all their queries) dealt to ranks
    pool = [synth.synth_video(opt, 200, a.queries, seed=3000 + i, tag=f'c{i}', text_len_range=(4, 12), n_events=1)
            for i in mine[:32]]
    videos = [pool[i % len(pool)] for i in range(len(mine))]
    ev = Evaluator(opt.clone(), dataset=videos, state_dict=sd, n_lanes=a.lanes)
    e2e, dev = _time_videos(ev, videos, a.steps, 1)
    t = torch.tensor([e2e, dev], dtype=torch.float64, device='cuda')
    if dist is not None:
        dist.all_reduce(t)                                          # sum of per-rank rates (weak scaling, no collective on the path)
    if rank == 0:
        print(json.dumps({'config': 4, 'workload': f'Charades/TACoS shape: t=200 (T=256), embd 128, 6 levels, win 5, '
                                                   f'{a.videos} videos x {a.queries} queries, query-sharded over {world} GPU(s)',
                          'n_gpus': world, 'lanes': a.lanes, 'e2e_pairs_per_s': float(t[0]), 'device_pairs_per_s': float(t[1])}))
    if dist is not None:
        dist.destroy_process_group()


def run_c512(a):
    """SURVEY.md section 8(d) config 3 names the C = 512 variant of the network: the NLQ shape (t = 2000 -> T = 2304, 16 queries)
    at embd 512 (FFN 2048, second heads 544 channels wide: conv -> row-wise LayerNorm instead of the fused epilogue)."""
    from decaf_b200 import synth
    from decaf_b200.worker_v2 import Evaluator
    torch.cuda.set_device(0)
    opt = synth.nlq_opt(embd_dim=512, n_heads=a.heads)      # 8 heads: head dim 64 (the tensor-core attention kernels cover 32 / 64)
    sd = _model(opt)
    videos = [synth.synth_video(opt, 2000, a.queries, seed=5120 + i, tag=f'w{i}', n_events=1) for i in range(8)]
    ev = Evaluator(opt.clone(), dataset=videos, state_dict=sd, n_lanes=a.lanes)
    e2e, dev = _time_videos(ev, videos * 4, a.steps, 1)
    print(json.dumps({'config': '3 (C = 512 variant)', 'workload': f'NLQ shape t=2000 (T=2304), {a.queries} queries, embd 512, {a.heads} heads, 8 levels, win 19',
                      'n_gpus': 1, 'lanes': a.lanes, 'e2e_pairs_per_s': e2e, 'device_pairs_per_s': dev,
                      'peak_mem_gib': torch.cuda.max_memory_allocated() / 2 ** 30}))


def run_lengths(a):
    from decaf_b200 import synth
    from decaf_b200.worker_v2 import Evaluator
    torch.cuda.set_device(0)
    for t in a.lengths:
        for sratio in (0.1, 0.3, 0.5):
            opt = synth.nlq_opt(sratio=sratio)
            sd = _model(opt)
            videos = [synth.synth_video(opt, t, a.queries, seed=4000 + i, tag=f'l{t}_{i}', n_events=1) for i in range(2)]
            lanes = 3 if t <= 10000 else 1
            ev = Evaluator(opt.clone(), dataset=videos, state_dict=sd, n_lanes=lanes)
            e2e, dev = _time_videos(ev, videos, 2 if t > 10000 else 6, 1)
            T = ev.padded_len(t)
            sel = float(ev.model.engine().plan(a.queries, T).sel.float().mean())
            print(json.dumps({'config': 5, 'sweep': 'length x sratio', 't': t, 'T': T, 'sratio': sratio, 'queries': a.queries,
                              'lanes': lanes, 'selected_fraction_of_T': sel, 'e2e_pairs_per_s': e2e, 'device_pairs_per_s': dev,
                              'device_ms_per_video': a.queries / dev * 1e3,
                              'peak_mem_gb': torch.cuda.max_memory_allocated() / 2 ** 30}), flush=True)
            del ev
            torch.cuda.empty_cache()


def run_nms(a):
    from decaf_b200 import _cabi as cabi
    torch.cuda.set_device(0)
    soft_ref = hard_ref = kind = None
    if a.cpu_reference:                                  # CPU baseline leg (checker / baseline only)
        from oracle import nms_oracle
        soft_ref, hard_ref = nms_oracle.reference_fns()
        kind = 'oracle/_ref nms_1d_cpu_vg (reference extension)'
        if soft_ref is None:
            soft_ref, hard_ref, kind = nms_oracle.softnms, nms_oracle.nms, 'oracle/nms_oracle.c (C twin)'
    g = torch.Generator().manual_seed(2022)
    for n in a.sizes:
        B = a.batch
        c = torch.rand(B, n, generator=g) * 2304.0
        ln = 1.0 + torch.rand(B, n, generator=g) * 199.0
        segs = torch.stack((c - 0.5 * ln, c + 0.5 * ln), -1).contiguous()
        scores = torch.rand(B, n, generator=g)
        d_segs, d_scores = segs.cuda(), scores.cuda()
        cnt = torch.full((B, ), n, dtype=torch.int32, device='cuda')
        ws = torch.empty(int(cabi.nms_workspace_bytes(B, n)), dtype=torch.uint8, device='cuda')
        for mode in ('soft_nms', 'nms'):
            prm = cabi.NmsParams()
            prm.mode = 2 if mode == 'soft_nms' else 1
            prm.iou_thresh, prm.sigma, prm.min_score, prm.max_num_segs, prm.voting_thresh = 0.1, 0.9, 1e-3, 5, 0.95
            out_s = torch.zeros(B, 5, 2, device='cuda'); out_c = torch.zeros(B, 5, device='cuda')
            out_n = torch.zeros(B, dtype=torch.int32, device='cuda')
            for _ in range(2):
                cabi.batched_nms(d_segs, d_scores, cnt, B, n, prm, out_s, out_c, out_n, ws)
            torch.cuda.synchronize()
            reps = 5 if n >= 100000 else 20
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                cabi.batched_nms(d_segs, d_scores, cnt, B, n, prm, out_s, out_c, out_n, ws)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            line = {'config': 5, 'sweep': 'nms', 'n': n, 'batch': B, 'mode': mode, 'gpu_ms_per_batch': ms,
                    'gpu_queries_per_s': B / (ms * 1e-3), 'algorithmic_gbs': B * n * 12 / (ms * 1e-3) / 1e9}
            if a.cpu_reference and n <= a.cpu_max:
                from oracle import grounder_oracle as go
                t0 = time.perf_counter()
                ref = go.batched_nms(segs[0], scores[0], 0.1, 1e-3, 5, mode, 0.9, 0.95, softnms_fn=soft_ref, nms_fn=hard_ref)
                line['cpu_ms_per_query'] = (time.perf_counter() - t0) * 1e3
                line['cpu_kind'] = kind
                k = int(out_n[0])
                line['keep_set_matches_cpu'] = bool(k == ref[0].shape[0] and
                                                    np.allclose(out_s[0, :k].cpu().numpy(), ref[0].numpy(), rtol=1e-5, atol=1e-4))
            print(json.dumps(line), flush=True)


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('what', choices=['charades', 'lengths', 'nms', 'c512'])
    ap.add_argument('--videos', type=int, default=512)
    ap.add_argument('--queries', type=int, default=16)
    ap.add_argument('--lanes', type=int, default=8)
    ap.add_argument('--heads', type=int, default=8, help='c512: attention heads')
    ap.add_argument('--steps', type=int, default=2)
    ap.add_argument('--batch', type=int, default=16)
    ap.add_argument('--lengths', type=int, nargs='*', default=[1000, 2300, 10000, 30000, 70000, 100000])
    ap.add_argument('--sizes', type=int, nargs='*', default=[1000, 10000, 100000, 1000000])
    ap.add_argument('--cpu-max', type=int, default=10000)
    ap.add_argument('--cpu-reference', action='store_true', help='also time the reference CPU NMS (oracle/_ref) on the host')
    a = ap.parse_args()
    {'charades': run_charades, 'lengths': run_lengths, 'nms': run_nms, 'c512': run_c512}[a.what](a)

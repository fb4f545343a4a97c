#!/usr/bin/env python
"""Time-sharded grounding of one long video over the ranks of a torchrun launch (NCCL), checked against the
unsharded path on rank 0 when --check is given (the unsharded run must fit one GPU).

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/run_time_shard.py --clips 70001 --queries 64
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
import torch
import torch.distributed as dist


def run_query_sharded(a, ev, data, rank, world, T):
    """The alternative partition of a long video when there are at least as many queries as GPUs: rank r grounds queries
    r::world on the WHOLE timeline (pair sharding) — no recompute halo, no exchange; the results
    (<= max_num_segs segments per query) are gathered at the end.  Bit-identical to the one-GPU run by construction (rows of
    different queries never interact)."""
    mine = list(range(rank, a.queries, world))
    part = dict(data)
    part['text'] = tuple(data['text'][i] for i in mine)
    part['text_cls'] = data['text_cls'][mine].contiguous()
    part['segment'] = data['segment'][mine]
    for _ in range(a.warmup):
        res = ev.predict_video(part)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        res = ev.predict_video(part)
        if world > 1:                                    # gather the final segments of every query on every rank
            pack = torch.zeros(len(mine), 5, 3, device='cuda')
            for j, r in enumerate(res):
                k = r['scores'].numel()
                pack[j, :k, :2] = r['segments'].cuda()
                pack[j, :k, 2] = r['scores'].cuda()
            allp = [torch.zeros_like(pack) for _ in range(world)] if a.queries % world == 0 else None
            if allp is not None:
                dist.all_gather(allp, pack)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device='cuda', dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    sec = float(dt.item()) / a.steps
    if rank == 0:
        print(json.dumps({'workload': f'MAD-shape video: t={a.clips} (T={T}), {a.queries} queries, QUERY-sharded over {world} GPU(s) '
                                      f'({len(mine)} queries per rank, whole timeline, no halo)',
                          'n_gpus': world, 'ms_per_video': sec * 1e3, 'pairs_per_s': a.queries / sec,
                          'peak_mem_gb': torch.cuda.max_memory_allocated() / 2 ** 30}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--clips', type=int, default=70001)
    ap.add_argument('--queries', type=int, default=64)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=1)
    ap.add_argument('--check', action='store_true')
    ap.add_argument('--dtype', default='bf16')
    ap.add_argument('--halo-mode', default='exchange', choices=['exchange', 'recompute'],
                    help="'exchange': window-sized halo refreshed from the neighbours after every encoder output (NCCL send/recv); "
                         "'recompute': one-shot halo covering the whole receptive field")
    ap.add_argument('--shard', default='time', choices=['time', 'queries'],
                    help="'queries': every rank grounds a slice of the queries on the whole timeline (no halo, no collective on the path)")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from decaf_b200 import synth
    from decaf_b200.time_shard import TimeShardedEvaluator, plan_shards
    from decaf_b200.worker_v2 import Evaluator, create_model
    opt = synth.nlq_opt()
    shapes = {k: tuple(v.shape) for k, v in create_model(opt.clone()).state_dict().items()}
    sd = synth.fill_state_dict(shapes, 2022)
    data = synth.synth_video(opt, a.clips, a.queries, seed=2022, tag='mad', n_events=2)
    act = torch.bfloat16 if a.dtype == 'bf16' else torch.float32
    ev = Evaluator(opt.clone(), dataset=[data], state_dict=sd, act_dtype=act, use_graphs=False)
    T = ev.padded_len(a.clips)
    if a.shard == 'queries':
        return run_query_sharded(a, ev, data, rank, world, T)
    tse = TimeShardedEvaluator(ev, rank=rank, world=world, halo_mode=a.halo_mode)
    for _ in range(a.warmup):
        res = tse.predict_video(data)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        res = tse.predict_video(data)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device='cuda', dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    sec = float(dt.item()) / a.steps
    out = {'workload': f'MAD-shape video: t={a.clips} (T={T}), {a.queries} queries, time-sharded over {world} GPU(s), '
                       f'halo {tse.halo} ({a.halo_mode}), {tse.exchange_bytes / 2 ** 20:.1f} MiB sent per rank and video in halo exchanges',
           'n_gpus': world, 'ms_per_video': sec * 1e3, 'pairs_per_s': a.queries / sec,
           'shards': [s['own'] for s in plan_shards(T, world, 8, tse.halo)],
           'peak_mem_gb': torch.cuda.max_memory_allocated() / 2 ** 30}
    if a.check and rank == 0:
        ref = ev.predict_video(data)
        worst = 0.0
        same = True
        for b in range(len(ref)):
            same &= ref[b]['segments'].shape == res[b]['segments'].shape
            if same and ref[b]['segments'].numel():
                worst = max(worst, float((ref[b]['segments'] - res[b]['segments']).abs().max()))
                worst = max(worst, float((ref[b]['scores'] - res[b]['scores']).abs().max()))
        out['check'] = {'same_shapes': bool(same), 'max_abs_diff_vs_unsharded': worst}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

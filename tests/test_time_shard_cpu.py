"""CPU: host-side logic of the time-sharded path (decaf_b200/time_shard.py): shard planning, the halo bound,
the all-gather assembly of per-step saliency rows over a world_size-2 gloo group, and the candidate merge rule
(global top-k by score, ties by global flat index) against a direct global sort."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from decaf_b200 import synth, time_shard as ts


def test_plan_shards_covers_timeline_and_aligns():
    for T, world, L, halo in [(71424, 8, 8, 3840), (71424, 2, 8, 3840), (3072, 4, 5, 320), (2304, 3, 8, 3840)]:
        shards = ts.plan_shards(T, world, L, halo)
        align = 2 ** (L - 1)
        assert shards[0]['own'][0] == 0 and shards[-1]['own'][1] == T
        for a, b in zip(shards[:-1], shards[1:]):
            assert a['own'][1] == b['own'][0]
        for s in shards:
            (o0, o1), (w0, w1) = s['own'], s['win']
            assert o0 % align == 0 and o1 % align == 0 and w0 % align == 0 and w1 % align == 0
            assert w0 == max(0, o0 - halo) and w1 == min(T, o1 + halo)
            assert o1 > o0
        sizes = [s['own'][1] - s['own'][0] for s in shards]
        assert max(sizes) - min(sizes) <= align


def test_receptive_halo_bounds_the_measured_field():
    # 3.3k (left) / 3.6k (right) level-0 steps measured on the reference for L = 8, window 19
    h = ts.receptive_halo(synth.nlq_opt())
    assert h % 128 == 0 and 3600 <= h <= 4096
    assert ts.receptive_halo(synth.tiny_opt(n_levels=5, win=9)) % 16 == 0


def test_merge_rule_equals_global_sort():
    g = torch.Generator().manual_seed(3)
    n_src, n, topk = 3, 4, 16
    # quantised scores force ties; idx are unique global point indices
    scores = (torch.randint(0, 8, (n_src, n, topk), generator=g).float() / 8).sort(dim=-1, descending=True).values
    idx = torch.stack([torch.randperm(1000, generator=g)[:n_src * topk].reshape(n_src, topk) for _ in range(n)], 1).int()
    segs = torch.rand(n_src, n, topk, 2, generator=g)
    count = torch.randint(5, topk + 1, (n_src, n), generator=g).int()
    out = ts.merge_candidates_reference(segs, scores, idx, count, topk)
    for q in range(n):
        rows = [(-float(scores[r, q, j]), int(idx[r, q, j]), r, j) for r in range(n_src) for j in range(int(count[r, q]))]
        rows.sort()
        want = rows[:topk]
        got_idx = out[q][2].tolist()
        assert got_idx == [w[1] for w in want]
        assert torch.equal(out[q][0], torch.stack([segs[w[2], q, w[3]] for w in want]))


def _worker(rank, world, port, T, n):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        shards = ts.plan_shards(T, world, 5, 64)
        full = torch.arange(n * T, dtype=torch.float32).reshape(n, T)
        a, e = shards[rank]['own']
        comm = ts._DistComm()
        got = ts.assemble_rows(comm, shards, [rank], [full[:, a:e].contiguous()], T)
        assert torch.equal(got, full), rank
        # candidate lists: all-gather + reference merge is the same on every rank
        g = torch.Generator().manual_seed(100 + rank)
        topk = 8
        sc = torch.rand(n, topk, generator=g).sort(dim=-1, descending=True).values
        ix = (torch.arange(topk)[None] * world + rank).repeat(n, 1).int()
        sg = torch.rand(n, topk, 2, generator=g)
        cnt = torch.full((n,), topk, dtype=torch.int32)
        G = [torch.stack(comm.all_gather([t])) for t in (sg, sc, ix, cnt)]
        merged = ts.merge_candidates_reference(*G, topk)
        chk = torch.stack([m[1] for m in merged])
        ref = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(ref, chk)
        assert all(torch.equal(r, chk) for r in ref)
        assert bool((chk[:, :-1] >= chk[:, 1:]).all())
    finally:
        dist.destroy_process_group()


def test_allgather_assembly_gloo_world2():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, 48 * 16, 3), nprocs=2, join=True)

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


def test_exchange_halo_is_window_sized():
    """halo_mode='exchange': (1 + half window) rows of the coarsest level + 1 row of margin, i.e. 1408 level-0 steps for the
    NLQ / MAD network (8 levels, window 19) against 3840 for the one-shot recompute halo; it also covers the head / TCN /
    pyramid chain that follows the last exchange."""
    opt = synth.nlq_opt()
    h = ts.exchange_halo(opt)
    assert h == 11 * 128 and h < ts.receptive_halo(opt) // 2
    assert ts.exchange_halo(opt, margin_rows=0) == 10 * 128
    h5 = ts.exchange_halo(synth.tiny_opt(embd_dim=128, n_levels=5, win=9))
    assert h5 % 16 == 0 and h5 // 16 >= 1 + 4


@pytest.mark.parametrize('T,world,L,halo', [(71424, 8, 8, 1408), (71424, 4, 8, 1408), (71424, 2, 8, 2560), (3072, 3, 5, 160)])
def test_halo_rows_reproduce_the_global_rows(T, world, L, halo):
    """The row arithmetic of one halo exchange: every shard holds its window of a per-level global array with the outermost
    `rows` rows of each interior halo corrupted (what a layer computes from zero padding); after copying the neighbours'
    send ranges into the receive ranges every window equals the global slice again — at every level, for every shard."""
    shards = ts.plan_shards(T, world, L, halo)
    for level in range(L):
        G = torch.arange(T >> level, dtype=torch.float32)
        rows = min(halo >> level, 11)
        wins = []
        for s in shards:
            w0, w1 = s['win']
            x = G[w0 >> level:w1 >> level].clone()
            left, right = ts.halo_rows(s, level, rows, T)
            assert (left is None) == (s['rank'] == 0) and (right is None) == (s['rank'] == world - 1)
            if left is not None:
                x[left[0][0]:left[0][1]] = -1
            if right is not None:
                x[right[0][0]:right[0][1]] = -1
            wins.append((x, left, right))
        for s, (x, left, right) in zip(shards, wins):
            if left is not None:
                nb_x, _, nb_right = wins[s['rank'] - 1]
                x[left[0][0]:left[0][1]] = nb_x[nb_right[1][0]:nb_right[1][1]]
            if right is not None:
                nb_x, nb_left, _ = wins[s['rank'] + 1]
                x[right[0][0]:right[0][1]] = nb_x[nb_left[1][0]:nb_left[1][1]]
        for s, (x, _, _) in zip(shards, wins):
            w0, w1 = s['win']
            assert torch.equal(x, G[w0 >> level:w1 >> level]), (level, s['rank'])


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
        # grouped neighbour send/recv of one halo exchange (the NCCL path's host logic, here over gloo)
        s = shards[rank]
        lvl, rows = 2, 3
        Gl = torch.arange(T >> lvl, dtype=torch.float32)
        x = Gl[s['win'][0] >> lvl:s['win'][1] >> lvl].clone()[None, :, None].repeat(2, 1, 4)
        left, right = ts.halo_rows(s, lvl, rows, T)
        snd = lambda side: None if side is None else x[:, side[1][0]:side[1][1]].contiguous()
        rcv = lambda side: None if side is None else torch.empty(2, rows, 4)
        rl, rr = rcv(left), rcv(right)
        for side in (left, right):
            if side is not None:
                x[:, side[0][0]:side[0][1]] = -1
        sl, sr = (None if left is None else Gl[None, (s['win'][0] >> lvl) + left[1][0]:(s['win'][0] >> lvl) + left[1][1], None].repeat(2, 1, 4).contiguous(),
                  None if right is None else Gl[None, (s['win'][0] >> lvl) + right[1][0]:(s['win'][0] >> lvl) + right[1][1], None].repeat(2, 1, 4).contiguous())
        comm.exchange(sl, rl, sr, rr)
        if left is not None:
            x[:, left[0][0]:left[0][1]] = rl
        if right is not None:
            x[:, right[0][0]:right[0][1]] = rr
        assert torch.equal(x[0, :, 0], Gl[s['win'][0] >> lvl:s['win'][1] >> lvl]), rank
    finally:
        dist.destroy_process_group()


def test_allgather_assembly_gloo_world2():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, 48 * 16, 3), nprocs=2, join=True)

"""GPU: the time-sharded path (decaf_b200/time_shard.py) against the unsharded path on the same video.  With a window-sized
halo refreshed from the neighbours after every encoder output (halo_mode='exchange') — or a one-shot halo that covers the
whole receptive field ('recompute') — every owned point sees the same inputs in the same arithmetic, so the
merged candidates (global coordinates, global flat indices, order) and the final segments must equal the
unsharded ones; S shards are run one after the other on one GPU (emulate=S).  The multi-process form over NCCL
is exercised by tools/run_time_shard.py under torchrun on >= 2 GPUs."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(act_dtype, vid_len=2900, n_query=5):
    from decaf_b200 import synth
    from decaf_b200.worker_v2 import Evaluator, create_model
    opt = synth.tiny_opt(embd_dim=128, n_levels=5, win=9, max_seq_len=256, sn=12, vid_in_dim=64, text_dim=64,
                         pre_nms_topk=300)
    shapes = {k: tuple(v.shape) for k, v in create_model(opt.clone()).state_dict().items()}
    sd = synth.fill_state_dict(shapes, 5)
    data = synth.synth_video(opt, vid_len, n_query, seed=9, tag='long', n_events=2)
    ev = Evaluator(opt.clone(), dataset=[data], state_dict=sd, act_dtype=act_dtype, use_graphs=False)
    return opt, ev, data


@pytest.mark.parametrize('halo_mode', ['exchange', 'recompute'])
@pytest.mark.parametrize('act_dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('S', [2, 3, 5])
def test_sharded_equals_unsharded(act_dtype, S, halo_mode):
    from decaf_b200.time_shard import TimeShardedEvaluator
    opt, ev, data = _setup(act_dtype)
    ref = ev.predict_video(data)
    eng = ev.model.engine()
    T = ev.padded_len(data['vid'].size(-1))
    p = eng.plan(len(ref), T)
    ref_cnt = p.cand_count.cpu()
    ref_scores, ref_segs, ref_idx = p.cand_scores.cpu().clone(), p.cand_segs.cpu().clone(), p.cand_idx.cpu().clone()
    tse = TimeShardedEvaluator(ev, emulate=S, halo_mode=halo_mode)
    if halo_mode == 'exchange':          # pinned features take the direct strided upload (decaf_upload_2d), pageable ones the staging copy
        data = dict(data, vid=data['vid'].pin_memory(), shallow_vid=data['shallow_vid'].pin_memory())
    if tse.halo >= T // S - 16:
        pytest.skip('test video too short for this many shards with this halo')
    res, (m_segs, m_scores, m_idx, m_cnt) = tse.predict_video(data, return_candidates=True)
    assert torch.equal(m_cnt.cpu(), ref_cnt)
    for b in range(len(ref)):
        k = int(ref_cnt[b])
        assert k > 10
        assert torch.equal(m_idx[b, :k].cpu(), ref_idx[b, :k]), 'candidate order'
        np.testing.assert_allclose(m_scores[b, :k].cpu().numpy(), ref_scores[b, :k].numpy(), rtol=0, atol=0)
        np.testing.assert_allclose(m_segs[b, :k].cpu().numpy(), ref_segs[b, :k].numpy(), rtol=0, atol=0)
        assert res[b]['segments'].shape == ref[b]['segments'].shape
        np.testing.assert_allclose(res[b]['segments'].numpy(), ref[b]['segments'].numpy(), rtol=0, atol=0)
        np.testing.assert_allclose(res[b]['scores'].numpy(), ref[b]['scores'].numpy(), rtol=0, atol=0)


def test_mad_size_sharded_equals_unsharded():
    """BASELINE.json config 3 at full length: t = 70,001 clips (T = 71,424, P = 142,290 points per query), the NLQ network,
    3 queries, bf16; 4 and 8 time shards with per-layer halo exchange (1408-step halo, all shards in this process, advanced in
    lockstep) against the unsharded run: candidate order, scores, coordinates and final segments bit-identical."""
    from decaf_b200 import synth
    from decaf_b200.time_shard import TimeShardedEvaluator
    from decaf_b200.worker_v2 import Evaluator, create_model
    opt = synth.nlq_opt()
    shapes = {k: tuple(v.shape) for k, v in create_model(opt.clone()).state_dict().items()}
    sd = synth.fill_state_dict(shapes, 2022)
    data = synth.synth_video(opt, 70001, 3, seed=2022, tag='mad', n_events=2)
    ev = Evaluator(opt.clone(), dataset=[data], state_dict=sd, act_dtype=torch.bfloat16, use_graphs=False)
    ref = ev.predict_video(data)
    T = ev.padded_len(70001)
    assert T == 71424
    p = ev.model.engine().plan(3, T)
    ref_cnt, ref_idx = p.cand_count.cpu().clone(), p.cand_idx.cpu().clone()
    ref_scores, ref_segs = p.cand_scores.cpu().clone(), p.cand_segs.cpu().clone()
    for S in (4, 8):
        tse = TimeShardedEvaluator(ev, emulate=S)
        assert tse.halo_mode == 'exchange' and tse.halo == 1408
        res, (m_segs, m_scores, m_idx, m_cnt) = tse.predict_video(data, return_candidates=True)
        assert torch.equal(m_cnt.cpu(), ref_cnt)
        for b in range(3):
            k = int(ref_cnt[b])
            assert torch.equal(m_idx[b, :k].cpu(), ref_idx[b, :k])
            assert torch.equal(m_scores[b, :k].cpu(), ref_scores[b, :k]) and torch.equal(m_segs[b, :k].cpu(), ref_segs[b, :k])
            assert torch.equal(res[b]['segments'], ref[b]['segments']) and torch.equal(res[b]['scores'], ref[b]['scores'])
        del tse
        torch.cuda.empty_cache()


def test_halo_too_small_is_detectably_wrong():
    """Sanity of the tests themselves: with a 1-unit recompute halo, or with the window-sized halo but the exchange switched
    off, the shards do NOT reproduce the unsharded candidates; a halo too narrow for the exchange is refused."""
    from decaf_b200.time_shard import TimeShardedEvaluator
    opt, ev, data = _setup(torch.float32)
    ref = ev.predict_video(data)
    eng = ev.model.engine()
    p = eng.plan(len(ref), ev.padded_len(data['vid'].size(-1)))
    ref_scores = p.cand_scores.cpu().clone()
    tse = TimeShardedEvaluator(ev, emulate=2, halo=16, halo_mode='recompute')
    _, (m_segs, m_scores, m_idx, m_cnt) = tse.predict_video(data, return_candidates=True)
    assert not torch.equal(m_scores.cpu(), ref_scores)
    tse = TimeShardedEvaluator(ev, emulate=2)
    tse._exchange_step = lambda *a, **k: None
    _, (m_segs, m_scores, m_idx, m_cnt) = tse.predict_video(data, return_candidates=True)
    assert not torch.equal(m_scores.cpu(), ref_scores)
    with pytest.raises(AssertionError):
        TimeShardedEvaluator(ev, emulate=2, halo=16).predict_video(data)


def test_merge_kernel_matches_reference_rule():
    from decaf_b200 import _cabi as cabi
    from decaf_b200.time_shard import merge_candidates_reference
    g = torch.Generator().manual_seed(1)
    n_src, n, topk = 4, 6, 300
    scores = (torch.randint(0, 64, (n_src, n, topk), generator=g).float() / 64).sort(dim=-1, descending=True).values
    idx = torch.stack([torch.randperm(200000, generator=g)[:n_src * topk].reshape(n_src, topk) for _ in range(n)], 1).int()
    segs = torch.rand(n_src, n, topk, 2, generator=g)
    count = torch.randint(0, topk + 1, (n_src, n), generator=g).int()
    count[0, 0] = 0
    count[:, 1] = 0                                                      # a query without any candidate
    want = merge_candidates_reference(segs, scores, idx, count, topk)
    o_segs = torch.zeros(n, topk, 2, device='cuda'); o_scores = torch.zeros(n, topk, device='cuda')
    o_idx = torch.zeros(n, topk, dtype=torch.int32, device='cuda'); o_cnt = torch.zeros(n, dtype=torch.int32, device='cuda')
    cabi.merge_candidates(segs.cuda(), scores.cuda(), idx.cuda(), count.cuda(), n_src, n, topk, o_segs, o_scores, o_idx, o_cnt)
    for q in range(n):
        k = len(want[q][1])
        assert int(o_cnt[q]) == k
        assert torch.equal(o_idx[q, :k].cpu(), want[q][2])
        assert torch.equal(o_scores[q, :k].cpu(), want[q][1])
        assert torch.equal(o_segs[q, :k].cpu(), want[q][0])

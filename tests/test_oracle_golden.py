"""CPU: the oracle restatement (oracle/grounder_oracle.py) against golden vectors produced by
the unmodified reference (tests/golden/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest
import torch

from golden_util import CASES, load_case
from oracle import grounder_oracle as go
from oracle import nms_oracle


@pytest.mark.parametrize('name', list(CASES))
def test_oracle_matches_reference_golden(name):
    opt, sd, data, g = load_case(name)
    out = go.predict(sd, opt, data, return_aux=True)
    nq = int(g['n_query'])
    assert nq == len(out['logits'])
    for b in range(nq):
        lg = torch.cat([x[0] for x in out['logits'][b]]).numpy()
        of = torch.cat([x[0] for x in out['offsets'][b]]).numpy()
        mk = torch.cat([x.reshape(-1) for x in out['masks'][b]]).numpy().astype(np.uint8)
        np.testing.assert_allclose(out['text'][b][0].numpy(), g[f'text{b}'], rtol=0, atol=1e-5)
        np.testing.assert_allclose(out['aux'][b]['correl'].numpy(), g['correl'][b], rtol=0, atol=1e-6)
        # discrete: selection mask and level masks are exact
        assert np.array_equal(out['aux'][b]['weight'].numpy().astype(np.uint8), g['weight'][b])
        assert np.array_equal(mk, g[f'masks{b}'])
        np.testing.assert_allclose(lg, g[f'logits{b}'], rtol=0, atol=2e-5)
        np.testing.assert_allclose(of, g[f'offsets{b}'], rtol=0, atol=2e-5)
        l1 = torch.cat([x[0] for x in out['aux'][b]['logits1']]).numpy()
        np.testing.assert_allclose(l1, g[f'logits1_{b}'], rtol=0, atol=2e-5)
        np.testing.assert_allclose(out['aux'][b]['vid_map'][0].numpy(), g[f'vid_map{b}'], rtol=0, atol=1e-5)
        cs, cc, _ = out['cands'][b]
        assert cs.shape == g[f'cand_segs{b}'].shape
        np.testing.assert_allclose(cs.numpy(), g[f'cand_segs{b}'], rtol=0, atol=1e-4)
        np.testing.assert_allclose(cc.numpy(), g[f'cand_scores{b}'], rtol=0, atol=1e-5)
        r = out['results'][b]
        assert r['segments'].shape == g[f'res_segs{b}'].shape
        np.testing.assert_allclose(r['segments'].numpy(), g[f'res_segs{b}'], rtol=0, atol=1e-4)
        np.testing.assert_allclose(r['scores'].numpy(), g[f'res_scores{b}'], rtol=0, atol=1e-5)


@pytest.mark.parametrize('name', ['tiny_msf', 'small_w9', 'tiny_hardnms'])
def test_decode_and_nms_exact_given_reference_logits(name):
    """Feed the REFERENCE's logits/offsets through the oracle's decode + NMS (numpy twin and
    C twin): candidate order and final keep-set must be identical to the reference's."""
    opt, sd, data, g = load_case(name)
    T = int(g['T'])
    L = opt.model.num_fpn_levels
    sizes = [T // 2 ** l for l in range(L)]
    for b in range(int(g['n_query'])):
        lg = torch.from_numpy(g[f'logits{b}']).split(sizes)
        of = torch.from_numpy(g[f'offsets{b}']).split(sizes)
        mk = torch.from_numpy(g[f'masks{b}'].astype(bool)).split(sizes)
        lg = [x[None] for x in lg]
        of = [x[None] for x in of]
        mk = [x[None] for x in mk]
        segs, scores, idx = go.collect_segments(lg, of, mk, opt.eval.pre_nms_thresh,
                                                opt.eval.pre_nms_topk, opt.eval.seg_len_thresh)
        assert np.array_equal(segs.numpy(), g[f'cand_segs{b}'])
        assert np.array_equal(scores.numpy(), g[f'cand_scores{b}'])
        for fns in ((None, None), (nms_oracle.softnms, nms_oracle.nms)):
            s, c = go.batched_nms(segs, scores, softnms_fn=fns[0], nms_fn=fns[1], **opt.nms)
            s = (s * data['clip_stride'] + 0.5 * data['clip_size']) / data['fps']
            s = torch.clamp(s, min=0, max=data['duration'])
            np.testing.assert_allclose(s.numpy(), g[f'res_segs{b}'], rtol=0, atol=2e-6)
            if fns[0] is None:      # numpy exp vs glibc expf: ulp-level per decay
                np.testing.assert_allclose(c.numpy(), g[f'res_scores{b}'], rtol=2e-6, atol=0)
            else:                   # C twin: same libm, same op order -> bit-exact
                assert np.array_equal(c.numpy(), g[f'res_scores{b}'])


def test_select_clips_quirks():
    """k == 0 selects everything; partial last block; fp32 nearest index."""
    x = torch.arange(10, dtype=torch.float32)
    pooled, sel, w = go.select_clips(x, 10, 4, 0.0)
    assert np.allclose(pooled, [1.5, 5.5, 8.5]) and w.all() and len(sel) == 3
    pooled, sel, w = go.select_clips(x, 10, 4, 0.34)     # int(0.34*3) = 1 -> last block only
    assert list(sel) == [2]
    # int(0.29*100) == 28 in double arithmetic
    x = torch.rand(6000)
    pooled, sel, w = go.select_clips(x, 6000, 60, 0.29)
    assert len(sel) == 28
    # nearest up-sampling must equal torch's op for awkward lengths
    import torch.nn.functional as F
    for n in (123, 165, 550, 2000, 2301):
        x = torch.rand(n)
        pooled, sel, w = go.select_clips(x, n, 60, 0.3)
        m = len(pooled)
        ww = torch.zeros(m)
        ww[torch.from_numpy(sel)] = 1
        ref = F.interpolate(ww[None, None], size=n, mode='nearest')[0, 0].bool()
        assert torch.equal(ref, w), n
        cb = F.avg_pool1d(x[None, None], kernel_size=60, stride=60, ceil_mode=True)[0, 0]
        np.testing.assert_allclose(pooled, cb.numpy(), rtol=1e-6)


@pytest.mark.parametrize('name', list(CASES))
def test_oracle_eval_loss_matches_reference(name):
    """oracle eval_loss (the restatement of Evaluator._calc_loss, libs/worker_v2.py:1029-1061) fed with the reference's own
    logits / offsets / masks reproduces the reference's eval-time loss statistics recorded in the fixture."""
    opt, sd, data, g = load_case(name)
    nq, T, L = int(g['n_query']), int(g['T']), opt.model.num_fpn_levels
    sizes = [T // 2 ** l for l in range(L)]
    lg = [[x[None] for x in torch.from_numpy(g[f'logits{b}']).split(sizes)] for b in range(nq)]
    of = [[x[None] for x in torch.from_numpy(g[f'offsets{b}']).split(sizes)] for b in range(nq)]
    mk = [[x[None] for x in torch.from_numpy(g[f'masks{b}'].astype(bool)).split(sizes)] for b in range(nq)]
    rr = [tuple(x) for x in g['regression_range'].tolist()]
    got = go.eval_loss(opt, data, lg, of, mk, rr)
    assert abs(got['cls_loss'] - float(g['loss_cls'])) <= 1e-6 * abs(float(g['loss_cls']))
    assert abs(got['reg_loss'] - float(g['loss_reg'])) <= 1e-6 * abs(float(g['loss_reg']))

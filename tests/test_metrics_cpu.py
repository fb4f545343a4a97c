"""CPU: the vectorised R@k x IoU accumulation of Evaluator.run (all queries of a video at once, from the padded result
arrays) against the per-query statement of libs/worker_v2.py:857-878 / libs/train_utils.py:81-96 (Evaluator._accumulate),
including empty results and degenerate (zero-length) segments whose IoU is NaN."""
import numpy as np
import torch


def test_accumulate_raw_equals_per_query_statement():
    from decaf_b200.worker_v2 import Evaluator
    ev = Evaluator.__new__(Evaluator)                      # host logic only: no model, no device
    ev.ranks, ev.topk, ev.iou_threshs = (1, 5), 5, np.array((0.3, 0.5))
    rng = np.random.default_rng(0)
    B, K = 16, 5
    total = np.zeros((2, 2))
    for _ in range(40):
        cnt = rng.integers(0, K + 1, B).astype(np.int32)
        c = rng.uniform(0, 100, (B, K)).astype(np.float32)
        ln = rng.uniform(0, 30, (B, K)).astype(np.float32)
        segs = np.stack([c - ln / 2, c + ln / 2], -1).astype(np.float32)
        scores = -np.sort(-rng.uniform(0, 1, (B, K)).astype(np.float32), axis=1)      # rows sorted like the finalize kernel's
        tg = np.sort(rng.uniform(0, 100, (B, 2)).astype(np.float32), axis=1)
        segs[0, 0] = tg[0] = (5.0, 5.0)                    # union 0 -> NaN IoU in both statements
        ev.counts, ev.text_cnt = np.zeros((2, 2)), 0
        ev._accumulate_raw(segs, scores, cnt, tg)
        got, got_n = ev.counts.copy(), ev.text_cnt
        ev.counts, ev.text_cnt = np.zeros((2, 2)), 0
        res = [{'segments': torch.from_numpy(segs[b, :cnt[b]].copy()), 'scores': torch.from_numpy(scores[b, :cnt[b]].copy())}
               for b in range(B)]
        ev._accumulate(res, tg)
        assert np.array_equal(got, ev.counts) and got_n == ev.text_cnt == B
        total += got
    assert total.min() > 0

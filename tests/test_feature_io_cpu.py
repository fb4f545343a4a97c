"""CPU: feature-file ingest (decaf_b200/feature_io.py) against a direct statement of the reference's loading rules
(libs/data/dataset.py:128-135 VID_LOAD_FUNC, :363-407 / :840-891 _load_vid_feats / _load_shallow_vid_feats): formats, multi-source
alignment by tail replication, down-sampling, (c, t) layout, optional normalisation; the prefetching loader yields every item
once, in order, and recycles its buffers."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from decaf_b200 import feature_io as fio


def _write(tmp, name, arr, fmt):
    if fmt == 'npy':
        np.save(os.path.join(tmp, name + '.npy'), arr)
    else:
        torch.save(torch.from_numpy(arr), os.path.join(tmp, name + '.pt'))


def _records(n, rng, C_tok=12, Cs=8):
    recs = []
    for i in range(n):
        nq = 2 + i % 3
        recs.append({'id': f'v{i}', 'fps': 30.0, 'duration': 100.0 + i, 'num_frames': 3000, 'clip_size': 32, 'clip_stride': 16,
                     'segment': np.zeros((nq, 2), np.float32), 'target': torch.zeros(nq, 2),
                     'text': tuple(torch.randn(C_tok, 3 + j) for j in range(nq)), 'text_cls': torch.randn(nq, Cs)})
    return recs


@pytest.mark.parametrize('fmt', ['npy', 'pt'])
@pytest.mark.parametrize('normalize', [False, True])
def test_dataset_follows_the_reference_loading_rules(tmp_path, fmt, normalize):
    rng = np.random.default_rng(0)
    d1, d2, ds_ = (str(tmp_path / x) for x in ('rgb', 'flow', 'shallow'))
    for d in (d1, d2, ds_):
        os.makedirs(d)
    lens = [(40, 37), (64, 64), (25, 30)]
    raw = []
    for i, (a, b) in enumerate(lens):
        x1, x2 = rng.standard_normal((a, 6)).astype(np.float32), rng.standard_normal((b, 4)).astype(np.float64 if fmt == 'npy' else np.float32)
        sh = rng.standard_normal((max(a, b), 8)).astype(np.float32)
        _write(d1, f'v{i}', x1, fmt); _write(d2, f'v{i}', x2, fmt); _write(ds_, f'v{i}', sh, fmt)
        raw.append((x1, x2.astype(np.float32), sh))
    ds = fio.FeatureFileDataset(_records(3, rng), [d1, d2], [ds_], vid_load=fmt, shallow_load=fmt, downsample_rate=1, shallow_ds=1,
                                normalize_vid=normalize)
    assert len(ds) == 3
    for i, item in enumerate(ds):
        x1, x2, sh = raw[i]
        mx = max(len(x1), len(x2))
        pad = lambda x: x if len(x) == mx else np.concatenate((x, np.tile(x[-1], (mx - len(x), 1))))
        want = torch.from_numpy(np.ascontiguousarray(np.concatenate((pad(x1), pad(x2)), -1).transpose()))
        want_sh = torch.from_numpy(np.ascontiguousarray(sh.transpose()))
        if normalize:
            want, want_sh = F.normalize(want, dim=0), F.normalize(want_sh, dim=0)
        assert item['vid'].shape == (10, mx) and item['vid'].is_contiguous() and torch.equal(item['vid'], want)
        assert torch.equal(item['shallow_vid'], want_sh)
        assert item['clip_id'] == f'v{i}' and item['ext_scores'] is None and len(item['text']) == item['text_cls'].size(0)
        assert {'fps', 'num_frames', 'duration', 'segment', 'clip_size', 'clip_stride', 'target', 'text_id'} <= set(item)
        ds.release(item)


def test_downsampling_and_misalignment_and_missing_files(tmp_path):
    rng = np.random.default_rng(1)
    d, s = str(tmp_path / 'a'), str(tmp_path / 'b')
    os.makedirs(d); os.makedirs(s)
    x = rng.standard_normal((50, 4)).astype(np.float32)
    _write(d, 'v0', x, 'npy'); _write(s, 'v0', x[::2], 'npy')
    ds = fio.FeatureFileDataset(_records(1, rng), [d], [s], downsample_rate=2, shallow_ds=1)
    item = ds[0]
    assert torch.equal(item['vid'], torch.from_numpy(np.ascontiguousarray(x[::2].T))) and torch.equal(item['vid'], item['shallow_vid'])
    with pytest.raises(ValueError):
        fio.FeatureFileDataset([dict(_records(1, rng)[0], id='nope')], [d], [s])[0]
    d2 = str(tmp_path / 'c')
    os.makedirs(d2)
    _write(d2, 'v0', x[:30], 'npy')
    with pytest.raises(AssertionError):
        fio.load_feature_files('v0', [d, d2], 'npy')
    with pytest.raises(ValueError):
        fio.load_feature_files('v0', [d], 'pk1')


def test_prefetching_loader_yields_in_order_and_recycles(tmp_path):
    rng = np.random.default_rng(2)
    d, s = str(tmp_path / 'a'), str(tmp_path / 'b')
    os.makedirs(d); os.makedirs(s)
    arrs = []
    for i in range(12):
        x = rng.standard_normal((20 + 3 * i, 4)).astype(np.float32)
        _write(d, f'v{i}', x, 'npy'); _write(s, f'v{i}', x * 2, 'npy')
        arrs.append(x)
    ds = fio.FeatureFileDataset(_records(12, rng), [d], [s])
    seen = []
    for item in fio.PrefetchingLoader(ds, depth=2, lag=3):
        i = len(seen)
        assert item['clip_id'] == f'v{i}'
        assert torch.equal(item['vid'], torch.from_numpy(np.ascontiguousarray(arrs[i].T)))
        seen.append(item['clip_id'])
    assert seen == [f'v{i}' for i in range(12)]
    assert sum(len(v) for v in ds.pool._free.values()) >= 2        # buffers came back

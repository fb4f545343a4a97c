"""Feature-file ingest for the evaluation loop: per-video clip-feature files -> items with the
schema of libs/data/dataset.py:977-994, read ahead by a background thread straight into PINNED host buffers, so that
`Evaluator.predict_videos` uploads them without a staging copy and the loop is not host-bound at 10^4+ pairs/s.

What is mirrored from the reference's dataset (libs/data/dataset.py):
  * the per-format loaders of VID_LOAD_FUNC (:128-135): 'npy' (np.load(...).astype(float32)) and 'pt' (torch.load(...).numpy()),
    each file a (t, c) array;
  * _load_vid_feats / _load_shallow_vid_feats (:363-407, :840-891): several feature directories are concatenated along the
    channel axis after padding the shorter ones (<= 10 steps) by replicating their last vector, temporal down-sampling by a
    stride, transposition to (c, t), optional L2 normalisation over channels.
What is not: annotation parsing / tokenisation (the caller supplies, per video, the query token features, the query
class embeddings and the metadata fields), training-time cropping, external scores.

    ds = FeatureFileDataset(records, vid_dirs=[...], shallow_dirs=[...], vid_load='npy', shallow_load='npy')
    evaluator = Evaluator(opt, dataset=PrefetchingLoader(ds, depth=3), ...)
    evaluator.run()
"""
import os
import queue
import threading

import numpy as np
import torch
import torch.nn.functional as F


def _load_npy(path):
    return np.load(path + '.npy').astype(np.float32)


def _load_pt(path):
    return torch.load(path + '.pt').numpy().astype(np.float32)


VID_LOAD_FUNC = {'npy': _load_npy, 'pt': _load_pt}           # libs/data/dataset.py:128-135 (the formats that hold raw (t, c) arrays)


def load_feature_files(vid_id, dirs, fmt, downsample=1):
    """(t, c) float32: the features of one video from every directory in `dirs`, aligned and concatenated along channels
    (libs/data/dataset.py:363-398)."""
    if fmt not in VID_LOAD_FUNC:
        raise ValueError(f'unsupported feature format {fmt!r} (supported: {sorted(VID_LOAD_FUNC)})')
    try:
        feats = [VID_LOAD_FUNC[fmt](os.path.join(d, vid_id)) for d in dirs]
    except (OSError, ValueError) as e:
        raise ValueError(f'failed to load features for video {vid_id}: {e}')
    if len(feats) > 1:
        lens = [len(x) for x in feats]
        mx, mn = max(lens), min(lens)
        assert mx - mn <= 10, f'misaligned features ([max] {mx}, [min] {mn}) for video {vid_id}'
        feats = [x if len(x) == mx else np.concatenate((x, np.tile(x[-1], (mx - len(x), 1)))) for x in feats]
        out = np.concatenate(feats, axis=-1)
    else:
        out = feats[0]
    if downsample > 1:
        out = out[::downsample]
    return out


class _PinnedPool:
    """Reusable flat pinned float32 buffers in power-of-two capacity classes; a (C, t) item is the CONTIGUOUS view of the first
    C * t elements (what Evaluator's direct upload requires).  Falls back to pageable memory where no CUDA runtime exists."""

    def __init__(self):
        self._free = {}
        self.pinned = torch.cuda.is_available()

    def get(self, C, t):
        cap = max(1 << 16, 1 << (int(C * t) - 1).bit_length())
        lst = self._free.setdefault(cap, [])
        buf = lst.pop() if lst else (torch.empty(cap).pin_memory() if self.pinned else torch.empty(cap))
        return buf, cap

    def put(self, buf, key):
        self._free.setdefault(key, []).append(buf)


class FeatureFileDataset:
    """records: list of dicts, one per video: {'id', 'fps', 'duration', 'num_frames', 'segment' (n, 2) seconds, 'target' (n, 2)
    grid units, 'text': tuple of n (C_tok, L_i) tensors, 'text_cls': (n, C_s), 'clip_size', 'clip_stride'} — everything of the
    reference item (libs/data/dataset.py:977-994) except the two feature tensors, which are read here."""

    def __init__(self, records, vid_dirs, shallow_dirs, vid_load='npy', shallow_load='npy', downsample_rate=1, shallow_ds=1,
                 normalize_vid=False, pool=None):
        self.records = list(records)
        self.vid_dirs, self.shallow_dirs = list(vid_dirs), list(shallow_dirs)
        self.vid_load, self.shallow_load = vid_load, shallow_load
        self.downsample_rate, self.shallow_ds, self.normalize_vid = int(downsample_rate), int(shallow_ds), bool(normalize_vid)
        self.pool = pool or _PinnedPool()

    def __len__(self):
        return len(self.records)

    def _to_ct(self, arr):
        """(t, c) numpy -> (c, t) view of a pinned buffer (one transposing copy), + the handle that returns the buffer."""
        t, c = arr.shape
        buf, key = self.pool.get(c, t)
        view = buf[:c * t].view(c, t)
        view.copy_(torch.from_numpy(arr).t())
        if self.normalize_vid:
            view.copy_(F.normalize(view, dim=0))
        return view, (buf, key)

    def __getitem__(self, idx):
        rec = self.records[idx]
        vid = load_feature_files(rec['id'], self.vid_dirs, self.vid_load, self.downsample_rate)
        sh = load_feature_files(rec['id'], self.shallow_dirs, self.shallow_load, self.shallow_ds)
        assert len(vid) == len(sh), f"expert / sidekick features of {rec['id']} differ in length ({len(vid)} vs {len(sh)})"
        v, hv = self._to_ct(vid)
        s, hs = self._to_ct(sh)
        item = {k: rec[k] for k in ('fps', 'num_frames', 'duration', 'segment', 'clip_size', 'clip_stride', 'target', 'text', 'text_cls')
                if k in rec}
        item.update({'clip_id': rec['id'], 'text_id': rec.get('text_id', list(range(len(rec['text'])))), 'vid': v, 'shallow_vid': s,
                     'ext_scores': None, '_buffers': (hv, hs)})
        return item

    def release(self, item):
        """Hand the item's pinned buffers back (call once its results are out; PrefetchingLoader does it automatically)."""
        for buf, key in item.pop('_buffers', ()):
            self.pool.put(buf, key)

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]


class PrefetchingLoader:
    """Iterates a FeatureFileDataset with a reader thread `depth` videos ahead (file reads, the transposing copy into pinned
    memory and np.load all release the GIL).  An item's pinned buffers are recycled `lag` items after it was handed out —
    `lag` must exceed the number of videos the consumer keeps in flight (Evaluator: n_lanes + 2 host slots)."""

    def __init__(self, dataset, depth=3, lag=8):
        self.ds, self.depth, self.lag = dataset, max(1, int(depth)), max(1, int(lag))

    def __len__(self):
        return len(self.ds)

    def __iter__(self):
        q = queue.Queue(maxsize=self.depth)
        stop = threading.Event()

        def reader():
            try:
                for i in range(len(self.ds)):
                    if stop.is_set():
                        return
                    q.put(self.ds[i])
                q.put(None)
            except BaseException as e:            # surfaces in the consumer
                q.put(e)
        th = threading.Thread(target=reader, daemon=True)
        th.start()
        handed = []
        try:
            while True:
                item = q.get()
                if item is None:
                    break
                if isinstance(item, BaseException):
                    raise item
                handed.append(item)
                if len(handed) > self.lag:
                    self.ds.release(handed.pop(0))
                yield item
        finally:
            stop.set()
            while not q.empty():
                q.get_nowait()
            if torch.cuda.is_available():
                torch.cuda.synchronize()          # nothing may still be uploading out of the buffers about to be recycled
            for it in handed:
                self.ds.release(it)

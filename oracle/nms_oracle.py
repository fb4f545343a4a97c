"""ORACLE (test infrastructure, NOT product code): ctypes front-end of oracle/nms_oracle.c.

Function signatures match the plug points of oracle/grounder_oracle.batched_nms
(``softnms_fn`` / ``nms_fn``).  ``reference_fns()`` returns the same two callables backed by
the compiled, unmodified reference extension (oracle/_ref), when it exists.
"""
import ctypes
import os

import numpy as np

from . import build_ref

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = build_ref.build_c_oracle()
        L = ctypes.CDLL(path)
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int64)
        L.oracle_softnms_1d.restype = ctypes.c_int64
        L.oracle_softnms_1d.argtypes = [fp, fp, ctypes.c_int64, fp, ip, ctypes.c_float,
                                        ctypes.c_float, ctypes.c_float, ctypes.c_int,
                                        ctypes.c_int64]
        L.oracle_nms_1d.restype = ctypes.c_int64
        L.oracle_nms_1d.argtypes = [fp, fp, ctypes.c_int64, ip, ctypes.c_float]
        L.oracle_decode.restype = ctypes.c_int64
        L.oracle_decode.argtypes = [fp, fp, fp, fp, ctypes.c_int64, ctypes.c_float,
                                    ctypes.c_int64, ctypes.c_float, fp, fp, ip]
        _lib = L
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def softnms(segs, scores, iou_thresh, sigma, min_score, method, max_iters=None):
    segs, ps = _f(segs)
    scores, pc = _f(scores)
    n = len(scores)
    dets = np.zeros((max(n, 1), 3), dtype=np.float32)
    inds = np.zeros(max(n, 1), dtype=np.int64)
    k = lib().oracle_softnms_1d(ps, pc, n, dets.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                                inds.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                iou_thresh, sigma, min_score, method,
                                -1 if max_iters is None else int(max_iters))
    return dets[:k], inds[:k]


def nms(segs, scores, iou_thresh):
    segs, ps = _f(segs)
    scores, pc = _f(scores)
    n = len(scores)
    keep = np.zeros(max(n, 1), dtype=np.int64)
    k = lib().oracle_nms_1d(ps, pc, n, keep.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), iou_thresh)
    return keep[:k]


def decode(scores, offsets, coord, stride, pre_nms_thresh, topk, seg_len_thresh):
    scores, p0 = _f(scores)
    offsets, p1 = _f(offsets)
    coord, p2 = _f(coord)
    stride, p3 = _f(stride)
    p = len(scores)
    segs = np.zeros((max(topk, 1), 2), dtype=np.float32)
    sc = np.zeros(max(topk, 1), dtype=np.float32)
    idx = np.zeros(max(topk, 1), dtype=np.int64)
    k = lib().oracle_decode(p0, p1, p2, p3, p, pre_nms_thresh, topk, seg_len_thresh,
                            segs.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                            sc.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                            idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
    return segs[:k], sc[:k], idx[:k]


def reference_fns():
    """(softnms_fn, nms_fn) backed by the compiled reference extension, or (None, None)."""
    mod = build_ref.load_reference_nms()
    if mod is None:
        return None, None
    import torch

    def ref_softnms(segs, scores, iou_thresh, sigma, min_score, method, max_iters=None):
        s = torch.from_numpy(np.ascontiguousarray(segs, dtype=np.float32))
        c = torch.from_numpy(np.ascontiguousarray(scores, dtype=np.float32))
        dets = torch.zeros((len(c), 3), dtype=torch.float32)
        inds = mod.softnms(s, c, dets, iou_thresh=float(iou_thresh), sigma=float(sigma),
                           min_score=float(min_score), method=int(method))
        k = len(inds) if max_iters is None else min(len(inds), int(max_iters))
        return dets[:k].numpy(), inds[:k].numpy()

    def ref_nms(segs, scores, iou_thresh):
        s = torch.from_numpy(np.ascontiguousarray(segs, dtype=np.float32))
        c = torch.from_numpy(np.ascontiguousarray(scores, dtype=np.float32))
        return mod.nms(s, c, iou_thresh=float(iou_thresh)).numpy()

    return ref_softnms, ref_nms

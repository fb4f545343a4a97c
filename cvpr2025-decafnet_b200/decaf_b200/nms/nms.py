"""Mirror of libs/nms/nms.py on top of the CUDA NMS kernels (C ABI: decaf_softnms_1d,
decaf_nms_1d, decaf_batched_nms in include/decaf_b200.h).

`nms_1d_gpu` plays the role of the reference's pybind module `nms_1d_cpu_vg`
(libs/nms/src/nms_cpu.cpp:184-194) with the same two functions and argument meaning; its inputs must be
CUDA tensors (the reference's extension insists on CPU tensors, nms_cpu.cpp:11-17 — here the check is
mirrored: CPU tensors raise).  The Python-level entry points (`batched_nms`, `NMSop`, `SoftNMSop`) are the drop-in
for libs/nms/nms.py and accept what the reference's call site passes — CPU tensors (libs/worker_v2.py:1083-1111)
— as well as CUDA tensors: CPU inputs are uploaded to the current CUDA device, the kernels run there (there is no
CPU implementation), and the results come back on the input's device.
"""
import torch

from .. import _cabi as cabi


def _to_cuda(x):
    """The tensor on the current CUDA device (a no-op for CUDA inputs), contiguous."""
    if not torch.cuda.is_available():
        raise RuntimeError('decaf_b200.nms needs a CUDA device (no CPU fallback exists)')
    return (x if x.is_cuda else x.cuda()).contiguous()


def _check_cuda(x, name):
    if not x.is_cuda:
        raise RuntimeError(f'{name} must be a CUDA tensor')
    if not x.is_contiguous():
        raise RuntimeError(f'{name} must be contiguous')


class nms_1d_gpu:
    """Drop-in for the `nms_1d_cpu_vg` module."""

    @staticmethod
    def nms(segs, scores, iou_thresh):
        """Kept indices (int64) in descending score order.  nms_cpu.cpp:20-70."""
        _check_cuda(segs, 'segs'); _check_cuda(scores, 'scores')
        n = scores.numel()
        if n == 0:
            return torch.empty(0, dtype=torch.long, device=segs.device)
        cnt = torch.tensor([n], dtype=torch.int32, device=segs.device)
        keep = torch.empty(n, dtype=torch.int32, device=segs.device)
        n_out = torch.zeros(1, dtype=torch.int32, device=segs.device)
        cabi.nms_1d(segs.float(), scores.float(), cnt, 1, n, keep, n_out, float(iou_thresh), 0.0, 0)
        return keep[:int(n_out.item())].long()

    @staticmethod
    def softnms(segs, scores, dets, iou_thresh, sigma, min_score, method, max_iters=0):
        """Writes dets (n,3) in place, returns original indices (int64).  nms_cpu.cpp:72-181.
        `max_iters` > 0 (extension): stop after that many outer steps."""
        _check_cuda(segs, 'segs'); _check_cuda(scores, 'scores'); _check_cuda(dets, 'dets')
        n = scores.numel()
        if n == 0:
            return torch.empty(0, dtype=torch.long, device=segs.device)
        assert dets.dtype == torch.float32 and dets.shape == (n, 3)
        cnt = torch.tensor([n], dtype=torch.int32, device=segs.device)
        inds = torch.empty(n, dtype=torch.int32, device=segs.device)
        n_out = torch.zeros(1, dtype=torch.int32, device=segs.device)
        ws = torch.empty(int(cabi.nms_workspace_bytes(1, n)), dtype=torch.uint8, device=segs.device)
        cabi.softnms_1d(segs.float(), scores.float(), cnt, 1, n, dets, inds, n_out, float(iou_thresh), float(sigma),
                        float(min_score), int(method), int(max_iters), ws)
        return inds[:int(n_out.item())].long()


class NMSop(torch.autograd.Function):
    """libs/nms/nms.py:6-31."""

    @staticmethod
    def forward(ctx, segs, scores, iou_thresh, min_score, max_num_segs):
        dev = segs.device
        segs, scores = _to_cuda(segs), _to_cuda(scores)
        if min_score > 0:
            mask = scores > min_score
            segs, scores = segs[mask], scores[mask]
        idx = nms_1d_gpu.nms(segs.contiguous(), scores.contiguous(), iou_thresh=float(iou_thresh))
        if max_num_segs > 0:
            idx = idx[:min(max_num_segs, len(idx))]
        return segs[idx].contiguous().to(dev), scores[idx].contiguous().to(dev)


class SoftNMSop(torch.autograd.Function):
    """libs/nms/nms.py:34-61."""

    @staticmethod
    def forward(ctx, segs, scores, iou_thresh, sigma, min_score, method, max_num_segs):
        dev = segs.device
        segs, scores = _to_cuda(segs), _to_cuda(scores)
        out = segs.new_empty((len(segs), 3))
        idx = nms_1d_gpu.softnms(segs.contiguous(), scores.contiguous(), out, iou_thresh=float(iou_thresh),
                                 sigma=float(sigma), min_score=float(min_score), method=int(method),
                                 max_iters=max(int(max_num_segs), 0))
        num_segs = len(idx)
        if max_num_segs > 0:
            num_segs = min(num_segs, max_num_segs)
        return out[:num_segs, :2].contiguous().to(dev), out[:num_segs, 2].contiguous().to(dev)


def batched_nms(segs, scores, iou_thresh, min_score, max_num_segs, mode='soft_nms', sigma=0.5,
                voting_thresh=0.75):
    """libs/nms/nms.py:106-148, same signature and return convention ((k,2) segs, (k,) scores;
    `zeros(0,2), zeros(0)` when empty).  One fused launch pair per call: (soft-)NMS + voting + sort.  CPU or CUDA
    tensors in, results on the input's device (the empty result too)."""
    in_dev = segs.device
    if len(segs) == 0:
        return torch.zeros(0, 2, device=in_dev), torch.zeros(0, device=in_dev)
    if mode not in (None, 'nms', 'soft_nms'):
        raise NotImplementedError('invalid NMS mode')
    segs, scores = _to_cuda(segs.float()), _to_cuda(scores.float())
    n = scores.numel()
    dev = segs.device
    prm = cabi.NmsParams()
    prm.mode = {None: 0, 'nms': 1, 'soft_nms': 2}[mode]
    prm.iou_thresh, prm.sigma, prm.min_score = float(iou_thresh), float(sigma), float(min_score)
    prm.max_num_segs, prm.voting_thresh = int(max_num_segs), float(voting_thresh)
    max_out = int(max_num_segs) if max_num_segs > 0 else n
    cnt = torch.tensor([n], dtype=torch.int32, device=dev)
    out_segs = torch.empty(1, max_out, 2, device=dev)
    out_scores = torch.empty(1, max_out, device=dev)
    out_count = torch.zeros(1, dtype=torch.int32, device=dev)
    ws = torch.empty(int(cabi.nms_workspace_bytes(1, n)), dtype=torch.uint8, device=dev)
    cabi.batched_nms(segs, scores, cnt, 1, n, prm, out_segs, out_scores, out_count, ws)
    k = int(out_count.item())
    return out_segs[0, :k].to(in_dev), out_scores[0, :k].to(in_dev)

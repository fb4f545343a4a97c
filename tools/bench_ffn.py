"""Fused FFN launch vs the fc + proj GEMM pair at the shapes of one NLQ step (CUDA events, warm, 20 repetitions).
    python tools/bench_ffn.py [C] [n_query]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'cvpr2025-decafnet_b200'))
from decaf_b200 import _cabi as cabi  # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 256
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dev = 'cuda'
g = torch.Generator(device=dev).manual_seed(0)
W1 = (torch.randn(4 * C, C, generator=g, device=dev) * C ** -0.5).to(torch.bfloat16)
W2 = (torch.randn(C, 4 * C, generator=g, device=dev) * (4 * C) ** -0.5).to(torch.bfloat16)
b1, b2, ls = torch.randn(4 * C, device=dev) * 0.1, torch.randn(C, device=dev) * 0.1, torch.rand(C, device=dev)
def timeit(fns, reps=5):
    """fns: the same launch on len(fns) different buffer sets (together larger than L2), captured back to back in one CUDA
    graph (no host launch cost between them); returns us per launch."""
    for f in fns:
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for f in fns:
            f()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * len(fns)) * 1e3


print(f'# C={C}, {B} queries; us per launch (CUDA-graph replay over rotating buffer sets > L2)')
for T in (2304, 1152, 576, 288, 144, 72, 36, 18):
    M = B * T
    nb = max(2, min(8, (300 << 20) // (M * C * 14)))
    sets = []
    for _ in range(nb):
        sets.append((torch.randn(M, C, generator=g, device=dev).to(torch.bfloat16), torch.randn(M, C, generator=g, device=dev),
                     torch.empty(M, 4 * C, dtype=torch.bfloat16, device=dev), torch.empty(M, C, device=dev)))
    mask = torch.ones(M, dtype=torch.uint8, device=dev)

    def unfused(A, X, H, O):
        def f():
            cabi.gemm(A, W1, 4 * C, C, 1, M, bias=b1, act=cabi.ACT_GELU, out_act=H)
            cabi.gemm(H, W2, C, 4 * C, 1, M, bias=b2, colscale=ls, resid=X, rowmask=mask, out_f32=O)
        return f

    def fused(A, X, H, O):
        return lambda: cabi.ffn(A, W1, b1, W2, b2, C, 1, M, colscale=ls, resid=X, rowmask=mask, out_f32=O)

    tu, tf = timeit([unfused(*s) for s in sets]), timeit([fused(*s) for s in sets])
    fl = 16.0 * M * C * C
    print(f'M={M:6d}  gemm pair {tu:7.1f} us   fused {tf:7.1f} us   ({fl / tf / 1e6:7.1f} TFLOP/s fused)')
    del sets

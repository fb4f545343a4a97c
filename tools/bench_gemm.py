"""Times decaf_gemm in isolation on the shapes of the NLQ step (CUDA events, L2 flushed between launches
by rotating over several operand sets) and prints achieved GB/s and TFLOP/s per shape."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'cvpr2025-decafnet_b200')):
    sys.path.insert(0, p)
import torch
from decaf_b200 import _cabi as cabi

B, T, C = 16, 2304, 256
rows = B * T
hrows = B * 4599
SHAPES = [
    # name, rows, K, N, taps, groups, resid, f32out, actout, gelu, ln
    ('qkv g3', rows, C, C, 1, 3, False, False, True, False, False),
    ('proj+resid', rows, C, C, 1, 1, True, True, False, False, False),
    ('fc gelu', rows, C, 4 * C, 1, 1, False, False, True, True, False),
    ('proj2+resid', rows, 4 * C, C, 1, 1, True, True, True, False, False),
    ('vid_map', rows, 512, C, 1, 1, False, True, False, False, False),
    ('head conv3 256 ln', hrows, C, C, 3, 1, False, False, True, False, True),
    ('head conv3 288 ln', hrows, 288, 288, 3, 1, False, False, True, False, True),
    ('head conv3 288 f32', hrows, 288, 288, 3, 1, False, True, False, False, False),
    ('lvl7 proj2', B * 18, 4 * C, C, 1, 1, True, True, True, False, False),
]
NSET = 6


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else None
    for name, M, K, N, taps, G, resid, f32o, acto, gelu, ln in SHAPES:
        if only and only not in name:
            continue
        sets = []
        for s in range(NSET):
            A = torch.randn(G, M, K, device='cuda').bfloat16()
            W = (torch.randn(G, N, taps, K, device='cuda') / (K * taps) ** 0.5).bfloat16()
            bias = torch.randn(G, N, device='cuda')
            r = torch.randn(M, N, device='cuda') if resid else None
            o32 = torch.empty(G, M, N, device='cuda') if f32o else None
            oa = torch.empty(G, M, N, device='cuda', dtype=torch.bfloat16) if acto else None
            lw = torch.randn(N, device='cuda')
            sets.append((A, W, bias, r, o32, oa, lw))

        def run(s):
            A, W, bias, r, o32, oa, lw = sets[s % NSET]
            cabi.gemm(A, W, N, K, 1, M, taps=taps, bias=bias, act=cabi.ACT_GELU if gelu else (cabi.ACT_RELU if ln else 0),
                      resid=r, out_f32=o32, out_act=oa, n_group=G, g_stride_a=M * K, g_stride_w=N * taps * K,
                      g_stride_bias=N, g_stride_out_f32=M * N, g_stride_out_act=M * N, ln=ln, ln_w=lw if ln else None,
                      ln_b=lw if ln else None, impl=2)
        for i in range(3):
            run(i)
        torch.cuda.synchronize()
        # one CUDA graph of 2 * NSET launches: the host cost of a launch (ctypes + tensor-map encode, ~30 us) would
        # otherwise hide kernels shorter than that
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for i in range(2 * NSET):
                run(i)
        graph.replay()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); graph.replay(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / (2 * NSET))
        ms = sorted(ts)[len(ts) // 2]
        nbytes = G * (M * K * 2 + M * N * ((4 if f32o else 0) + (2 if acto else 0) + (4 if resid else 0)))
        flops = 2.0 * G * M * N * K * taps
        print(f'{name:22s} M={M:6d} K={K:4d} N={N:4d} taps={taps} G={G}: {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:7.0f} GB/s  '
              f'{flops / ms / 1e9:7.1f} TFLOP/s')


if __name__ == '__main__':
    main()

"""Debug: per-role clock64 timeline of CTA 0 of one fused FFN launch (decaf_debug_ffn_trace).
    python tools/trace_ffn.py [M] [C]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'cvpr2025-decafnet_b200'))
import torch
from decaf_b200 import _cabi as cabi

M = int(sys.argv[1]) if len(sys.argv) > 1 else 288
C = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = 'cuda'
A = torch.randn(M, C, device=dev).bfloat16()
W1 = (torch.randn(4 * C, C, device=dev) * C ** -0.5).bfloat16()
W2 = (torch.randn(C, 4 * C, device=dev) * (4 * C) ** -0.5).bfloat16()
b1, b2, ls = torch.randn(4 * C, device=dev) * 0.1, torch.randn(C, device=dev) * 0.1, torch.rand(C, device=dev)
X, O = torch.randn(M, C, device=dev), torch.empty(M, C, device=dev)
run = lambda: cabi.ffn(A, W1, b1, W2, b2, C, 1, M, colscale=ls, resid=X, out_f32=O)
for _ in range(3):
    run()
torch.cuda.synchronize()
buf = torch.zeros(3, 512, dtype=torch.int64, device=dev)
cabi.debug_ffn_trace(buf.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record()
torch.cuda.synchronize()
cabi.debug_ffn_trace(None)
print(f'M={M} C={C}: event time {e0.elapsed_time(e1) * 1e3:.1f} us')
b = buf.cpu()
t0 = int(b[b > 0].min())
names = ['producer: W stage acquired', 'mma thread: [G1 stage full x2] / [G2 wait H, H ready, stage full x2]',
         'epilogue warp 0: per slice [wait acc1, acc1 ready, H free, done]; per tile [acc2 ready, E2 done]']
for r in range(3):
    v = [int(x) - t0 for x in b[r] if x > 0]
    print(names[r], len(v))
    print('  ', v[:120])

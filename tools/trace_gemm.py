"""Debug: per-role clock64 timeline of CTA 0 of one tcgen05 GEMM launch (decaf_debug_gemm_trace)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'cvpr2025-decafnet_b200')):
    sys.path.insert(0, p)
import torch
from decaf_b200 import _cabi as cabi

M, K, N, taps, G = [int(x) for x in (sys.argv[1:6] if len(sys.argv) > 5 else (36864, 256, 256, 1, 1))]
mode = sys.argv[6] if len(sys.argv) > 6 else 'plain'
A = torch.randn(G, M, K, device='cuda').bfloat16()
W = (torch.randn(G, N, taps, K, device='cuda') / (K * taps) ** 0.5).bfloat16()
bias = torch.randn(G, N, device='cuda')
oa = torch.empty(G, M, N, device='cuda', dtype=torch.bfloat16)
o32 = torch.empty(M, N, device='cuda')
cs = torch.randn(N, device='cuda')
kw = dict(out_act=oa)
if mode == 'gelu':
    kw = dict(out_act=oa, act=cabi.ACT_GELU)
elif mode == 'resid':
    kw = dict(out_f32=o32, resid=o32, colscale=cs)
elif mode == 'resid2':
    kw = dict(out_f32=o32, resid=torch.randn(M, N, device='cuda'), colscale=cs)
elif mode == 'ln':
    kw = dict(out_act=oa, ln=True, ln_w=cs, ln_b=cs, act=cabi.ACT_RELU)


def run():
    cabi.gemm(A, W, N, K, 1, M, taps=taps, bias=bias, n_group=G, g_stride_a=M * K, g_stride_w=N * taps * K, g_stride_bias=N,
              g_stride_out_act=M * N, impl=2, **kw)


for _ in range(3):
    run()
torch.cuda.synchronize()
# evict the operands from L2 (126 MB) so the traced launch sees HBM like a real step does
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device='cuda')
flush.fill_(1)
torch.cuda.synchronize()
buf = torch.zeros(4, 2048, dtype=torch.int64, device='cuda')
cabi.debug_gemm_trace(buf.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record()
torch.cuda.synchronize()
print(f'event time {e0.elapsed_time(e1) * 1e3:.1f} us')
cabi.debug_gemm_trace(None)
b4 = buf.cpu()[3]
b = buf.cpu()[:3]
t0 = int(b[b > 0].min())
names = ['producer: stage acquired', 'mma: stage full', 'epilogue team 0 leader: setup|acc ready|(chunk read, chunk stored)*']
for r in range(3):
    v = [int(x) - t0 for x in b[r] if x > 0]
    print(names[r], len(v))
    print('  ', v[:64])
    if len(v) > 64:
        print('   ...', v[-16:])
v = [int(x) - t0 for x in b4 if x > 0]
if v:
    print('mma thread per k-block: [loop top, stage full, MMAs issued, committed] (deltas to the previous stamp)')
    for i in range(0, min(len(v), 4 * 24), 4):
        q = v[i:i + 4]
        prev = v[i - 1] if i else q[0]
        print('   ', q[0] - prev, [q[j] - q[j - 1] for j in range(1, len(q))])

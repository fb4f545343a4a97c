"""Debug: every decaf_ffn call of one grounder pass is re-run as the fc + proj GEMM pair on the same inputs (copies of the
in-place residual) and the two results are compared call by call."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'cvpr2025-decafnet_b200')):
    sys.path.insert(0, p)
import torch
from decaf_b200 import _cabi as cabi, synth
from decaf_b200.worker_v2 import Evaluator, create_model

opt = synth.tiny_opt(embd_dim=128, n_levels=5, win=9, max_seq_len=256, sn=12, vid_in_dim=64, text_dim=64)
shapes = {k: tuple(v.shape) for k, v in create_model(opt.clone()).state_dict().items()}
sd = synth.fill_state_dict(shapes, 21)
data = synth.synth_video(opt, 230, 4, seed=21, tag='ffn', n_events=1)
ev = Evaluator(opt.clone(), dataset=[data], state_dict=sd, act_dtype=torch.bfloat16, use_graphs=False)
eng = ev.model.engine()
eng.fused_ffn, eng.ffn_min_rows = True, 0
real_ffn = cabi.ffn
n_call = [0]


def checked_ffn(A, W1, b1, W2, b2, C, n_seq, rows, colscale=None, resid=None, rowmask=None, m_seq_stride=0, out_f32=None,
                out_act=None, ldo2=0, o2_seq_stride=0, **kw):
    M = n_seq * rows
    r0 = resid.clone() if resid is not None else None
    ref32 = torch.zeros(M, C, device='cuda')
    H = torch.empty(M, 4 * C, dtype=torch.bfloat16, device='cuda')
    cabi.gemm(A, W1, 4 * C, C, 1, M, bias=b1, act=cabi.ACT_GELU, out_act=H)
    cabi.gemm(H, W2, C, 4 * C, n_seq, rows, bias=b2, colscale=colscale, resid=r0, rowmask=rowmask, m_seq_stride=m_seq_stride, out_f32=ref32)
    real_ffn(A, W1, b1, W2, b2, C, n_seq, rows, colscale=colscale, resid=resid, rowmask=rowmask, m_seq_stride=m_seq_stride,
             out_f32=out_f32, out_act=out_act, ldo2=ldo2, o2_seq_stride=o2_seq_stride, **kw)
    torch.cuda.synchronize()
    got = out_f32.view(-1)[:M * C].view(M, C)
    d = (got - ref32).abs()
    # second opinion: torch on the same operands
    h = torch.nn.functional.gelu(A.view(-1)[:M * C].view(M, C).float() @ W1.float().view(4 * C, C).t() + b1, approximate='tanh').to(torch.bfloat16)
    dh = (h.float() - H.float()).abs().max().item()
    print(f'call {n_call[0]:2d}  n_seq {n_seq} rows {rows} M {M}: max |fused - pair| {d.max().item():.3e} at row {int(d.max(1).values.argmax())} '
          f'(scale {ref32.abs().max().item():.2f}); H(pair) vs torch {dh:.3e}; A finite {bool(torch.isfinite(A.float()).all())}')
    n_call[0] += 1


cabi.ffn = checked_ffn
import decaf_b200.engine as E
E.cabi.ffn = checked_ffn
ev.predict_video(data)

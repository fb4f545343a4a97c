"""GPU: the fused FFN launch (decaf_ffn: fc -> GELU -> proj with LayerScale / residual / mask, hidden tensor kept on the SM;
libs/modeling/blocks.py:523-538, 587-590) against (a) the two decaf_gemm launches it replaces — same arithmetic in the same
accumulation order, so the results are expected to be equal bit for bit — and (b) a plain torch fp32 statement of the same op
on the same bf16-rounded operands."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(C, n_seq, rows, seq_pad, ld_out2, seed):
    g = torch.Generator(device='cuda').manual_seed(seed)
    M = n_seq * rows
    r = lambda *s, scale=1.0: torch.randn(*s, generator=g, device='cuda') * scale
    A = r(M, C).to(torch.bfloat16)
    W1 = r(4 * C, C, scale=C ** -0.5).to(torch.bfloat16)
    W2 = r(C, 4 * C, scale=(4 * C) ** -0.5).to(torch.bfloat16)
    b1, b2, ls = r(4 * C, scale=0.1), r(C, scale=0.1), torch.rand(C, generator=g, device='cuda') * 0.5 + 0.5
    resid = r(M, C)
    stride = rows + seq_pad
    mask = (torch.rand(n_seq, stride, generator=g, device='cuda') > 0.2).to(torch.uint8)
    return A, W1, b1, W2, b2, ls, resid, mask, stride


CASES = [
    (256, 1, 36864, 0, 256, 'nlq level 0, flat'),
    (256, 16, 288, 37, 288, 'level 3, strided mask / bf16 copy into a wider buffer'),
    (256, 16, 18, 5, 288, 'level 7: tiles span several sequences'),
    (256, 3, 100, 0, 256, 'odd tile count: phantom partner tile, partial last tile'),
    (256, 1, 128, 0, 256, 'one tile'),
    (256, 5, 77, 3, 264, 'ragged'),
    (128, 8, 256, 0, 160, 'embd 128 (Charades shape)'),
    (128, 7, 33, 2, 128, 'embd 128 ragged'),
]


@pytest.mark.parametrize('C,n_seq,rows,seq_pad,ld2,what', CASES, ids=[c[-1] for c in CASES])
def test_fused_ffn_equals_gemm_pair_and_torch(C, n_seq, rows, seq_pad, ld2, what):
    from decaf_b200 import _cabi as cabi
    assert cabi.ffn_supported(C, cabi.BF16)
    A, W1, b1, W2, b2, ls, resid, mask, stride = _mk(C, n_seq, rows, seq_pad, ld2, seed=C + rows)
    M = n_seq * rows
    # (a) the unfused pair
    H = torch.empty(M, 4 * C, dtype=torch.bfloat16, device='cuda')
    ref32 = torch.full((M, C), 7.0, device='cuda')
    ref16 = torch.full((n_seq, stride, ld2), 3.0, dtype=torch.bfloat16, device='cuda')
    cabi.gemm(A, W1, 4 * C, C, 1, M, bias=b1, act=cabi.ACT_GELU, out_act=H)
    cabi.gemm(H, W2, C, 4 * C, n_seq, rows, bias=b2, colscale=ls, resid=resid, rowmask=mask, m_seq_stride=stride,
              out_f32=ref32, out_act=ref16, ldo2=ld2, o2_seq_stride=stride)
    # fused
    out32 = torch.full((M, C), 7.0, device='cuda')
    out16 = torch.full((n_seq, stride, ld2), 3.0, dtype=torch.bfloat16, device='cuda')
    cabi.ffn(A, W1, b1, W2, b2, C, n_seq, rows, colscale=ls, resid=resid, rowmask=mask, m_seq_stride=stride,
             out_f32=out32, out_act=out16, ldo2=ld2, o2_seq_stride=stride)
    torch.cuda.synchronize()
    d = (out32 - ref32).abs().max().item()
    scale = ref32.abs().max().item()
    assert torch.equal(out32, ref32), (what, d, scale)          # same operations in the same order: bit-identical
    assert torch.equal(out16, ref16), what
    print(f'[ffn {what}] fused vs gemm pair: max |delta| {d:.3e} (bit-equal fp32: {torch.equal(out32, ref32)}, bf16: {torch.equal(out16, ref16)})')
    # untouched padding of the strided outputs
    assert bool((out16[:, rows:, :] == 3.0).all()) and bool((out16[:, :, C:] == 3.0).all())
    # (b) torch fp32 on the same operands (hidden tensor rounded to bf16 like both CUDA paths; erf-form GELU vs tanh form:
    # <= 5e-4 absolute on the hidden values)
    h = torch.nn.functional.gelu(A.float() @ W1.float().t() + b1, approximate='tanh').to(torch.bfloat16).float()
    want = ((h @ W2.float().t() + b2) * ls + resid) * mask[:, :rows].reshape(M, 1).float()
    assert (out32 - want).abs().max().item() <= 2e-3 * want.abs().max().item(), what


def test_fused_ffn_in_place_residual_and_no_optional_inputs():
    """out_f32 may alias resid (the engine updates the residual stream in place); bias / colscale / resid / mask optional."""
    from decaf_b200 import _cabi as cabi
    C, M = 256, 1000
    A, W1, b1, W2, b2, ls, resid, mask, stride = _mk(C, 1, M, 0, C, seed=3)
    x = resid.clone()
    cabi.ffn(A, W1, b1, W2, b2, C, 1, M, colscale=ls, resid=x, out_f32=x)
    H = torch.empty(M, 4 * C, dtype=torch.bfloat16, device='cuda')
    ref = torch.empty(M, C, device='cuda')
    cabi.gemm(A, W1, 4 * C, C, 1, M, bias=b1, act=cabi.ACT_GELU, out_act=H)
    cabi.gemm(H, W2, C, 4 * C, 1, M, bias=b2, colscale=ls, resid=resid, out_f32=ref)
    torch.cuda.synchronize()
    assert torch.equal(x, ref)
    out = torch.empty(M, C, device='cuda')
    cabi.ffn(A, W1, None, W2, None, C, 1, M, out_f32=out)
    want = torch.nn.functional.gelu(A.float() @ W1.float().t(), approximate='tanh').to(torch.bfloat16).float() @ W2.float().t()
    torch.cuda.synchronize()
    assert (out - want).abs().max().item() <= 2e-3 * want.abs().max().item()


def test_engine_fused_ffn_equals_unfused():
    """Whole grounder with and without the fused FFN: final logits / offsets equal (same arithmetic per element)."""
    from decaf_b200 import synth
    from decaf_b200.worker_v2 import Evaluator, create_model
    opt = synth.tiny_opt(embd_dim=128, n_levels=5, win=9, max_seq_len=256, sn=12, vid_in_dim=64, text_dim=64)
    shapes = {k: tuple(v.shape) for k, v in create_model(opt.clone()).state_dict().items()}
    sd = synth.fill_state_dict(shapes, 21)
    data = synth.synth_video(opt, 230, 4, seed=21, tag='ffn', n_events=1)
    outs = []
    for fused in (True, False):
        ev = Evaluator(opt.clone(), dataset=[data], state_dict=sd, act_dtype=torch.bfloat16, use_graphs=False)
        eng = ev.model.engine()
        assert eng.fused_ffn
        eng.fused_ffn, eng.ffn_min_rows = fused, 0
        ev.predict_video(data)
        p = eng.plan(4, ev.padded_len(230))
        outs.append((p.logits2.clone(), p.offsets.clone()))
    assert torch.equal(outs[0][0], outs[1][0])
    assert (outs[0][1] - outs[1][1]).abs().max().item() <= 1e-5 * outs[1][1].abs().max().item()

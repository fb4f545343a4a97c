"""GPU: the tcgen05/TMEM/TMA GEMM (impl=2) against the SIMT fp32-FMA GEMM (impl=1) on identical bf16
operands — the two differ only in fp32 accumulation order — and against a torch fp32 statement."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rand(*s, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*s, generator=g).cuda()


def _rel(a, b):
    return (a.double() - b.double()).abs().max().item() / max(b.double().abs().max().item(), 1e-12)


CASES = [
    # n_seq, T, K, N, taps, dil, lda_extra
    (1, 300, 64, 64, 1, 1, 0),          # flat, single k-block, partial last M tile
    (2, 2304, 256, 256, 1, 1, 0),       # canonical 1x1
    (3, 200, 96, 96, 3, 1, 0),          # k=3 conv, partial K block (96 = 64 + 32), partial M tile per sequence
    (1, 1500, 288, 288, 3, 1, 0),       # head tower shape: N split into 2 x 144, K tail, one long padded sequence
    (2, 700, 256, 1024, 1, 1, 0),       # FFN fc: 4 N tiles
    (2, 700, 1024, 256, 1, 1, 0),       # FFN proj: 16 k-blocks (ring wraps 4x)
    (1, 640, 256, 256, 3, 1, 32),       # lda > K (reads the first K columns of a wider buffer)
    (4, 18, 128, 128, 1, 1, 0),         # tiny top FPN level, flat mode spans sequences
    (2, 130, 64, 16, 3, 2, 0),          # dilation 2, narrow N
    (1, 256, 80, 64, 1, 1, 0),          # K = 80 (not a multiple of 64)
]


@pytest.mark.parametrize('n_seq,T,K,N,taps,dil,lda_extra', CASES)
def test_gemm_tc_matches_simt(n_seq, T, K, N, taps, dil, lda_extra):
    from decaf_b200 import _cabi as cabi
    lda = K + lda_extra
    Abuf = _rand(n_seq, T, lda, seed=1).bfloat16()
    W = (_rand(N, taps, K, seed=2) / math.sqrt(K * taps)).bfloat16()
    bias, cs = _rand(N, seed=3), _rand(N, seed=4)
    resid = _rand(n_seq, T, N, seed=5)
    mask = (torch.rand(n_seq, T, generator=torch.Generator().manual_seed(6)) > 0.3).cuda().to(torch.uint8)
    outs = {}
    for impl in (1, 2):
        o32 = torch.full((n_seq, T, N), 3.0, device='cuda')
        oa = torch.full((n_seq, T, N), 3.0, device='cuda', dtype=torch.bfloat16)
        cabi.gemm(Abuf, W, N, K, n_seq, T, lda=lda, taps=taps, dil=dil, bias=bias, act=cabi.ACT_GELU, colscale=cs,
                  resid=resid, rowmask=mask, out_f32=o32, out_act=oa, impl=impl)
        outs[impl] = (o32, oa)
    torch.cuda.synchronize()
    assert _rel(outs[2][0], outs[1][0]) < 2e-5
    assert _rel(outs[2][1].float(), outs[1][1].float()) < 1e-2
    x = Abuf[..., :K].float().permute(0, 2, 1)
    w = W.float().permute(0, 2, 1).contiguous()
    ref = F.conv1d(x, w, bias, padding=(taps // 2) * dil, dilation=dil)
    ref = ((F.gelu(ref) * cs[None, :, None] + resid.permute(0, 2, 1)) * mask[:, None, :].float()).permute(0, 2, 1)
    assert _rel(outs[2][0], ref) < 2e-5


def test_gemm_tc_grouped_and_remapped_output():
    from decaf_b200 import _cabi as cabi
    rows, K, N = 500, 128, 128
    A = _rand(3, rows, K, seed=1).bfloat16()
    W = (_rand(3, N, 1, K, seed=2) / 11).bfloat16()
    b = _rand(3, N, seed=3)
    res = {}
    for impl in (1, 2):
        out = torch.zeros(3, rows, N, device='cuda', dtype=torch.bfloat16)
        cabi.gemm(A, W, N, K, 1, rows, bias=b, out_act=out, n_group=3, g_stride_a=rows * K, g_stride_w=N * K,
                  g_stride_bias=N, g_stride_out_act=rows * N, impl=impl)
        res[impl] = out
    assert _rel(res[2].float(), res[1].float()) < 1e-2
    ref = torch.einsum('grk,gnk->grn', A.float(), W[:, :, 0].float()) + b[:, None, :]
    assert _rel(res[2].float(), ref) < 1e-2
    # second output with its own row mapping / pitch (FPN level written into the padded head buffer)
    n_seq, T, Pp, C2 = 4, 72, 100, 160
    A2 = _rand(n_seq, T, K, seed=7).bfloat16()
    mask = torch.ones(n_seq, Pp, dtype=torch.uint8, device='cuda')
    res = {}
    for impl in (1, 2):
        x = torch.zeros(n_seq, T, N, device='cuda')
        cat = torch.zeros(n_seq * Pp, C2, device='cuda', dtype=torch.bfloat16)
        cabi.gemm(A2, W[0], N, K, n_seq, T, rowmask=mask.view(-1)[5:], m_seq_stride=Pp, out_f32=x,
                  out_act=cat[5:], ldo2=C2, o2_seq_stride=Pp, impl=impl)
        res[impl] = (x, cat)
    assert _rel(res[2][0], res[1][0]) < 2e-5
    assert torch.equal(res[2][1] != 0, res[1][1] != 0)
    assert _rel(res[2][1].float(), res[1][1].float()) < 1e-2


LN_CASES = [
    # n_seq, T, K, N, taps, affine, relu, use_pe, f32_out
    (2, 300, 64, 64, 1, True, True, False, False),
    (1, 1500, 256, 256, 3, True, True, True, True),      # embed conv: LN -> ReLU -> +PE -> mask, fp32 residual stream out
    (1, 1100, 288, 288, 3, True, True, False, False),    # head tower C2 = 288: two N = 144 MMAs, one 288-column accumulator
    (3, 700, 288, 288, 3, True, True, False, False),     # the same over several padded sequences (register-resident LN epilogue)
    (2, 260, 96, 288, 1, False, True, False, False),     # ... flat tiling, no affine
    (3, 130, 128, 160, 1, False, False, False, True),    # affine=False, no activation
    (1, 700, 96, 32, 3, True, True, False, False),       # one 32-column chunk: second column half idle
    (2, 260, 320, 512, 1, True, False, False, False),    # N = 512: the whole TMEM
]


@pytest.mark.parametrize('n_seq,T,K,N,taps,affine,relu,use_pe,f32_out', LN_CASES)
def test_gemm_tc_fused_layernorm(n_seq, T, K, N, taps, affine, relu, use_pe, f32_out):
    from decaf_b200 import _cabi as cabi
    A = _rand(n_seq, T, K, seed=11).bfloat16()
    W = (_rand(N, taps, K, seed=12) / math.sqrt(K * taps)).bfloat16()
    bias = _rand(N, seed=13)
    lw, lb = (_rand(N, seed=14), _rand(N, seed=15)) if affine else (None, None)
    pe = _rand(T, N, seed=16) if use_pe else None
    mask = (torch.rand(n_seq, T, generator=torch.Generator().manual_seed(17)) > 0.3).cuda().to(torch.uint8)
    oa = torch.full((n_seq, T, N), 3.0, device='cuda', dtype=torch.bfloat16)
    o32 = torch.full((n_seq, T, N), 3.0, device='cuda') if f32_out else None
    cabi.gemm(A, W, N, K, n_seq, T, taps=taps, bias=bias, act=cabi.ACT_RELU if relu else cabi.ACT_NONE, rowmask=mask,
              out_f32=o32, out_act=oa, ln=True, ln_w=lw, ln_b=lb, pe=pe, impl=2)
    torch.cuda.synchronize()
    x = A.float().permute(0, 2, 1)
    w = W.float().permute(0, 2, 1).contiguous()
    y = F.conv1d(x, w, bias, padding=taps // 2).permute(0, 2, 1)            # (n_seq, T, N)
    mu = y.mean(-1, keepdim=True)
    r = y - mu
    sig = (r * r).mean(-1, keepdim=True)
    y = r / torch.sqrt(sig + 1e-5)
    if affine:
        y = y * lw + lb
    if relu:
        y = y.relu()
    if use_pe:
        y = y + pe[None]
    y = y * mask[..., None].float()
    if f32_out:
        assert _rel(o32, y) < 2e-5
    assert _rel(oa.float(), y) < 1e-2


def test_gemm_ln_needs_tensor_core_path():
    from decaf_b200 import _cabi as cabi
    A = _rand(1, 128, 64, seed=1)
    W = _rand(64, 1, 64, seed=2)
    out = torch.zeros(1, 128, 64, device='cuda')
    with pytest.raises(RuntimeError, match='tcgen05'):
        cabi.gemm(A, W, 64, 64, 1, 128, out_f32=out, ln=True)


def test_gemm_tc_many_tiles_persistent_ring():
    """More tiles than SMs and more k-iterations than pipeline stages: every CTA wraps both mbarrier rings
    and alternates the two TMEM accumulator stages many times."""
    from decaf_b200 import _cabi as cabi
    n_seq, T, K, N = 16, 2304, 256, 1024
    A = _rand(n_seq, T, K, seed=21).bfloat16()
    W = (_rand(N, 1, K, seed=22) / 16).bfloat16()
    b = _rand(N, seed=23)
    out = torch.zeros(n_seq, T, N, device='cuda', dtype=torch.bfloat16)
    cabi.gemm(A, W, N, K, 1, n_seq * T, bias=b, act=cabi.ACT_GELU, out_act=out, impl=2)
    ref = F.gelu(A.float() @ W[:, 0].float().t() + b)
    assert _rel(out.float(), ref) < 1e-2


@pytest.mark.parametrize('width', [2, 48, 74])
def test_launch_width_does_not_change_results(width):
    """decaf_set_gemm_sms caps the persistent grids of the tensor-core GEMM and of the fused FFN (lanes space-share the
    device); a tile's arithmetic does not depend on which CTA computes it: bit-identical outputs at any width."""
    from decaf_b200 import _cabi as cabi
    n_seq, T, K, N = 3, 1000, 256, 256
    A = _rand(n_seq, T, K, seed=31).bfloat16()
    W = (_rand(N, 3, K, seed=32) / 16).bfloat16()
    b, lw, lb = _rand(N, seed=33), _rand(N, seed=34), _rand(N, seed=35)
    W1 = (_rand(4 * K, K, seed=36) / 16).bfloat16()
    W2 = (_rand(K, 4 * K, seed=37) / 32).bfloat16()
    b1, b2, ls = _rand(4 * K, seed=38), _rand(K, seed=39), _rand(K, seed=40)
    resid = _rand(n_seq * T, K, seed=41)

    def run():
        conv = torch.zeros(n_seq, T, N, device='cuda', dtype=torch.bfloat16)
        cabi.gemm(A, W, N, K, n_seq, T, taps=3, bias=b, act=cabi.ACT_RELU, ln=True, ln_w=lw, ln_b=lb, out_act=conv, impl=2)
        ffn = torch.zeros(n_seq * T, K, device='cuda')
        cabi.ffn(A.view(-1, K), W1, b1, W2, b2, K, 1, n_seq * T, colscale=ls, resid=resid, out_f32=ffn)
        return conv, ffn
    prev = cabi.set_gemm_sms(0)
    try:
        want = run()
        assert cabi.set_gemm_sms(width) == 0
        got = run()
        assert cabi.set_gemm_sms(0) == width
    finally:
        cabi.set_gemm_sms(prev)
    assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])

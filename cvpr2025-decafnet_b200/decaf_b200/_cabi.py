"""ctypes binding of libdecaf_b200.so (C ABI declared in include/decaf_b200.h).

The library is REQUIRED: importing this module raises if the shared object is missing, and
every wrapper raises RuntimeError with the library's message on a non-zero status.  There is
no CPU or PyTorch fallback anywhere in the product path.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libdecaf_b200.so')

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2
MAX_LEVELS = 16

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f'{LIB_PATH} not found: build it with `make -C cvpr2025-decafnet_b200/csrc` or '
        '`python -c "import __graft_entry__ as g; g.build()"` (no fallback path exists)')

lib = C.CDLL(LIB_PATH)

vp, i32, i64, f32, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double


class Levels(C.Structure):
    _fields_ = [('n_levels', i32), ('Pp', i32), ('off', i32 * MAX_LEVELS), ('len', i32 * MAX_LEVELS)]


class GemmParams(C.Structure):
    _fields_ = [
        ('A', vp), ('dtype', i32), ('lda', i64), ('a_seq_stride', i64),
        ('n_seq', i32), ('rows_per_seq', i32),
        ('W', vp),
        ('N', i32), ('K', i32), ('taps', i32), ('dil', i32),
        ('bias', vp),
        ('act', i32),
        ('colscale', vp),
        ('resid', vp), ('ldr', i64), ('r_seq_stride', i64),
        ('rowmask', vp), ('m_seq_stride', i64),
        ('out_f32', vp), ('ldo', i64), ('o_seq_stride', i64),
        ('out_act', vp), ('ldo2', i64), ('o2_seq_stride', i64),
        ('n_group', i32),
        ('g_stride_a', i64), ('g_stride_w', i64), ('g_stride_bias', i64),
        ('g_stride_out_f32', i64), ('g_stride_out_act', i64),
        ('impl', i32),
        ('ln', i32), ('ln_w', vp), ('ln_b', vp), ('ln_eps', f32),
        ('pe', vp),
    ]


class FfnParams(C.Structure):
    _fields_ = [
        ('A', vp), ('dtype', i32), ('lda', i64), ('a_seq_stride', i64),
        ('n_seq', i32), ('rows_per_seq', i32), ('C', i32),
        ('W1', vp), ('b1', vp),
        ('W2', vp), ('b2', vp),
        ('colscale', vp),
        ('resid', vp), ('ldr', i64), ('r_seq_stride', i64),
        ('rowmask', vp), ('m_seq_stride', i64),
        ('out_f32', vp), ('ldo', i64), ('o_seq_stride', i64),
        ('out_act', vp), ('ldo2', i64), ('o2_seq_stride', i64),
    ]


class LayerNormParams(C.Structure):
    _fields_ = [
        ('x', vp), ('ldx', i64), ('x_seq_stride', i64),
        ('n_seq', i32), ('rows_per_seq', i32), ('C', i32),
        ('w', vp), ('b', vp),
        ('eps', f32), ('relu', i32),
        ('pe', vp),
        ('rowmask', vp), ('m_seq_stride', i64),
        ('out_f32', vp), ('ldo', i64), ('o_seq_stride', i64),
        ('out_act', vp), ('dtype', i32), ('ldo2', i64), ('o2_seq_stride', i64),
    ]


class PreAttnParams(C.Structure):
    _fields_ = [
        ('x', vp), ('n_seq', i32), ('T_in', i32), ('C', i32), ('stride', i32),
        ('mask_in', vp), ('mi_seq_stride', i64),
        ('w_pre', vp), ('b_pre', vp),
        ('n_branch', i32),
        ('wd', vp),
        ('w_br', vp), ('b_br', vp),
        ('eps', f32),
        ('out_act', vp), ('dtype', i32), ('out_branch_stride', i64),
        ('skip_out', vp),
        ('mask_out', vp), ('mo_seq_stride', i64),
    ]


class AdaLNParams(C.Structure):
    _fields_ = [
        ('q', vp), ('rows', i32), ('C', i32),
        ('ss', vp), ('ss_dtype', i32),
        ('rowmask', vp),
        ('w_ffn', vp), ('b_ffn', vp), ('eps', f32),
        ('out_q', vp), ('out_act', vp), ('dtype', i32),
    ]


class DecodeWindow(C.Structure):
    _fields_ = [('t0', i32), ('own_lo', i32), ('own_hi', i32), ('T_global', i32)]


class NmsParams(C.Structure):
    _fields_ = [
        ('mode', i32), ('iou_thresh', f32), ('sigma', f32), ('min_score', f32),
        ('max_num_segs', i32), ('voting_thresh', f32),
        ('to_seconds', i32), ('vid_stride', f32), ('clip_stride', f32), ('half_clip_size', f32),
        ('fps', f32), ('duration', f32),
        ('video_meta', vp),
    ]


class TextEncoderParams(C.Structure):
    _fields_ = [
        ('tokens', vp), ('lens', vp),
        ('n_query', i32), ('Lmax', i32), ('Ctok', i32), ('Ct', i32), ('n_heads', i32), ('n_layers', i32),
        ('n_fusion', i32), ('C', i32),
        ('wblob', vp), ('pblob', vp), ('pe', vp), ('pe_rows', i32),
        ('eps', f32),
        ('text_out', vp), ('kv_out', vp), ('kv_len_out', vp),
    ]


def _sig(name, restype, *argtypes):
    fn = getattr(lib, name)
    fn.restype = restype
    fn.argtypes = list(argtypes)
    return fn


_last_error = _sig('decaf_last_error', C.c_char_p)
version = _sig('decaf_version', i32)
device_is_sm100 = _sig('decaf_device_is_sm100', i32)
set_gemm_sms = _sig('decaf_set_gemm_sms', i32, i32)
_gemm = _sig('decaf_gemm', i32, C.POINTER(GemmParams), vp)
debug_gemm_trace = _sig('decaf_debug_gemm_trace', i32, vp)
_upload_2d = _sig('decaf_upload_2d', i32, vp, i64, vp, i64, i64, i64, vp)
_ffn = _sig('decaf_ffn', i32, C.POINTER(FfnParams), vp)
ffn_supported = _sig('decaf_ffn_supported', i32, i32, i32)
debug_ffn_trace = _sig('decaf_debug_ffn_trace', i32, vp)
_layernorm = _sig('decaf_layernorm', i32, C.POINTER(LayerNormParams), vp)
_preattn = _sig('decaf_preattn', i32, C.POINTER(PreAttnParams), vp)
_adaln = _sig('decaf_adaln', i32, C.POINTER(AdaLNParams), vp)
_local_attn = _sig('decaf_local_attn_phase', i32, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, i64, i32, vp)
_xattn = _sig('decaf_xattn', i32, vp, i32, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp)
xattn_packed_elems = _sig('decaf_xattn_packed_elems', i64, i32, i32, i32)
xattn_packed_supported = _sig('decaf_xattn_packed_supported', i32, i32, i32, i32)
_xattn_pack_kv = _sig('decaf_xattn_pack_kv', i32, vp, vp, vp, vp, i32, i32, i32, vp)
_xattn_packed = _sig('decaf_xattn_packed', i32, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp)
_split_bf16x3 = _sig('decaf_split_bf16x3', i32, vp, i64, i32, i64, vp, i32, vp)
_text_init = _sig('decaf_text_init', i32, vp, i32, i32, i32, vp, vp, vp, vp)
_saliency = _sig('decaf_saliency', i32, vp, vp, vp, i32, i32, i32, i32, vp)
_select = _sig('decaf_select', i32, vp, vp, vp, vp, vp, i32, i32, i32, i32, f64, i32, vp, vp)
_merge = _sig('decaf_merge', i32, vp, i32, vp, i32, vp, i32, vp, vp, vp, i32, i64, i32, i32, vp)
_map_combine = _sig('decaf_map_combine', i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp)
_scatter_clips = _sig('decaf_scatter_clips', i32, vp, i64, vp, i32, i32, vp, i32, vp)
_cast_bf16 = _sig('decaf_cast_bf16', i32, vp, vp, i64, vp)
_build_masks = _sig('decaf_build_masks', i32, vp, i64, vp, C.POINTER(Levels), i32, vp)
_head_out = _sig('decaf_head_out', i32, vp, i32, i64, i32, i32, vp, vp, i32, i32, vp, C.POINTER(Levels), vp, vp)
_tcn_in = _sig('decaf_tcn_in', i32, vp, vp, C.POINTER(Levels), vp, vp, i32, vp, i32, vp)
_tcn_layer = _sig('decaf_tcn_layer', i32, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp, f32, i32, i32, i32, i32, vp)
_tcn_out = _sig('decaf_tcn_out', i32, vp, vp, i64, vp, vp, i32, vp, i32, i64, i32, C.POINTER(Levels), i32, vp)
_refine_pool = _sig('decaf_refine_pool', i32, vp, i32, i64, i32, i32, vp, C.POINTER(Levels), i32, i32, vp)
tcn_fused_supported = _sig('decaf_tcn_fused_supported', i32, i32, i32)
_tcn_fused = _sig('decaf_tcn_fused', i32, vp, vp, C.POINTER(Levels), vp, vp, vp, vp, i32, vp, vp, i32, f32, vp, i64, i32, i32, vp, vp)
refine_pyramid_supported = _sig('decaf_refine_pyramid_supported', i32, i32)
_refine_pyramid = _sig('decaf_refine_pyramid', i32, vp, i32, i64, i32, i32, vp, C.POINTER(Levels), i32, vp)
_text_prep = _sig('decaf_text_prep', i32, vp, i32, i32, i32, vp, vp, i32, vp, vp)
text_encoder_supported = _sig('decaf_text_encoder_supported', i32, i32, i32, i32, i32, i32, i32, i32)
debug_text_max_clusters = _sig('decaf_debug_text_max_clusters', i32)
debug_text_trace = _sig('decaf_debug_text_trace', i32, vp)
_text_encoder = _sig('decaf_text_encoder', i32, C.POINTER(TextEncoderParams), vp)
text_encoder_wblob_floats = _sig('decaf_text_encoder_wblob_floats', i64, i32, i32, i32, i32, i32)
text_encoder_pblob_floats = _sig('decaf_text_encoder_pblob_floats', i64, i32, i32, i32, i32)
_decode = _sig('decaf_decode', i32, vp, vp, vp, C.POINTER(Levels), i32, i32, f32, i32, f32, vp, vp, vp, vp, vp)
_eval_loss = _sig('decaf_eval_loss', i32, vp, vp, vp, C.POINTER(Levels), i32, vp, vp, i32, f32, f32, f32, vp, vp)
_decode_window = _sig('decaf_decode_window', i32, vp, vp, vp, C.POINTER(Levels), i32, i32, f32, i32, f32, C.POINTER(DecodeWindow), vp, vp, vp, vp, vp)
_merge_candidates = _sig('decaf_merge_candidates', i32, vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp)
nms_workspace_bytes = _sig('decaf_nms_workspace_bytes', i64, i32, i32)
_softnms = _sig('decaf_softnms_1d', i32, vp, vp, vp, i32, i32, vp, vp, vp, f32, f32, f32, i32, i32, vp, vp)
_nms = _sig('decaf_nms_1d', i32, vp, vp, vp, i32, i32, vp, vp, f32, f32, i32, vp, vp)
_batched_nms = _sig('decaf_batched_nms', i32, vp, vp, vp, i32, i32, C.POINTER(NmsParams), vp, vp, vp, vp, vp)

EXPORTED = [
    'decaf_last_error', 'decaf_version', 'decaf_device_is_sm100', 'decaf_set_gemm_sms', 'decaf_gemm', 'decaf_debug_gemm_trace', 'decaf_layernorm',
    'decaf_preattn', 'decaf_adaln', 'decaf_local_attn', 'decaf_xattn', 'decaf_saliency', 'decaf_select',
    'decaf_merge', 'decaf_map_combine', 'decaf_scatter_clips', 'decaf_cast_bf16', 'decaf_build_masks', 'decaf_head_out', 'decaf_tcn_in', 'decaf_tcn_layer',
    'decaf_tcn_out', 'decaf_refine_pool', 'decaf_tcn_fused', 'decaf_tcn_fused_supported', 'decaf_refine_pyramid',
    'decaf_refine_pyramid_supported', 'decaf_text_prep', 'decaf_decode', 'decaf_nms_workspace_bytes',
    'decaf_softnms_1d', 'decaf_nms_1d', 'decaf_batched_nms', 'decaf_text_encoder_supported', 'decaf_text_encoder', 'decaf_debug_text_trace', 'decaf_debug_text_max_clusters',
    'decaf_text_encoder_wblob_floats', 'decaf_text_encoder_pblob_floats', 'decaf_decode_window', 'decaf_merge_candidates',
    'decaf_split_bf16x3', 'decaf_text_init', 'decaf_xattn_packed_elems', 'decaf_xattn_packed_supported', 'decaf_xattn_pack_kv', 'decaf_xattn_packed',
    'decaf_ffn', 'decaf_ffn_supported', 'decaf_debug_ffn_trace', 'decaf_local_attn_phase', 'decaf_eval_loss', 'decaf_upload_2d',
]


# launch accounting / optional per-GEMM CUDA-event timing (bench.py: roofline of the dominant kernel)
counters = {'launches': 0}
gemm_prof = None            # set to a list to record (flops, bytes, start_event, end_event, tag) per GEMM launch
gemm_record = None          # set to a list to record (GemmParams copy, flops, bytes, tensors kept alive, tag) per GEMM launch
gemm_tag = None             # label attached to recorded launches (the engine marks the text encoder's GEMMs 'text')


def check(status, what, n_launch=1):
    if status != 0:
        raise RuntimeError(f'{what}: {_last_error().decode()}')
    counters['launches'] += n_launch


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return None if t is None else t.data_ptr()


def dtype_code(t_or_dtype):
    d = t_or_dtype.dtype if isinstance(t_or_dtype, torch.Tensor) else t_or_dtype
    if d == torch.float32:
        return F32
    if d == torch.bfloat16:
        return BF16
    raise TypeError(f'unsupported activation dtype {d}')


def make_levels(lens):
    lv = Levels()
    lv.n_levels = len(lens)
    off = 1
    for l, n in enumerate(lens):
        lv.off[l] = off
        lv.len[l] = n
        off += n + 1
    lv.Pp = off
    return lv


# ----------------------------------------------------------------------------- wrappers
def gemm(A, W, N, K, n_seq, rows_per_seq, *, lda=None, a_seq_stride=0, taps=1, dil=1, bias=None,
         act=ACT_NONE, colscale=None, resid=None, ldr=0, r_seq_stride=0, rowmask=None, m_seq_stride=0,
         out_f32=None, ldo=0, o_seq_stride=0, out_act=None, ldo2=0, o2_seq_stride=0, n_group=1,
         g_stride_a=0, g_stride_w=0, g_stride_bias=0, g_stride_out_f32=0, g_stride_out_act=0, impl=0,
         ln=False, ln_w=None, ln_b=None, ln_eps=1e-5, pe=None):
    p = GemmParams()
    p.A, p.dtype, p.lda, p.a_seq_stride = ptr(A), dtype_code(A), (lda or K), a_seq_stride
    p.n_seq, p.rows_per_seq = n_seq, rows_per_seq
    assert W.dtype == A.dtype, (W.dtype, A.dtype)
    p.W, p.N, p.K, p.taps, p.dil = ptr(W), N, K, taps, dil
    p.bias, p.act, p.colscale = ptr(bias), act, ptr(colscale)
    p.resid, p.ldr, p.r_seq_stride = ptr(resid), (ldr or N), r_seq_stride
    p.rowmask, p.m_seq_stride = ptr(rowmask), m_seq_stride
    p.out_f32, p.ldo, p.o_seq_stride = ptr(out_f32), (ldo or N), o_seq_stride
    p.out_act, p.ldo2, p.o2_seq_stride = ptr(out_act), (ldo2 or N), o2_seq_stride
    if out_act is not None:
        assert out_act.dtype == A.dtype
    p.n_group = n_group
    p.g_stride_a, p.g_stride_w, p.g_stride_bias = g_stride_a, g_stride_w, g_stride_bias
    p.g_stride_out_f32, p.g_stride_out_act = g_stride_out_f32, g_stride_out_act
    p.impl = impl
    p.ln, p.ln_w, p.ln_b, p.ln_eps, p.pe = int(ln), ptr(ln_w), ptr(ln_b), ln_eps, ptr(pe)
    if gemm_record is not None:
        M = n_seq * rows_per_seq
        eb = A.element_size()
        flops = 2.0 * M * N * K * taps * n_group
        nbytes = n_group * (M * K * eb + N * K * taps * eb + M * N * ((4 if out_f32 is not None else 0) +
                            (eb if out_act is not None else 0) + (4 if resid is not None else 0)))
        q = GemmParams()
        C.memmove(C.byref(q), C.byref(p), C.sizeof(GemmParams))
        gemm_record.append((q, flops, nbytes, (A, W, bias, colscale, resid, rowmask, out_f32, out_act, ln_w, ln_b, pe), gemm_tag))
    if gemm_prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(_gemm(C.byref(p), stream_ptr()), 'decaf_gemm')
        e1.record()
        M = n_seq * rows_per_seq
        eb = A.element_size()
        flops = 2.0 * M * N * K * taps * n_group
        nbytes = n_group * (M * K * eb + N * K * taps * eb + M * N * ((4 if out_f32 is not None else 0) +
                            (eb if out_act is not None else 0) + (4 if resid is not None else 0)))
        gemm_prof.append((flops, nbytes, e0, e1, (M, N, K, taps, n_group)))
        return
    check(_gemm(C.byref(p), stream_ptr()), 'decaf_gemm')


def gemm_replay(q):
    """Re-launch a recorded GEMM / fused FFN (bench.py: time the tensor-core launches of one step in isolation)."""
    if isinstance(q, FfnParams):
        check(_ffn(C.byref(q), stream_ptr()), 'decaf_ffn')
    else:
        check(_gemm(C.byref(q), stream_ptr()), 'decaf_gemm')


def upload_2d(dst, src):
    """dst (device, 2-D, unit inner stride) <- src (host, same shape, unit inner stride; pinned for an asynchronous copy)."""
    assert dst.dim() == 2 and src.shape == dst.shape and dst.stride(1) == 1 and src.stride(1) == 1 and dst.dtype == src.dtype
    es = dst.element_size()
    check(_upload_2d(dst.data_ptr(), dst.stride(0) * es, src.data_ptr(), src.stride(0) * es, dst.size(1) * es, dst.size(0),
                     stream_ptr()), 'decaf_upload_2d', n_launch=0)


def ffn(A, W1, b1, W2, b2, C_, n_seq, rows_per_seq, *, lda=None, colscale=None, resid=None, ldr=0, r_seq_stride=0,
        rowmask=None, m_seq_stride=0, out_f32=None, ldo=0, o_seq_stride=0, out_act=None, ldo2=0, o2_seq_stride=0):
    """decaf_ffn: out = ((GELU(A W1^T + b1) W2^T + b2) * colscale + resid) * rowmask in one tcgen05 launch."""
    p = FfnParams()
    p.A, p.dtype, p.lda, p.a_seq_stride = ptr(A), dtype_code(A), (lda or C_), 0
    p.n_seq, p.rows_per_seq, p.C = n_seq, rows_per_seq, C_
    assert W1.dtype == A.dtype and W2.dtype == A.dtype
    p.W1, p.b1, p.W2, p.b2, p.colscale = ptr(W1), ptr(b1), ptr(W2), ptr(b2), ptr(colscale)
    p.resid, p.ldr, p.r_seq_stride = ptr(resid), (ldr or C_), r_seq_stride
    p.rowmask, p.m_seq_stride = ptr(rowmask), m_seq_stride
    p.out_f32, p.ldo, p.o_seq_stride = ptr(out_f32), (ldo or C_), o_seq_stride
    p.out_act, p.ldo2, p.o2_seq_stride = ptr(out_act), (ldo2 or C_), o2_seq_stride
    M = n_seq * rows_per_seq
    flops = 2.0 * M * (4 * C_) * C_ * 2
    nbytes = M * C_ * 2 + 2 * 4 * C_ * C_ * 2 + M * C_ * ((4 if out_f32 is not None else 0) + (2 if out_act is not None else 0) +
                                                    (4 if resid is not None else 0))
    if gemm_record is not None:
        q = FfnParams()
        C.memmove(C.byref(q), C.byref(p), C.sizeof(FfnParams))
        gemm_record.append((q, flops, nbytes, (A, W1, b1, W2, b2, colscale, resid, rowmask, out_f32, out_act), gemm_tag))
    if gemm_prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(_ffn(C.byref(p), stream_ptr()), 'decaf_ffn')
        e1.record()
        gemm_prof.append((flops, nbytes, e0, e1, (M, C_, 4 * C_, 'ffn', 1)))
        return
    check(_ffn(C.byref(p), stream_ptr()), 'decaf_ffn')


def layernorm(x, C_, n_seq, rows_per_seq, *, ldx=None, x_seq_stride=0, w=None, b=None, eps=1e-5,
              relu=False, pe=None, rowmask=None, m_seq_stride=0, out_f32=None, ldo=0, o_seq_stride=0,
              out_act=None, ldo2=0, o2_seq_stride=0):
    p = LayerNormParams()
    p.x, p.ldx, p.x_seq_stride = ptr(x), (ldx or C_), x_seq_stride
    p.n_seq, p.rows_per_seq, p.C = n_seq, rows_per_seq, C_
    p.w, p.b, p.eps, p.relu, p.pe = ptr(w), ptr(b), eps, int(relu), ptr(pe)
    p.rowmask, p.m_seq_stride = ptr(rowmask), m_seq_stride
    p.out_f32, p.ldo, p.o_seq_stride = ptr(out_f32), (ldo or C_), o_seq_stride
    p.out_act, p.ldo2, p.o2_seq_stride = ptr(out_act), (ldo2 or C_), o2_seq_stride
    p.dtype = dtype_code(out_act) if out_act is not None else F32
    check(_layernorm(C.byref(p), stream_ptr()), 'decaf_layernorm')


def preattn(x, n_seq, T_in, C_, stride, mask_in, mi_seq_stride, w_pre, b_pre, n_branch, wd, w_br, b_br,
            out_act, out_branch_stride, skip_out=None, mask_out=None, mo_seq_stride=0, eps=1e-5):
    p = PreAttnParams()
    p.x, p.n_seq, p.T_in, p.C, p.stride = ptr(x), n_seq, T_in, C_, stride
    p.mask_in, p.mi_seq_stride = ptr(mask_in), mi_seq_stride
    p.w_pre, p.b_pre, p.n_branch, p.wd, p.w_br, p.b_br, p.eps = ptr(w_pre), ptr(b_pre), n_branch, ptr(wd), ptr(w_br), ptr(b_br), eps
    p.out_act, p.dtype, p.out_branch_stride = ptr(out_act), dtype_code(out_act), out_branch_stride
    p.skip_out, p.mask_out, p.mo_seq_stride = ptr(skip_out), ptr(mask_out), mo_seq_stride
    check(_preattn(C.byref(p), stream_ptr()), 'decaf_preattn')


def adaln(q, rows, C_, ss, rowmask, w_ffn, b_ffn, out_q, out_act, eps=1e-5):
    p = AdaLNParams()
    p.q, p.rows, p.C, p.ss, p.ss_dtype, p.rowmask = ptr(q), rows, C_, ptr(ss), dtype_code(ss), ptr(rowmask)
    p.w_ffn, p.b_ffn, p.eps, p.out_q, p.out_act, p.dtype = ptr(w_ffn), ptr(b_ffn), eps, ptr(out_q), ptr(out_act), dtype_code(out_act)
    check(_adaln(C.byref(p), stream_ptr()), 'decaf_adaln')


def local_attn(q, k, v, out, n_seq, T, C_, n_heads, window, mask, m_seq_stride, phase=0):
    check(_local_attn(ptr(q), ptr(k), ptr(v), ptr(out), dtype_code(q), n_seq, T, C_, n_heads, window,
                      ptr(mask), m_seq_stride, int(phase), stream_ptr()), 'decaf_local_attn')


def xattn(q, k, v, out, n_seq, Tq, Lk, C_, n_heads, kv_len):
    check(_xattn(ptr(q), dtype_code(q), ptr(k), ptr(v), ptr(out), dtype_code(out), n_seq, Tq, Lk, C_, n_heads,
                 ptr(kv_len), stream_ptr()), 'decaf_xattn')


def text_init(x, n_query, L1, C_, lens, kv_len, tmask):
    check(_text_init(ptr(x), n_query, L1, C_, ptr(lens), ptr(kv_len), ptr(tmask), stream_ptr()), 'decaf_text_init')


def split_bf16x3(src, rows, K, ld_src, dst, order, src_offset=0):
    """dst (rows, 3K) bf16 <- hi / lo bf16 parts of src (fp32, rows of K values, pitch ld_src, starting src_offset elements in)."""
    check(_split_bf16x3(src.data_ptr() + 4 * src_offset, rows, K, ld_src, ptr(dst), order, stream_ptr()), 'decaf_split_bf16x3')


def xattn_pack_kv(k, v, kv_len, packed, n_seq, Lk, C_):
    check(_xattn_pack_kv(ptr(k), ptr(v), ptr(kv_len), ptr(packed), n_seq, Lk, C_, stream_ptr()), 'decaf_xattn_pack_kv')


def xattn_packed(q, packed, out, n_seq, Tq, Lk, C_, n_heads, kv_len):
    check(_xattn_packed(ptr(q), ptr(packed), ptr(out), n_seq, Tq, Lk, C_, n_heads, ptr(kv_len), stream_ptr()), 'decaf_xattn_packed')


def saliency(shallow, text_cls, correl, Cs, T, n_query, norm):
    check(_saliency(ptr(shallow), ptr(text_cls), ptr(correl), Cs, T, n_query, int(norm), stream_ptr()), 'decaf_saliency')


def select(correl, vid_mask, sel, out_mask, pooled, max_blocks, T, n_query, sn, sratio, and_mask, vid_len_out=None):
    check(_select(ptr(correl), ptr(vid_mask), ptr(sel), ptr(out_mask), ptr(pooled), max_blocks, T, n_query, sn,
                  float(sratio), int(and_mask), ptr(vid_len_out), stream_ptr()), 'decaf_select')


def merge(vid, Ce, shallow, Cs, correl, scat, sel, out_mask, x0, ldx, T, n_query):
    check(_merge(ptr(vid), Ce, ptr(shallow), Cs, ptr(correl), int(scat), ptr(sel), ptr(out_mask), ptr(x0),
                 dtype_code(x0), ldx, T, n_query, stream_ptr()), 'decaf_merge')


def map_combine(E, S, bias, correl, wc, sel, mask, X, T, C_, n_query):
    check(_map_combine(ptr(E), ptr(S), ptr(bias), ptr(correl), ptr(wc), ptr(sel), ptr(mask), ptr(X), T, C_, n_query, stream_ptr()),
          'decaf_map_combine')


def scatter_clips(compact, ld, index, Ce, K, dense, T):
    check(_scatter_clips(ptr(compact), ld, ptr(index), Ce, K, ptr(dense), T, stream_ptr()), 'decaf_scatter_clips')


def cast_bf16(src, dst):
    check(_cast_bf16(ptr(src), ptr(dst), src.numel(), stream_ptr()), 'decaf_cast_bf16')


def build_masks(mask0, m0_seq_stride, hmask, lv, n_query):
    check(_build_masks(ptr(mask0), m0_seq_stride, ptr(hmask), C.byref(lv), n_query, stream_ptr()), 'decaf_build_masks')


def head_out(x, ldx, rows_total, C_, w, bias, n_out, mode, level_scale, lv, out):
    check(_head_out(ptr(x), dtype_code(x), ldx, rows_total, C_, ptr(w), ptr(bias), n_out, mode, ptr(level_scale),
                    C.byref(lv), ptr(out), stream_ptr()), 'decaf_head_out')


def tcn_in(logits1, hmask, lv, w_in, b_in, R, r0, n_query):
    check(_tcn_in(ptr(logits1), ptr(hmask), C.byref(lv), ptr(w_in), ptr(b_in), R, ptr(r0), n_query, stream_ptr()), 'decaf_tcn_in')


def tcn_layer(r_in, r_out, mask0, m_seq_stride, wd, bd, w1, b1, ln_w, ln_b, R, dil, n_query, T, eps=1e-5):
    check(_tcn_layer(ptr(r_in), ptr(r_out), ptr(mask0), m_seq_stride, ptr(wd), ptr(bd), ptr(w1), ptr(b1), ptr(ln_w),
                     ptr(ln_b), eps, R, dil, n_query, T, stream_ptr()), 'decaf_tcn_layer')


def tcn_out(r_in, mask0, m_seq_stride, w_out, b_out, R, cat, ldc, col0, lv, n_query):
    check(_tcn_out(ptr(r_in), ptr(mask0), m_seq_stride, ptr(w_out), ptr(b_out), R, ptr(cat), dtype_code(cat), ldc, col0,
                   C.byref(lv), n_query, stream_ptr()), 'decaf_tcn_out')


def tcn_fused(logits1, hmask, lv, w_in, b_in, wblob, vblob, n_layers, w_out, b_out, R, cat, ldc, col0, n_query, eps=1e-5, scratch=None):
    check(_tcn_fused(ptr(logits1), ptr(hmask), C.byref(lv), ptr(w_in), ptr(b_in), ptr(wblob), ptr(vblob), n_layers, ptr(w_out),
                     ptr(b_out), R, eps, ptr(cat), ldc, col0, n_query, ptr(scratch), stream_ptr()), 'decaf_tcn_fused',
          n_launch=2 if (scratch is not None and n_layers >= 6) else 1)


def refine_pyramid(cat, ldc, col0, R, hmask, lv, n_query):
    check(_refine_pyramid(ptr(cat), dtype_code(cat), ldc, col0, R, ptr(hmask), C.byref(lv), n_query, stream_ptr()),
          'decaf_refine_pyramid')


def refine_pool(cat, ldc, col0, R, hmask, lv, level, n_query):
    check(_refine_pool(ptr(cat), dtype_code(cat), ldc, col0, R, ptr(hmask), C.byref(lv), level, n_query, stream_ptr()),
          'decaf_refine_pool')


def text_prep(x, n_query, L1, C_, bkgd, pe, lens):
    """pe: the raw (max_seq_len, C) table or None; interpolated per query on the device when len > max_seq_len."""
    check(_text_prep(ptr(x), n_query, L1, C_, ptr(bkgd), ptr(pe), 0 if pe is None else int(pe.shape[0]), ptr(lens), stream_ptr()),
          'decaf_text_prep')


def text_encoder(prm):
    """prm: a filled TextEncoderParams (the engine keeps one per shape)."""
    check(_text_encoder(C.byref(prm), stream_ptr()), 'decaf_text_encoder')


def eval_loss(logits, offsets, hmask, lv, n_query, targets, reg_range, center_sampling, radius_mul, smoothing, alpha, out):
    check(_eval_loss(ptr(logits), ptr(offsets), ptr(hmask), C.byref(lv), n_query, ptr(targets), ptr(reg_range), int(center_sampling),
                     radius_mul, smoothing, alpha, ptr(out), stream_ptr()), 'decaf_eval_loss')


def decode(logits, offsets, hmask, lv, n_query, from_logits, pre_nms_thresh, topk, seg_len_thresh,
           cand_segs, cand_scores, cand_idx, cand_count):
    check(_decode(ptr(logits), ptr(offsets), ptr(hmask), C.byref(lv), n_query, int(from_logits), pre_nms_thresh, topk,
                  seg_len_thresh, ptr(cand_segs), ptr(cand_scores), ptr(cand_idx), ptr(cand_count), stream_ptr()),
          'decaf_decode')


def decode_window(logits, offsets, hmask, lv, n_query, from_logits, pre_nms_thresh, topk, seg_len_thresh, t0, own_lo, own_hi,
                  T_global, cand_segs, cand_scores, cand_idx, cand_count):
    win = DecodeWindow(t0, own_lo, own_hi, T_global)
    check(_decode_window(ptr(logits), ptr(offsets), ptr(hmask), C.byref(lv), n_query, int(from_logits), pre_nms_thresh, topk,
                         seg_len_thresh, C.byref(win), ptr(cand_segs), ptr(cand_scores), ptr(cand_idx), ptr(cand_count),
                         stream_ptr()), 'decaf_decode_window')


def merge_candidates(segs, scores, idx, count, n_src, n_query, topk, out_segs, out_scores, out_idx, out_count):
    check(_merge_candidates(ptr(segs), ptr(scores), ptr(idx), ptr(count), n_src, n_query, topk, ptr(out_segs), ptr(out_scores),
                            ptr(out_idx), ptr(out_count), stream_ptr()), 'decaf_merge_candidates')


def softnms_1d(segs, scores, n, n_query, cand_stride, dets, inds, n_out, iou_thresh, sigma, min_score, method,
               max_iters, workspace):
    check(_softnms(ptr(segs), ptr(scores), ptr(n), n_query, cand_stride, ptr(dets), ptr(inds), ptr(n_out), iou_thresh,
                   sigma, min_score, method, max_iters, ptr(workspace), stream_ptr()), 'decaf_softnms_1d')


def nms_1d(segs, scores, n, n_query, cand_stride, keep, n_out, iou_thresh, min_score, max_keep, workspace=None):
    check(_nms(ptr(segs), ptr(scores), ptr(n), n_query, cand_stride, ptr(keep), ptr(n_out), iou_thresh, min_score,
               max_keep, ptr(workspace), stream_ptr()), 'decaf_nms_1d')


def batched_nms(segs, scores, n, n_query, cand_stride, prm, out_segs, out_scores, out_count, workspace):
    check(_batched_nms(ptr(segs), ptr(scores), ptr(n), n_query, cand_stride, C.byref(prm), ptr(out_segs), ptr(out_scores),
                       ptr(out_count), ptr(workspace), stream_ptr()), 'decaf_batched_nms', 2)

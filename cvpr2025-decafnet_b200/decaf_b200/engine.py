"""Host-side orchestration of the DeCaf-Grounder forward on one B200.

`GrounderEngine` owns the repacked weights, per-shape workspaces ("plans") and the launch
sequence; every arithmetic step is a kernel of libdecaf_b200.so (see include/decaf_b200.h).
PyTorch is used for device memory and streams only.

Data layout in HBM: activations are channels-last (n_query, T_l, C);
the residual stream is fp32, GEMM operands are `act_dtype` (bf16 by default, fp32 for the FP32
configuration); the heads work on a level-major, zero-row-padded flat point layout of
Pp = 1 + sum_l (T_l + 1) rows per query so one launch covers every FPN level.

Reference call graph reproduced: PtTransformerEarlyFusionIterative.
_drop_forward_eval (libs/modeling/model.py:480-565) with all queries of a video batched,
Evaluator._collect_segments / _generate_proposals (libs/worker_v2.py:1063-1187) and
libs/nms/nms.py:batched_nms.
"""
import math
import os

import torch
import torch.nn.functional as F

from . import _cabi as cabi

R_REFINE = 32          # TCN width, libs/modeling/model.py:424


def _sinusoid_pe(max_seq_len, embd_dim, t):
    """PE table (t, C) fp32, channels-last.  libs/modeling/blocks.py:134-142 and
    video_net.py:74-79,143-152 (constant-table preparation, done once per length on the host)."""
    tics = torch.arange(max_seq_len, dtype=torch.float)
    n_freqs = embd_dim // 2
    freqs = 10000 ** torch.linspace(0, 1, n_freqs + 1)[:n_freqs]
    x = tics[None, :] / freqs[:, None]
    pe = torch.cat((torch.sin(x), torch.cos(x))) / embd_dim ** 0.5
    if t > max_seq_len:
        pe = F.interpolate(pe[None], size=t, mode='linear', align_corners=True)[0]
    return pe[:, :t].t().contiguous()


class _Plan:
    """Workspaces for one (n_query, T) shape."""
    pass


class GrounderEngine:
    fp32_tc = False                # FP32 configuration on the tensor cores (set per instance in __init__ / _setup)
    def __init__(self, opt, state_dict, act_dtype=torch.bfloat16, device='cuda', gemm_impl=0, fused_text=True):
        if not torch.cuda.is_available():
            raise RuntimeError('GrounderEngine needs a CUDA device (no CPU fallback exists)')
        self.opt = opt
        self.dev = torch.device(device)
        self.act_dtype = act_dtype
        self.gemm_impl = gemm_impl
        # one-launch cluster kernel for the text encoder (when its shape limits allow); False composes the same
        # computation from GEMM / LayerNorm / attention launches
        self.fused_text = fused_text and os.environ.get('DECAF_FUSED_TEXT', '1') != '0'
        # bf16 configuration: the text encoder as ~47 small tensor-core launches (bf16 operands like the rest of the
        # configuration) on the forked stream.  Its latency equals the one-launch cluster kernel's, but it occupies 4-16 SMs
        # at a time instead of 128 for ~350 us, which the other lanes' GEMMs (one CTA per SM, nothing co-resides) get back.
        self.text_tc = act_dtype == torch.bfloat16 and gemm_impl != 1 and os.environ.get('DECAF_TEXT_TC', '1') != '0'
        m = opt['model']
        vn, tn, fu = m['vid_net'], m['text_net'], m['fusion']
        self.C = vn['embd_dim']
        self.Ct = tn['embd_dim']
        self.Ctok = tn['in_dim']
        self.Cin = vn['in_dim']
        self.n_heads = vn['n_heads']
        self.win = vn['mha_win_size']
        self.arch = tuple(vn['arch'])
        self.L = self.arch[2]
        self.C2 = self.C + R_REFINE
        self.msf, self.scat, self.sfonly, self.norm = m['msf'], m['scat'], m['sfonly'], m['norm']
        self.fusion_heads = fu['n_heads']
        self.sn, self.sratio = int(m['sn']), float(m['sratio'])
        assert vn['stride'] == 1, 'vid_net.stride > 1 is not on the released eval path'
        assert self.arch[0] >= 1, 'at least one embedding conv expected'
        assert self.win > 0, 'local attention window expected (mha_win_size > 0)'
        assert tn['name'] == 'transformer' and tn.get('use_bkgd_token', True)
        assert fu.get('xattn_mode', 'adaln') == 'adaln'
        assert self.L <= cabi.MAX_LEVELS
        # vid_map input channels (libs/modeling/model.py:410-416, 543-551)
        if not self.msf:
            self.Ce_eff, self.Cs_eff = self.Cin, 0
        elif self.sfonly:
            self.Ce_eff, self.Cs_eff = 0, self.Cin
        else:
            self.Ce_eff, self.Cs_eff = self.Cin, self.Cin
        cin_map = self.Ce_eff + self.Cs_eff + (1 if self.scat else 0)
        self.K0 = (cin_map + 63) // 64 * 64 if cin_map % 8 else cin_map
        self._pack(state_dict, cin_map)
        self._plans = {}
        self._pe_cache = {}
        self._text_ws = {}
        # conv -> LayerNorm -> ReLU chains run as ONE tcgen05 launch (LN in the epilogue) on the bf16 path; the
        # fp32 configuration (SIMT GEMM) keeps the separate row-wise LayerNorm kernel
        self.fuse_ln = (act_dtype == torch.bfloat16 and gemm_impl != 1 and bool(cabi.device_is_sm100()))
        # FP32 configuration: dense convolutions on the tensor cores through a bf16 hi/lo split of both operands
        # (decaf_split_bf16x3 + the tcgen05 GEMM over K' = 3K), instead of the fp32-FMA kernel
        self.fp32_tc = (act_dtype == torch.float32 and gemm_impl != 1 and bool(cabi.device_is_sm100()) and
                        os.environ.get('DECAF_FP32_TC', '1') != '0')
        self._w3 = {}                  # fp32 weight tensor (data_ptr, shape) -> its [hi | lo | hi] bf16 form
        self._a3 = {}                  # lane -> bf16 scratch for the split activations
        self.capture = None            # set to a dict to record intermediate tensors (tests/debugging)
        # workspace lane: videos in flight on different streams (Evaluator.predict_videos) use disjoint plans /
        # text workspaces; the packed weights, PE tables and weight blobs are shared
        self.lane = 0
        self.linear_map = os.environ.get('DECAF_LINEAR_MAP', '1') != '0'   # vid_map once per video + per-query combine
        self.fused_tcn = True          # False: the per-layer TCN / per-level pooling launches (tests compare both)
        # FFN fc -> GELU -> proj as ONE tcgen05 launch (decaf_ffn: the hidden tensor never reaches HBM); False keeps the two
        # GEMM launches with the bf16 hidden tensor in between (tests compare both: same arithmetic, same accumulation order)
        self.fused_ffn = (act_dtype == torch.bfloat16 and gemm_impl != 1 and os.environ.get('DECAF_FUSED_FFN', '1') != '0' and
                          bool(cabi.ffn_supported(self.C, cabi.BF16)))
        self.ffn_min_rows = int(os.environ.get('DECAF_FFN_MIN_ROWS', '0'))

    def _cap(self, name, t):
        if self.capture is not None:
            self.capture[name] = t.detach().float().clone()

    # ------------------------------------------------------------------ weights
    def _f32(self, t):
        return t.detach().to(self.dev, torch.float32).contiguous()

    def _act(self, t):
        return t.detach().to(self.dev, torch.float32).to(self.act_dtype).contiguous()

    def _conv_w(self, w, act=True):
        """(Cout, Cin, k) -> (Cout, k, Cin) K-major."""
        w = w.detach().permute(0, 2, 1).contiguous()
        return self._act(w) if act else self._f32(w)

    def _pack_text(self, W, sd, pre, tn):
        """TextTransformer weights (libs/modeling/text_net.py:102-156) under the state-dict prefix `pre` -> W['t.*'] (fp32)."""
        f32, conv = self._f32, self._conv_w
        W['t.embd.w'] = conv(sd[pre + 'embd_fc.conv.weight'], act=False)
        W['t.embd.b'] = f32(sd[pre + 'embd_fc.conv.bias'])
        W['t.bkgd'] = f32(sd[pre + 'bkgd_token'].reshape(-1))
        self.text_layers = tn.get('n_layers', 5)
        for i in range(self.text_layers):
            p = pre + f'transformer.{i}.'
            a = p + 'attn.attn.'
            W[f't{i}.qkv.w'] = torch.stack([conv(sd[a + f'{n}.weight'], act=False) for n in ('query', 'key', 'value')]).contiguous()
            W[f't{i}.qkv.b'] = torch.stack([f32(sd[a + f'{n}.bias']) for n in ('query', 'key', 'value')]).contiguous()
            W[f't{i}.proj.w'] = conv(sd[a + 'proj.weight'], act=False)
            W[f't{i}.proj.b'] = f32(sd[a + 'proj.bias'])
            for n in ('ln_attn', 'ln_ffn'):
                W[f't{i}.{n}.w'] = f32(sd[p + n + '.weight'].reshape(-1))
                W[f't{i}.{n}.b'] = f32(sd[p + n + '.bias'].reshape(-1))
            W[f't{i}.ls_attn'] = f32(sd[p + 'drop_path_attn.scale'].reshape(-1))
            W[f't{i}.ls_ffn'] = f32(sd[p + 'drop_path_ffn.scale'].reshape(-1))
            W[f't{i}.fc.w'] = conv(sd[p + 'ffn.fc.weight'], act=False)
            W[f't{i}.fc.b'] = f32(sd[p + 'ffn.fc.bias'])
            W[f't{i}.proj2.w'] = conv(sd[p + 'ffn.proj.weight'], act=False)
            W[f't{i}.proj2.b'] = f32(sd[p + 'ffn.proj.bias'])

    def _pack_fusion(self, W, sd, pre):
        """XAttNFusion weights (libs/modeling/fusion.py:21-54) under the state-dict prefix `pre` -> W['f*']."""
        f32, conv = self._f32, self._conv_w
        C = self.C
        for i in range(self.fusion_layers):
            p = pre + f'layers.{i}.'
            a = p + 'xattn.'
            W[f'f{i}.lnq.w'] = f32(sd[p + 'ln_xattn_q.weight'].reshape(-1))
            W[f'f{i}.lnq.b'] = f32(sd[p + 'ln_xattn_q.bias'].reshape(-1))
            W[f'f{i}.lnkv.w'] = f32(sd[p + 'ln_xattn_kv.weight'].reshape(-1))
            W[f'f{i}.lnkv.b'] = f32(sd[p + 'ln_xattn_kv.bias'].reshape(-1))
            W[f'f{i}.dw'] = f32(sd[a + 'q_conv.conv.weight'].reshape(1, C, 3))
            W[f'f{i}.qn.w'] = f32(sd[a + 'q_norm.weight'].reshape(1, C))
            W[f'f{i}.qn.b'] = f32(sd[a + 'q_norm.bias'].reshape(1, C))
            W[f'f{i}.q.w'] = conv(sd[a + 'xattn.query.weight'])
            W[f'f{i}.q.b'] = f32(sd[a + 'xattn.query.bias'])
            W[f'f{i}.kv.w'] = torch.stack([conv(sd[a + f'xattn.{n}.weight'], act=False) for n in ('key', 'value')]).contiguous()
            W[f'f{i}.kv.b'] = torch.stack([f32(sd[a + f'xattn.{n}.bias']) for n in ('key', 'value')]).contiguous()
            W[f'f{i}.proj.w'] = conv(sd[a + 'xattn.proj.weight'])
            W[f'f{i}.proj.b'] = f32(sd[a + 'xattn.proj.bias'])
            W[f'f{i}.lnf.w'] = f32(sd[p + 'ln_ffn.weight'].reshape(-1))
            W[f'f{i}.lnf.b'] = f32(sd[p + 'ln_ffn.bias'].reshape(-1))
            W[f'f{i}.fc.w'] = conv(sd[p + 'ffn.fc.weight'])
            W[f'f{i}.fc.b'] = f32(sd[p + 'ffn.fc.bias'])
            W[f'f{i}.proj2.w'] = conv(sd[p + 'ffn.proj.weight'])
            W[f'f{i}.proj2.b'] = f32(sd[p + 'ffn.proj.bias'])
            W[f'f{i}.ls_ffn'] = f32(sd[p + 'drop_path_ffn.scale'].reshape(-1))
        W['f.lnout.w'] = f32(sd[pre + 'ln_out.weight'].reshape(-1))
        W['f.lnout.b'] = f32(sd[pre + 'ln_out.bias'].reshape(-1))

    def _pack_video_net(self, W, sd, pre):
        """VideoTransformer weights (libs/modeling/video_net.py:32-121) under the state-dict prefix `pre` -> W['v.*'], W['e*']."""
        f32, conv = self._f32, self._conv_w
        C = self.C
        W['v.embd.w'] = conv(sd[pre + 'embd_fc.conv.weight'])
        W['v.embd.b'] = f32(sd[pre + 'embd_fc.conv.bias'])
        for i in range(self.arch[0]):
            W[f'v.conv{i}.w'] = conv(sd[pre + f'embd_convs.{i}.conv.weight'])
            W[f'v.norm{i}.w'] = f32(sd[pre + f'embd_norms.{i}.weight'].reshape(-1))
            W[f'v.norm{i}.b'] = f32(sd[pre + f'embd_norms.{i}.bias'].reshape(-1))
        self.enc_names = [pre + f'stem.{i}.' for i in range(self.arch[1])] + \
                         [pre + f'branch.{i}.' for i in range(self.arch[2])]
        for j, p in enumerate(self.enc_names):
            a = p + 'attn.'
            W[f'e{j}.ln.w'] = f32(sd[p + 'ln_attn.weight'].reshape(-1))
            W[f'e{j}.ln.b'] = f32(sd[p + 'ln_attn.bias'].reshape(-1))
            W[f'e{j}.dw'] = torch.stack([f32(sd[a + f'{n}_conv.conv.weight'].reshape(C, 3)) for n in 'qkv']).contiguous()
            W[f'e{j}.brn.w'] = torch.stack([f32(sd[a + f'{n}_norm.weight'].reshape(-1)) for n in 'qkv']).contiguous()
            W[f'e{j}.brn.b'] = torch.stack([f32(sd[a + f'{n}_norm.bias'].reshape(-1)) for n in 'qkv']).contiguous()
            W[f'e{j}.qkv.w'] = torch.stack([conv(sd[a + f'attn.{n}.weight']) for n in ('query', 'key', 'value')]).contiguous()
            W[f'e{j}.qkv.b'] = torch.stack([f32(sd[a + f'attn.{n}.bias']) for n in ('query', 'key', 'value')]).contiguous()
            W[f'e{j}.proj.w'] = conv(sd[a + 'attn.proj.weight'])
            W[f'e{j}.proj.b'] = f32(sd[a + 'attn.proj.bias'])
            W[f'e{j}.ls_attn'] = f32(sd[p + 'drop_path_attn.scale'].reshape(-1))
            W[f'e{j}.lnf.w'] = f32(sd[p + 'ln_ffn.weight'].reshape(-1))
            W[f'e{j}.lnf.b'] = f32(sd[p + 'ln_ffn.bias'].reshape(-1))
            W[f'e{j}.fc.w'] = conv(sd[p + 'ffn.fc.weight'])
            W[f'e{j}.fc.b'] = f32(sd[p + 'ffn.fc.bias'])
            W[f'e{j}.proj2.w'] = conv(sd[p + 'ffn.proj.weight'])
            W[f'e{j}.proj2.b'] = f32(sd[p + 'ffn.proj.bias'])
            W[f'e{j}.ls_ffn'] = f32(sd[p + 'drop_path_ffn.scale'].reshape(-1))

    def _pack(self, sd, cin_map):
        W = {}
        f32, conv = self._f32, self._conv_w
        C = self.C
        self._pack_text(W, sd, 'text_net.', self.opt['model']['text_net'])
        # ---- vid_map (K zero-padded to K0)
        wm = sd['vid_map.conv.weight'].detach().float()[:, :, 0]
        wpad = torch.zeros(C, self.K0)
        wpad[:, :cin_map] = wm
        W['map.w'] = self._act(wpad)
        W['map.b'] = f32(sd['vid_map.conv.bias'])
        # linear split of vid_map (decaf_map_combine): per-part weight blocks (n_parts, C, Cin) and the correl column
        parts = []
        if self.Ce_eff:
            parts.append(wm[:, :self.Ce_eff])
        if self.Cs_eff:
            parts.append(wm[:, self.Ce_eff:self.Ce_eff + self.Cs_eff])
        W['map.w2'] = self._act(torch.stack(parts).contiguous())
        W['map.wc'] = f32(wm[:, cin_map - 1]) if self.scat else None
        self.fusion_layers = self.opt['model']['fusion']['n_layers']
        self._pack_fusion(W, sd, 'fusion.')
        self._pack_video_net(W, sd, 'vid_net.')
        # ---- heads
        self.head_layers = self.opt['model']['cls_head']['n_layers']
        self.reg_layers = self.opt['model']['reg_head']['n_layers']
        for name, pre, nl, fin in (('h1', 'cls_head.', self.head_layers, 'cls_head'),
                                   ('h2', 'cls_head2.', self.head_layers, 'cls_head'),
                                   ('hr', 'reg_head.', self.reg_layers, 'reg_head')):
            for i in range(nl):
                W[f'{name}.conv{i}.w'] = conv(sd[f'{pre}convs.{i}.conv.weight'])
                W[f'{name}.norm{i}.w'] = f32(sd[f'{pre}norms.{i}.weight'].reshape(-1))
                W[f'{name}.norm{i}.b'] = f32(sd[f'{pre}norms.{i}.bias'].reshape(-1))
            W[f'{name}.out.w'] = conv(sd[f'{pre}{fin}.conv.weight'], act=False)     # (n_out, 3, C)
            W[f'{name}.out.b'] = f32(sd[f'{pre}{fin}.conv.bias'])
        W['hr.scales'] = torch.stack([f32(sd[f'reg_head.scales.{l}.scale'].reshape(())) for l in range(self.L)]).contiguous()
        # ---- TCN
        W['r.in.w'] = f32(sd['refine.conv_1x1.weight'].reshape(R_REFINE, self.L))
        W['r.in.b'] = f32(sd['refine.conv_1x1.bias'])
        for i in range(self.L):
            p = f'refine.layers.{i}.'
            W[f'r{i}.wd'] = f32(sd[p + 'conv_dilated.weight'])
            W[f'r{i}.bd'] = f32(sd[p + 'conv_dilated.bias'])
            W[f'r{i}.w1'] = f32(sd[p + 'conv_1x1.weight'].reshape(R_REFINE, R_REFINE))
            W[f'r{i}.b1'] = f32(sd[p + 'conv_1x1.bias'])
            W[f'r{i}.lnw'] = f32(sd[p + 'norm.weight'])
            W[f'r{i}.lnb'] = f32(sd[p + 'norm.bias'])
        W['r.out.w'] = f32(sd['refine.conv_out.weight'].reshape(R_REFINE, R_REFINE))
        W['r.out.b'] = f32(sd['refine.conv_out.bias'])
        # blobs of decaf_tcn_fused (layout: include/decaf_b200.h): Wd^T [cout][tap * R + cin] | W1^T [cout][cin] per layer
        bf = lambda t: t.to(torch.bfloat16).contiguous()
        W['r.wblob'] = bf(torch.cat([torch.cat((W[f'r{i}.wd'].permute(0, 2, 1).reshape(-1), W[f'r{i}.w1'].reshape(-1)))
                                     for i in range(self.L)]))
        W['r.vblob'] = torch.cat([torch.cat((W[f'r{i}.bd'], W[f'r{i}.b1'], W[f'r{i}.lnw'], W[f'r{i}.lnb'])) for i in range(self.L)]).contiguous()
        W['r.out.wb'] = bf(W['r.out.w'])
        self.W = W

    # ------------------------------------------------------------------ plans
    def level_lens(self, T):
        lens = [T]
        for _ in range(1, self.L):
            assert lens[-1] % 2 == 0, f'T={T} is not divisible by 2^(L-1)'
            lens.append(lens[-1] // 2)
        return lens

    def plan(self, B, T):
        key = (self.lane, B, T)
        if key in self._plans:
            return self._plans[key]
        p = _Plan()
        dev, ad = self.dev, self.act_dtype
        C, C2 = self.C, self.C2
        p.B, p.T = B, T
        p.lens = self.level_lens(T)
        p.lv = cabi.make_levels(p.lens)
        p.Pp = p.lv.Pp
        p.off = [p.lv.off[l] for l in range(self.L)]
        rows = B * T
        z = lambda *s, dtype=torch.float32: torch.zeros(*s, dtype=dtype, device=dev)
        p.correl = z(B, T)
        p.max_blocks = (T + self.sn - 1) // self.sn
        p.pooled = z(B, p.max_blocks)
        p.sel = z(B, T, dtype=torch.uint8)
        p.mask0 = z(B, T, dtype=torch.uint8)
        p.vid_len = z(1, dtype=torch.int32)
        p.hmask = z(B * p.Pp, dtype=torch.uint8)
        p.x0 = z(rows, self.K0, dtype=ad) if not self.linear_map else None
        n_parts = (1 if self.Ce_eff else 0) + (1 if self.Cs_eff else 0)
        p.x1 = z(T, n_parts * self.Cin, dtype=ad)              # [vid | shallow] of the video, channels-last
        p.ES = z(n_parts, T, C)                                # E = W_e vid, S = W_s shallow
        p.ones_T = torch.ones(T, dtype=torch.uint8, device=dev)
        p.XA = z(rows, C)
        p.XB = z(rows, C)
        p.SKIP = z(max(rows // 2, 1), C)
        p.A1 = z(3, rows, C, dtype=ad)
        p.QKV = z(3, rows, C, dtype=ad)
        p.ATT = z(rows, C, dtype=ad)
        p.SS = z(rows, 2 * C, dtype=ad)
        p.H4 = z(rows, 4 * C, dtype=ad)
        p.TMPF = z(rows, C)
        hrows = B * p.Pp
        p.CAT = z(hrows, C2, dtype=ad)
        p.HA = z(hrows, C2, dtype=ad)
        p.HB = z(hrows, C2, dtype=ad)
        p.TMPH = z(hrows, C2)
        p.logits1 = z(hrows)
        p.logits2 = z(hrows)
        p.offsets = z(hrows, 2)
        p.R0 = z(rows, R_REFINE)
        p.R1 = z(rows, R_REFINE)
        ev = self.opt['eval']
        p.topk = int(ev['pre_nms_topk'])
        p.cand_segs = z(B, p.topk, 2)
        p.cand_scores = z(B, p.topk)
        p.cand_idx = z(B, p.topk, dtype=torch.int32)
        p.cand_count = z(B, dtype=torch.int32)
        nm = self.opt['nms']
        p.max_out = int(nm['max_num_segs']) if nm['max_num_segs'] > 0 else p.topk
        # final results in one buffer so the host needs a single D2H copy per video
        p.out_buf = z(B * (3 * p.max_out + 1))
        p.out_segs = p.out_buf[:2 * B * p.max_out].view(B, p.max_out, 2)
        p.out_scores = p.out_buf[2 * B * p.max_out:3 * B * p.max_out].view(B, p.max_out)
        p.out_count = p.out_buf[3 * B * p.max_out:].view(torch.int32)
        p.out_host = torch.zeros(B * (3 * p.max_out + 1)).pin_memory()
        p.nms_ws = z(int(cabi.nms_workspace_bytes(B, p.topk)), dtype=torch.uint8)
        self._plans[key] = p
        return p

    def drop_workspaces(self, B, T, Lmax, drop_plan=True, drop_text=True):
        """Release the per-shape workspaces of every lane for (B queries, T steps) / (B queries, Lmax tokens): called by the
        Evaluator's shape LRU.  The caller guarantees nothing is in flight on them (and that no CUDA graph still uses them)."""
        if drop_plan:
            for k in [k for k in self._plans if k[1] == B and k[2] == T]:
                del self._plans[k]
            self._pe_cache.pop(T, None)
        if drop_text:
            for k in list(self._text_ws):                  # ('fused'|'tc', lane, n, Lmax), ('kv', lane, n, Lmax + 1), (lane, n, Lmax)
                kind = k[0] if isinstance(k[0], str) else 'composed'
                if (k[-2], k[-1]) == ((B, Lmax + 1) if kind == 'kv' else (B, Lmax)):
                    del self._text_ws[k]

    def pe_table(self, T):
        if T not in self._pe_cache:
            vn = self.opt['model']['vid_net']
            self._pe_cache[T] = _sinusoid_pe(vn['max_seq_len'], self.C, T).to(self.dev)
        return self._pe_cache[T]

    # ------------------------------------------------------------------ text encoder
    def encode_text_batch(self, tokens, lens):
        """tokens: (n, Lmax, C_tok) fp32 channels-last on device (rows >= len zero), lens (n,)
        int32 on device.  Returns text (n, Lmax+1, C_t) fp32, kv_len = lens + 1 (int32) and the
        fusion layers' key/value projections of the text, kv (n_fusion, 2, n * (Lmax+1), C) fp32.
        Restates TextTransformer.forward (libs/modeling/text_net.py:158-188) for a padded batch;
        padded keys are masked with -inf exactly like the reference's kv_mask.
        One launch (decaf_text_encoder: a cluster of 8 CTAs per query) when the shape fits it,
        otherwise composed from the GEMM / LayerNorm / attention entry points."""
        n, Lmax, Ctok = tokens.shape
        assert Ctok == self.Ctok and tokens.dtype == torch.float32 and tokens.is_contiguous()
        tn = self.opt['model']['text_net']
        if self.text_tc and bool(cabi.device_is_sm100()) and Ctok % 8 == 0 and self.Ct % 32 == 0:
            # (fewer than 64 rows: decaf_gemm falls back to the SIMT kernel on the same bf16 operands - same arithmetic)
            cabi.gemm_tag = 'text'
            try:
                XT, kv_len = self._encode_text_tc(tokens, lens)
                return XT, kv_len, self._pack_text_kv(self._text_kv_tc(XT, n, Lmax + 1), kv_len, n, Lmax + 1)
            finally:
                cabi.gemm_tag = None
        if self.fused_text and cabi.text_encoder_supported(Lmax, self.Ct, Ctok, tn['n_heads'], self.text_layers, self.C,
                                                          self.fusion_layers):
            XT, kv_len, KV = self._encode_text_fused(tokens, lens)
            return XT, kv_len, self._pack_text_kv(KV, kv_len, n, Lmax + 1)
        XT, kv_len = self._encode_text_composed(tokens, lens)
        return XT, kv_len, self._pack_text_kv(self.text_kv(XT, n, Lmax + 1), kv_len, n, Lmax + 1)

    def _pack_text_kv(self, KV, kv_len, n, L1):
        """bf16 configuration: the fusion layers' text keys / values are converted ONCE per video into the shared-memory
        image of the tensor-core cross-attention kernel (decaf_xattn_pack_kv) - on the text stream, next to the GEMM that
        produced them - instead of by each of the n * T / 128 CTAs of every fusion layer.  The image rides along as an
        attribute of the fp32 tensor (same lifetime as the workspace it belongs to); _fusion falls back to decaf_xattn for a
        key/value tensor that does not carry one."""
        if self.act_dtype != torch.bfloat16 or not cabi.xattn_packed_supported(L1, self.C, self.fusion_heads):
            return KV
        P = getattr(KV, '_decaf_packed', None)
        per = int(cabi.xattn_packed_elems(n, L1, self.C))
        if P is None or P.shape != (self.fusion_layers, per):
            P = torch.empty(self.fusion_layers, per, dtype=torch.bfloat16, device=self.dev)
            KV._decaf_packed = P
        for i in range(self.fusion_layers):
            cabi.xattn_pack_kv(KV[i][0], KV[i][1], kv_len, P[i], n, L1, self.C)
        return KV

    def _text_pe(self):
        """Raw sinusoid table (max_seq_len, C_t) of the text encoder, or None (libs/modeling/text_net.py:121-127).  The
        kernels pick or interpolate its rows PER QUERY from that query's own length: the reference encodes every query
        alone (libs/worker_v2.py:945-955), so neither the batch's longest query nor the length bucket may leak into it."""
        tn = self.opt['model']['text_net']
        if not tn.get('use_abs_pe', True):
            return None
        if 'text' not in self._pe_cache:
            self._pe_cache['text'] = _sinusoid_pe(tn['max_seq_len'], self.Ct, tn['max_seq_len']).to(self.dev)
        return self._pe_cache['text']

    def _text_blobs(self):
        """Weight / parameter blobs of decaf_text_encoder (layout: include/decaf_b200.h): every (stage, CTA)
        weight slice transposed and contiguous, every per-layer vector in one block."""
        if getattr(self, '_tblobs', None) is None:
            W, Ct, C = self.W, self.Ct, self.C
            CL, cpc, H = 8, self.Ct // 8, 4 * self.Ct
            m2 = lambda t: t.reshape(t.shape[0], -1)                      # conv (N, 1, K) -> (N, K)
            wl, pl = [], []
            ew = m2(W['t.embd.w'])
            wl += [ew[r * cpc:(r + 1) * cpc].t().contiguous().view(-1) for r in range(CL)]
            pl += [W['t.embd.b'], W['t.bkgd']]
            for i in range(self.text_layers):
                qkv = W[f't{i}.qkv.w'].reshape(3, Ct, Ct)
                wl += [torch.cat([qkv[j, r * cpc:(r + 1) * cpc] for j in range(3)], 0).t().contiguous().view(-1) for r in range(CL)]
                pw, fw, p2 = m2(W[f't{i}.proj.w']), m2(W[f't{i}.fc.w']), m2(W[f't{i}.proj2.w'])
                wl += [pw[r * cpc:(r + 1) * cpc].t().contiguous().view(-1) for r in range(CL)]
                wl += [fw[r * (H // CL):(r + 1) * (H // CL)].t().contiguous().view(-1) for r in range(CL)]
                wl += [p2[r * cpc:(r + 1) * cpc].t().contiguous().view(-1) for r in range(CL)]
                pl += [W[f't{i}.ln_attn.w'], W[f't{i}.ln_attn.b'], W[f't{i}.qkv.b'].reshape(-1), W[f't{i}.proj.b'],
                       W[f't{i}.ls_attn'], W[f't{i}.ln_ffn.w'], W[f't{i}.ln_ffn.b'], W[f't{i}.fc.b'], W[f't{i}.proj2.b'],
                       W[f't{i}.ls_ffn']]
            ns = 2 * C // CL
            for i in range(self.fusion_layers):
                kv = W[f'f{i}.kv.w'].reshape(2 * C, Ct)
                wl += [kv[r * ns:(r + 1) * ns].t().contiguous().view(-1) for r in range(CL)]
                pl += [W[f'f{i}.lnkv.w'], W[f'f{i}.lnkv.b'], W[f'f{i}.kv.b'].reshape(-1)]
            wblob, pblob = torch.cat(wl).contiguous(), torch.cat([x.reshape(-1) for x in pl]).contiguous()
            assert wblob.numel() == cabi.text_encoder_wblob_floats(Ct, self.Ctok, self.text_layers, C, self.fusion_layers)
            assert pblob.numel() == cabi.text_encoder_pblob_floats(Ct, self.text_layers, C, self.fusion_layers)
            self._tblobs = (wblob, pblob)
        return self._tblobs

    def _encode_text_fused(self, tokens, lens):
        Ct, C = self.Ct, self.C
        n, Lmax, Ctok = tokens.shape
        L1 = Lmax + 1
        key = ('fused', self.lane, n, Lmax)
        ws = self._text_ws.get(key)
        if ws is None:
            e = lambda *sh, dtype=torch.float32: torch.zeros(*sh, dtype=dtype, device=self.dev)
            ws = dict(XT=e(n, L1, Ct), KV=e(self.fusion_layers, 2, n * L1, C), kv_len=e(n, dtype=torch.int32))
            wblob, pblob = self._text_blobs()
            prm = cabi.TextEncoderParams()
            prm.n_query, prm.Lmax, prm.Ctok, prm.Ct = n, Lmax, Ctok, Ct
            prm.n_heads, prm.n_layers = self.opt['model']['text_net']['n_heads'], self.text_layers
            prm.n_fusion, prm.C = self.fusion_layers, C
            pe = self._text_pe()
            prm.wblob, prm.pblob, prm.pe = cabi.ptr(wblob), cabi.ptr(pblob), cabi.ptr(pe)
            prm.pe_rows = 0 if pe is None else int(pe.shape[0])
            prm.eps = 1e-5
            prm.text_out, prm.kv_out, prm.kv_len_out = cabi.ptr(ws['XT']), cabi.ptr(ws['KV']), cabi.ptr(ws['kv_len'])
            ws['prm'] = prm
            self._text_ws[key] = ws
        prm = ws['prm']
        prm.tokens, prm.lens = cabi.ptr(tokens), cabi.ptr(lens)
        cabi.text_encoder(prm)
        return ws['XT'], ws['kv_len'], ws['KV']

    def text_kv(self, text, n, L1):
        """Key/value projections of an already encoded text for every fusion layer (the composed
        path; libs/modeling/blocks.py:640-641, 348-350)."""
        W, Ct, C = self.W, self.Ct, self.C
        trow = n * L1
        key = ('kv', self.lane, n, L1)
        ws = self._text_ws.get(key)
        if ws is None:
            ws = dict(TLNF=torch.empty(trow, Ct, device=self.dev), KV=torch.empty(self.fusion_layers, 2, trow, C, device=self.dev))
            self._text_ws[key] = ws
        for i in range(self.fusion_layers):
            cabi.layernorm(text, Ct, 1, trow, w=W[f'f{i}.lnkv.w'], b=W[f'f{i}.lnkv.b'], out_f32=ws['TLNF'])
            cabi.gemm(ws['TLNF'], W[f'f{i}.kv.w'], C, Ct, 1, trow, bias=W[f'f{i}.kv.b'], out_f32=ws['KV'][i], n_group=2,
                      g_stride_a=0, g_stride_w=C * Ct, g_stride_bias=C, g_stride_out_f32=trow * C, impl=1)
        return ws['KV']

    def _text_w_bf16(self):
        if getattr(self, '_tw16', None) is None:
            W = self.W
            b16 = lambda t: t.to(torch.bfloat16).contiguous()
            d = {'embd': b16(W['t.embd.w'])}
            for i in range(self.text_layers):
                d[f'{i}.q'] = b16(W[f't{i}.qkv.w'][0])
                d[f'{i}.kv'] = b16(W[f't{i}.qkv.w'][1:])
                d[f'{i}.kv.b'] = W[f't{i}.qkv.b'][1:].contiguous()
                d[f'{i}.q.b'] = W[f't{i}.qkv.b'][0].contiguous()
                for k in ('proj', 'fc', 'proj2'):
                    d[f'{i}.{k}'] = b16(W[f't{i}.{k}.w'])
            for i in range(self.fusion_layers):
                d[f'f{i}.kv'] = b16(W[f'f{i}.kv.w'])
            self._tw16 = d
        return self._tw16

    def _encode_text_tc(self, tokens, lens):
        """TextTransformer.forward (libs/modeling/text_net.py:158-188) for the padded query batch on the tensor-core GEMM,
        bf16 operands / fp32 accumulation, residual stream, LayerNorm and softmax (the bf16 configuration's arithmetic)."""
        W, Ct = self.W, self.Ct
        Wb = self._text_w_bf16()
        n, Lmax, Ctok = tokens.shape
        L1 = Lmax + 1
        rows = n * L1
        dev = self.dev
        tn = self.opt['model']['text_net']
        key = ('tc', self.lane, n, Lmax)
        ws = self._text_ws.get(key)
        if ws is None:
            e = lambda *sh, dtype=torch.float32: torch.zeros(*sh, dtype=dtype, device=dev)
            bf = torch.bfloat16
            ws = dict(XT=e(n, L1, Ct), TOK=e(n, Lmax, Ctok, dtype=bf), TLN=e(rows, Ct, dtype=bf), Q=e(rows, Ct, dtype=bf),
                      KVt=e(2, rows, Ct), TATT=e(rows, Ct, dtype=bf), TH4=e(rows, 4 * Ct, dtype=bf),
                      tmask=e(n, L1, dtype=torch.uint8), kv_len=e(n, dtype=torch.int32),
                      ar=torch.arange(L1, device=dev, dtype=torch.int32),
                      KV=e(self.fusion_layers, 2, rows, self.C), TLNF=e(rows, Ct, dtype=bf))
            self._text_ws[key] = ws
        XT, kv_len, tmask = ws['XT'], ws['kv_len'], ws['tmask']
        cabi.text_init(XT, n, L1, Ct, lens, kv_len, tmask)
        cabi.cast_bf16(tokens, ws['TOK'])
        # embedding projection of the word tokens into rows 1..Lmax
        cabi.gemm(ws['TOK'], Wb['embd'], Ct, Ctok, n, Lmax, bias=W['t.embd.b'], rowmask=tmask.view(-1)[1:], m_seq_stride=L1,
                  out_f32=XT.view(-1)[Ct:], ldo=Ct, o_seq_stride=L1)
        cabi.text_prep(XT, n, L1, Ct, W['t.bkgd'], self._text_pe(), lens)
        TLN, Q, KVt, TATT, TH4 = ws['TLN'], ws['Q'], ws['KVt'], ws['TATT'], ws['TH4']
        nh = tn['n_heads']
        for i in range(self.text_layers):
            cabi.layernorm(XT, Ct, 1, rows, w=W[f't{i}.ln_attn.w'], b=W[f't{i}.ln_attn.b'], out_act=TLN)
            cabi.gemm(TLN, Wb[f'{i}.q'], Ct, Ct, 1, rows, bias=Wb[f'{i}.q.b'], out_act=Q)
            cabi.gemm(TLN, Wb[f'{i}.kv'], Ct, Ct, 1, rows, bias=Wb[f'{i}.kv.b'], out_f32=KVt, n_group=2, g_stride_a=0,
                      g_stride_w=Ct * Ct, g_stride_bias=Ct, g_stride_out_f32=rows * Ct)
            cabi.xattn(Q, KVt[0], KVt[1], TATT, n, L1, L1, Ct, nh, kv_len)
            cabi.gemm(TATT, Wb[f'{i}.proj'], Ct, Ct, 1, rows, bias=W[f't{i}.proj.b'], colscale=W[f't{i}.ls_attn'], resid=XT,
                      rowmask=tmask, out_f32=XT)
            cabi.layernorm(XT, Ct, 1, rows, w=W[f't{i}.ln_ffn.w'], b=W[f't{i}.ln_ffn.b'], out_act=TLN)
            cabi.gemm(TLN, Wb[f'{i}.fc'], 4 * Ct, Ct, 1, rows, bias=W[f't{i}.fc.b'], act=cabi.ACT_GELU, out_act=TH4)
            cabi.gemm(TH4, Wb[f'{i}.proj2'], Ct, 4 * Ct, 1, rows, bias=W[f't{i}.proj2.b'], colscale=W[f't{i}.ls_ffn'], resid=XT,
                      rowmask=tmask, out_f32=XT)
        return XT, kv_len

    def _text_kv_tc(self, text, n, L1):
        """Fusion layers' key / value projections of the encoded text (libs/modeling/blocks.py:640-641, 348-350), tensor cores."""
        W, Ct, C = self.W, self.Ct, self.C
        Wb = self._text_w_bf16()
        rows = n * L1
        ws = self._text_ws[('tc', self.lane, n, L1 - 1)]
        for i in range(self.fusion_layers):
            cabi.layernorm(text, Ct, 1, rows, w=W[f'f{i}.lnkv.w'], b=W[f'f{i}.lnkv.b'], out_act=ws['TLNF'])
            cabi.gemm(ws['TLNF'], Wb[f'f{i}.kv'], C, Ct, 1, rows, bias=W[f'f{i}.kv.b'], out_f32=ws['KV'][i], n_group=2,
                      g_stride_a=0, g_stride_w=C * Ct, g_stride_bias=C, g_stride_out_f32=rows * C)
        return ws['KV']

    def _encode_text_composed(self, tokens, lens):
        W, Ct = self.W, self.Ct
        n, Lmax, Ctok = tokens.shape
        L1 = Lmax + 1
        rows = n * L1
        dev = self.dev
        tn = self.opt['model']['text_net']
        ws = self._text_ws.get((self.lane, n, Lmax))
        if ws is None:
            e = lambda *sh: torch.empty(*sh, device=dev)
            ws = dict(XT=e(n, L1, Ct), TLN=e(rows, Ct), TQKV=e(3, rows, Ct), TATT=e(rows, Ct), TH4=e(rows, 4 * Ct),
                      tmask=torch.empty(n, L1, dtype=torch.uint8, device=dev), kv_len=torch.empty(n, dtype=torch.int32, device=dev),
                      ar=torch.arange(L1, device=dev, dtype=torch.int32))
            self._text_ws[(self.lane, n, Lmax)] = ws
        XT, kv_len, tmask = ws['XT'], ws['kv_len'], ws['tmask']
        cabi.text_init(XT, n, L1, Ct, lens, kv_len, tmask)
        # embedding projection of the word tokens into rows 1..Lmax
        cabi.gemm(tokens, W['t.embd.w'], Ct, Ctok, n, Lmax, bias=W['t.embd.b'],
                  rowmask=tmask.view(-1)[1:], m_seq_stride=L1,
                  out_f32=XT.view(-1)[Ct:], ldo=Ct, o_seq_stride=L1, impl=1)
        cabi.text_prep(XT, n, L1, Ct, W['t.bkgd'], self._text_pe(), lens)
        TLN, TQKV, TATT, TH4 = ws['TLN'], ws['TQKV'], ws['TATT'], ws['TH4']
        nh = tn['n_heads']
        for i in range(self.text_layers):
            cabi.layernorm(XT, Ct, 1, rows, w=W[f't{i}.ln_attn.w'], b=W[f't{i}.ln_attn.b'], out_f32=TLN)
            cabi.gemm(TLN, W[f't{i}.qkv.w'], Ct, Ct, 1, rows, bias=W[f't{i}.qkv.b'], out_f32=TQKV,
                      n_group=3, g_stride_a=0, g_stride_w=Ct * Ct, g_stride_bias=Ct,
                      g_stride_out_f32=rows * Ct, impl=1)
            cabi.xattn(TQKV[0], TQKV[1], TQKV[2], TATT, n, L1, L1, Ct, nh, kv_len)
            cabi.gemm(TATT, W[f't{i}.proj.w'], Ct, Ct, 1, rows, bias=W[f't{i}.proj.b'],
                      colscale=W[f't{i}.ls_attn'], resid=XT, rowmask=tmask, out_f32=XT, impl=1)
            cabi.layernorm(XT, Ct, 1, rows, w=W[f't{i}.ln_ffn.w'], b=W[f't{i}.ln_ffn.b'], out_f32=TLN)
            cabi.gemm(TLN, W[f't{i}.fc.w'], 4 * Ct, Ct, 1, rows, bias=W[f't{i}.fc.b'], act=cabi.ACT_GELU,
                      out_f32=TH4, impl=1)
            cabi.gemm(TH4, W[f't{i}.proj2.w'], Ct, 4 * Ct, 1, rows, bias=W[f't{i}.proj2.b'],
                      colscale=W[f't{i}.ls_ffn'], resid=XT, rowmask=tmask, out_f32=XT, impl=1)
        return XT, kv_len

    # ------------------------------------------------------------------ grounder forward
    def _g(self, A, Wt, N, K, n_seq, rows, **kw):
        if self.fp32_tc and A.dtype == torch.float32 and self._g_fp32_tc(A, Wt, N, K, n_seq, rows, kw):
            return
        kw.setdefault('impl', self.gemm_impl)
        cabi.gemm(A, Wt, N, K, n_seq, rows, **kw)

    def _g_fp32_tc(self, A, Wt, N, K, n_seq, rows, kw):
        """One fp32 GEMM / implicit conv of the FP32 configuration as a tcgen05 launch: both operands split into bf16
        hi / lo parts and concatenated along K (see decaf_split_bf16x3), fp32 accumulation, the same fused epilogue.
        Returns False when the call does not fit (the caller then takes the fp32-FMA kernel)."""
        G = kw.get('n_group', 1)
        lda = kw.get('lda') or K
        if K % 8 or lda % 4 or kw.get('ln') or n_seq * rows < 64:
            return False
        out32, outa = kw.get('out_f32'), kw.get('out_act')
        if out32 is not None and outa is not None and G != 1:
            return False
        a_ss = kw.get('a_seq_stride') or rows
        R = (n_seq - 1) * a_ss + rows                          # rows of A the launch addresses (per group)
        gsa = kw.get('g_stride_a', 0)
        n_a = G if (G > 1 and gsa) else 1
        need = n_a * R * 3 * K
        buf = self._a3.get(self.lane)
        if buf is None or buf.numel() < need:
            buf = self._a3[self.lane] = torch.empty(need, dtype=torch.bfloat16, device=self.dev)
        A3 = buf[:need].view(n_a, R, 3 * K)
        for g in range(n_a):
            cabi.split_bf16x3(A, R, K, lda, A3[g], 0, src_offset=g * gsa)
        key = (Wt.data_ptr(), tuple(Wt.shape))
        W3 = self._w3.get(key)
        if W3 is None:
            rows_w = Wt.numel() // K
            W3 = torch.empty(rows_w, 3 * K, dtype=torch.bfloat16, device=self.dev)
            cabi.split_bf16x3(Wt, rows_w, K, K, W3, 1)
            self._w3[key] = (W3, Wt)                           # (keeps the fp32 tensor alive: the key is its address)
        else:
            W3 = W3[0]
        kw2 = {k: v for k, v in kw.items() if k not in ('out_f32', 'out_act', 'ldo', 'ldo2', 'o_seq_stride', 'o2_seq_stride',
                                                         'g_stride_out_f32', 'g_stride_out_act', 'lda', 'g_stride_a',
                                                         'g_stride_w', 'impl')}
        if out32 is not None:
            o = dict(out_f32=out32, ldo=kw.get('ldo', 0), o_seq_stride=kw.get('o_seq_stride', 0),
                     g_stride_out_f32=kw.get('g_stride_out_f32', 0))
        else:                                                   # the act-dtype output IS fp32 in this configuration
            o = dict(out_f32=outa, ldo=kw.get('ldo2', 0), o_seq_stride=kw.get('o2_seq_stride', 0),
                     g_stride_out_f32=kw.get('g_stride_out_act', 0))
        cabi.gemm(A3, W3, N, 3 * K, n_seq, rows, lda=3 * K, g_stride_a=(R * 3 * K if n_a > 1 else 0),
                  g_stride_w=3 * kw.get('g_stride_w', 0), impl=0, **o, **kw2)
        if out32 is not None and outa is not None:              # second fp32 copy (FPN output into the head-input buffer)
            ld1, ss1 = kw.get('ldo') or N, kw.get('o_seq_stride') or rows
            ld2, ss2 = kw.get('ldo2') or N, kw.get('o2_seq_stride') or rows
            src = out32.as_strided((n_seq, rows, N), (ss1 * ld1, ld1, 1), out32.storage_offset())
            outa.as_strided((n_seq, rows, N), (ss2 * ld2, ld2, 1), outa.storage_offset()).copy_(src)
        return True

    def _encoder(self, p, j, X_in, T_in, stride, lvl_in, lvl_out, cat_level, phase=0):
        """One TransformerEncoder (libs/modeling/blocks.py:578-591).  X_in: (B, T_in, C) fp32
        stored masked.  Returns the buffer holding X_out (B, T_out, C)."""
        W, C, B = self.W, self.C, p.B
        T_out = T_in // stride
        rows = B * T_out
        mask_in = p.hmask[p.off[lvl_in]:]
        mask_out = p.hmask[p.off[lvl_out]:]
        A1 = p.A1.view(-1)[:3 * rows * C].view(3, rows, C)
        QKV = p.QKV.view(-1)[:3 * rows * C].view(3, rows, C)
        skip = p.SKIP if stride == 2 else None
        cabi.preattn(X_in, B, T_in, C, stride, mask_in, p.Pp, W[f'e{j}.ln.w'], W[f'e{j}.ln.b'], 3,
                     W[f'e{j}.dw'], W[f'e{j}.brn.w'], W[f'e{j}.brn.b'], A1, rows * C, skip_out=skip)
        self._g(A1, W[f'e{j}.qkv.w'], C, C, 1, rows, bias=W[f'e{j}.qkv.b'], out_act=QKV, n_group=3,
                g_stride_a=rows * C, g_stride_w=C * C, g_stride_bias=C, g_stride_out_act=rows * C)
        cabi.local_attn(QKV[0], QKV[1], QKV[2], p.ATT, B, T_out, C, self.n_heads, self.win, mask_out, p.Pp, phase=phase)
        if stride == 2:
            X_out = p.XB if X_in.data_ptr() == p.XA.data_ptr() else p.XA
            resid = p.SKIP
        else:
            X_out = X_in
            resid = X_in
        self._g(p.ATT, W[f'e{j}.proj.w'], C, C, B, T_out, bias=W[f'e{j}.proj.b'], colscale=W[f'e{j}.ls_attn'],
                resid=resid, rowmask=mask_out, m_seq_stride=p.Pp, out_f32=X_out)
        cabi.layernorm(X_out, C, 1, rows, w=W[f'e{j}.lnf.w'], b=W[f'e{j}.lnf.b'], out_act=p.A1[0])
        cat = dict(out_act=p.CAT[p.off[cat_level]:], ldo2=self.C2, o2_seq_stride=p.Pp) if cat_level is not None else {}
        # (cat: FPN output, also written in the act dtype into the head-input buffer)
        if self._use_fused_ffn(rows):
            cabi.ffn(p.A1[0], W[f'e{j}.fc.w'], W[f'e{j}.fc.b'], W[f'e{j}.proj2.w'], W[f'e{j}.proj2.b'], C, B, T_out,
                     colscale=W[f'e{j}.ls_ffn'], resid=X_out, rowmask=mask_out, m_seq_stride=p.Pp, out_f32=X_out, **cat)
        else:
            self._g(p.A1[0], W[f'e{j}.fc.w'], 4 * C, C, 1, rows, bias=W[f'e{j}.fc.b'], act=cabi.ACT_GELU, out_act=p.H4)
            self._g(p.H4, W[f'e{j}.proj2.w'], C, 4 * C, B, T_out, bias=W[f'e{j}.proj2.b'], colscale=W[f'e{j}.ls_ffn'],
                    resid=X_out, rowmask=mask_out, m_seq_stride=p.Pp, out_f32=X_out, **cat)
        return X_out

    def _use_fused_ffn(self, rows):
        """decaf_ffn streams the whole (2 x 4C x C) weight pair per 256 rows out of L2 and is bound by that stream: measured
        (tools/bench_ffn.py, profiles/) it beats the fc + proj GEMM pair at every size for C = 128 and from ~16k rows for
        C = 256 (36,864 rows: 60 vs 64 us, 18,432: 34 vs 40 us); below that the weight-resident fc GEMM wins."""
        return self.fused_ffn and (self.C <= 128 or rows >= self.ffn_min_rows)

    def _ln_fusable(self, N, rows):
        return self.fuse_ln and N <= 512 and rows >= 64

    def _tower(self, p, name, n_layers, x, ldx, Cw):
        """2 x (k3 conv -> LN -> ReLU) over the padded flat layout (libs/modeling/head.py:55-58)."""
        W = self.W
        hrows = p.B * p.Pp
        bufs = (p.HA, p.HB)
        cur, ld = x, ldx
        # the LayerNorm epilogue needs the whole output row in one CTA's TMEM (<= 512 fp32 columns) and a tensor-core
        # tile's worth of rows; wider rows (embd 512: C + 32 = 544) take the conv -> row-wise LayerNorm pair instead
        fuse = self._ln_fusable(Cw, hrows)
        for i in range(n_layers):
            dst = bufs[i % 2]
            if fuse:
                self._g(cur, W[f'{name}.conv{i}.w'], Cw, Cw, 1, hrows, lda=ld, taps=3, ln=True, ln_w=W[f'{name}.norm{i}.w'],
                        ln_b=W[f'{name}.norm{i}.b'], act=cabi.ACT_RELU, rowmask=p.hmask, out_act=dst, ldo2=Cw)
            else:
                self._g(cur, W[f'{name}.conv{i}.w'], Cw, Cw, 1, hrows, lda=ld, taps=3, out_f32=p.TMPH, ldo=Cw)
                cabi.layernorm(p.TMPH, Cw, 1, hrows, ldx=Cw, w=W[f'{name}.norm{i}.w'], b=W[f'{name}.norm{i}.b'],
                               relu=True, rowmask=p.hmask, out_act=dst, ldo2=Cw)
            cur, ld = dst, Cw
        return cur, ld

    def _fusion(self, p, X, B, T, text, kv_len, text_kv, text_ready, out_act=None, out_f32=None):
        """XAttNFusion._forward (libs/modeling/fusion.py:56-66): n_layers x ConvXAttNLayer + FFN on the residual stream X
        (B * T, C) fp32 (updated in place), then ln_out (masked) into out_act / out_f32."""
        W, C = self.W, self.C
        rows = B * T
        L1 = text.shape[1]
        mask0 = p.mask0
        KVall = text_kv
        for i in range(self.fusion_layers):
            cabi.preattn(X, B, T, C, 1, mask0, T, W[f'f{i}.lnq.w'], W[f'f{i}.lnq.b'], 1, W[f'f{i}.dw'],
                         W[f'f{i}.qn.w'], W[f'f{i}.qn.b'], p.A1[0], rows * C)
            self._g(p.A1[0], W[f'f{i}.q.w'], C, C, 1, rows, bias=W[f'f{i}.q.b'], out_act=p.QKV[0])
            if i == 0:
                if text_ready is not None:
                    text_ready()
                if KVall is None:
                    KVall = self.text_kv(text, B, L1)
            packed = getattr(KVall, '_decaf_packed', None)
            if packed is not None and self.act_dtype == torch.bfloat16:
                cabi.xattn_packed(p.QKV[0], packed[i], p.ATT, B, T, L1, C, self.fusion_heads, kv_len)
            else:
                KV = KVall[i]
                cabi.xattn(p.QKV[0], KV[0], KV[1], p.ATT, B, T, L1, C, self.fusion_heads, kv_len)
            self._g(p.ATT, W[f'f{i}.proj.w'], 2 * C, C, 1, rows, bias=W[f'f{i}.proj.b'], out_act=p.SS)
            cabi.adaln(X, rows, C, p.SS, mask0, W[f'f{i}.lnf.w'], W[f'f{i}.lnf.b'], X, p.A1[0])
            if self._use_fused_ffn(rows):
                cabi.ffn(p.A1[0], W[f'f{i}.fc.w'], W[f'f{i}.fc.b'], W[f'f{i}.proj2.w'], W[f'f{i}.proj2.b'], C, 1, rows,
                         colscale=W[f'f{i}.ls_ffn'], resid=X, rowmask=mask0, out_f32=X)
            else:
                self._g(p.A1[0], W[f'f{i}.fc.w'], 4 * C, C, 1, rows, bias=W[f'f{i}.fc.b'], act=cabi.ACT_GELU, out_act=p.H4)
                self._g(p.H4, W[f'f{i}.proj2.w'], C, 4 * C, 1, rows, bias=W[f'f{i}.proj2.b'], colscale=W[f'f{i}.ls_ffn'],
                        resid=X, rowmask=mask0, out_f32=X)
        cabi.layernorm(X, C, 1, rows, w=W['f.lnout.w'], b=W['f.lnout.b'], rowmask=mask0, out_act=out_act, out_f32=out_f32)

    def _backbone_steps(self, p, src, K_in, B, T, window, halo_steps):
        """VideoTransformer.forward (libs/modeling/video_net.py:123-164) from src (B * T, K_in) act dtype: embd_fc, the
        embedding convs (+ PE on the last), stem and branch encoders; FPN level l lands in the residual buffers and (act
        dtype) in p.CAT.  Generator: yields (level, X, cat, rows) after every encoder when halo_steps (see forward_steps)."""
        W, C, C2 = self.W, self.C, self.C2
        rows = B * T
        mask0 = p.mask0
        X = p.XA
        self._g(src, W['v.embd.w'], C, K_in, 1, rows, bias=W['v.embd.b'], rowmask=mask0, out_act=p.A1[1])
        n_convs = self.arch[0]
        use_pe = self.opt['model']['vid_net']['use_abs_pe']
        for i in range(n_convs):
            last = i == n_convs - 1
            pe = None
            if last and use_pe:
                pe = self.pe_table(T) if window is None else self.pe_table(window[0])[window[1]:window[1] + T]
            if self._ln_fusable(C, rows):
                a1, a2 = (p.A1[1], p.A1[2]) if i % 2 == 0 else (p.A1[2], p.A1[1])
                self._g(a1, W[f'v.conv{i}.w'], C, C, B, T, taps=3, ln=True, ln_w=W[f'v.norm{i}.w'], ln_b=W[f'v.norm{i}.b'],
                        act=cabi.ACT_RELU, pe=pe, rowmask=mask0, m_seq_stride=T,
                        out_f32=X if last else None, out_act=None if last else a2)
            else:
                self._g(p.A1[1], W[f'v.conv{i}.w'], C, C, B, T, taps=3, out_f32=p.TMPF)
                cabi.layernorm(p.TMPF, C, B, T, w=W[f'v.norm{i}.w'], b=W[f'v.norm{i}.b'], relu=True, pe=pe, rowmask=mask0,
                               out_f32=X if last else None, out_act=None if last else p.A1[1])
        self._cap('embed', X.view(B, T, C))
        j = 0
        # rows next to a window edge that differ from the unsharded run (time shards only): every k = 3 convolution adds one
        # row, every encoder its depthwise conv (1) + half window; a stride-2 encoder halves what it inherits.  After a halo
        # exchange the count restarts from zero.
        s_half = self.win // 2
        inv = self.fusion_layers + n_convs
        w_first = 0 if window is None else int(window[1])      # global step of the window's first row (0 when unsharded)
        for _ in range(self.arch[1]):                               # stem (stride 1, not an FPN level)
            X = self._encoder(p, j, X, T, 1, 0, 0, None, phase=w_first & 15)
            inv += 1 + s_half
            if halo_steps:
                yield 0, X.view(B, T, C), None, inv
                inv = 0
            j += 1
        T_l = T
        for l in range(self.L):                                     # branch -> FPN
            stride = 2 if l > 0 else 1
            X = self._encoder(p, j, X, T_l, stride, max(l - 1, 0), l, l, phase=(w_first >> l) & 15)
            T_l //= stride
            inv = (inv + 1 if stride == 1 else (inv + 2) // 2) + s_half
            if halo_steps:
                cat = p.CAT.view(B, p.Pp, C2)[:, p.off[l]:p.off[l] + T_l, :C]
                yield l, X.view(-1)[:B * T_l * C].view(B, T_l, C), cat, inv
                inv = 0
            self._cap(f'fpn{l}', X.view(-1)[:B * T_l * C].view(B, T_l, C))
            j += 1

    def forward(self, vid, shallow, vid_mask, text, kv_len, text_cls, text_kv=None, text_ready=None, window=None,
                exchange=None):
        """The grounder forward (see forward_steps for the arguments).  exchange: None, or a callable (level, X, cat, rows)
        invoked after every encoder output when the timeline is time-sharded with per-layer halo exchange: it must
        overwrite the outermost `rows` rows on each interior side of X (B, T_l, C) fp32 — and of cat, the bf16 copy the
        heads read, (B, T_l, C) strided or None — with the neighbour shard's values."""
        gen = self.forward_steps(vid, shallow, vid_mask, text, kv_len, text_cls, text_kv=text_kv, text_ready=text_ready,
                                 window=window, halo_steps=exchange is not None)
        try:
            while True:
                level, X, cat, rows = next(gen)
                exchange(level, X, cat, rows)
        except StopIteration as e:
            return e.value

    def forward_steps(self, vid, shallow, vid_mask, text, kv_len, text_cls, text_kv=None, text_ready=None, window=None,
                      halo_steps=False):
        """Generator form of forward: with halo_steps it yields (level, X, cat, rows) after every encoder output (the points
        where time shards refresh their halo rows from their neighbours, decaf_b200/time_shard.py) and returns the plan;
        several shards living in one process are advanced in lockstep through these points.
        vid (Ce, T) / shallow (Cs, T) fp32 with T contiguous (the reference layout, zero
        padded), vid_mask (T,) uint8/bool, text (n, L1, C_t) fp32 from encode_text_batch, kv_len
        (n,) int32, text_cls (n, Cs) fp32, text_kv = the third return value of encode_text_batch
        (computed here when None) — all on device.  text_ready: optional callable invoked right
        before the first kernel that reads text_kv / kv_len (the caller may have produced them on
        another stream: everything before that point depends on the video only).
        window: None, or (T_global, w0) when vid / shallow are the columns [w0, w0 + T) of a longer
        timeline processed by time shards (decaf_b200.time_shard): the caller has already filled
        plan.correl / plan.sel / plan.mask0 for this window from the GLOBAL saliency selection, and
        the positional encoding is the global table's rows [w0, w0 + T).  Fills plan.logits2 /
        offsets / hmask and returns the plan."""
        W, C, C2 = self.W, self.C, self.C2
        T = vid.shape[-1]
        B, L1, Ct = text.shape
        p = self.plan(B, T)
        rows = B * T
        vm = vid_mask.view(torch.uint8) if vid_mask.dtype == torch.bool else vid_mask
        # (1) saliency -> exact top-k selection -> merge (libs/modeling/model.py:500-554)
        if window is None:
            cabi.saliency(shallow, text_cls, p.correl, shallow.shape[0], T, B, self.norm)
            cabi.select(p.correl, vm, p.sel, p.mask0, p.pooled, p.max_blocks, T, B, self.sn, self.sratio,
                        and_mask=not self.msf, vid_len_out=p.vid_len)
        cabi.build_masks(p.mask0, T, p.hmask, p.lv, B)
        X = p.XA
        if self.linear_map:
            # vid_map is linear and the selection a 0/1 row mask: project the video ONCE (T rows), combine per query
            n_parts = p.ES.shape[0]
            cabi.merge(vid if self.Ce_eff else None, self.Ce_eff, shallow if self.Cs_eff else None, self.Cs_eff,
                       None, False, p.ones_T, p.ones_T, p.x1, n_parts * self.Cin, T, 1)
            self._g(p.x1, W['map.w2'], C, self.Cin, 1, T, lda=n_parts * self.Cin, out_f32=p.ES, n_group=n_parts,
                    g_stride_a=self.Cin, g_stride_w=C * self.Cin, g_stride_out_f32=T * C)
            E = p.ES[0] if self.Ce_eff else None
            S = p.ES[n_parts - 1] if self.Cs_eff else None
            cabi.map_combine(E, S, W['map.b'], p.correl if self.scat else None, W['map.wc'], p.sel, p.mask0, X, T, C, B)
        else:
            cabi.merge(vid if self.Ce_eff else None, self.Ce_eff, shallow if self.Cs_eff else None, self.Cs_eff,
                       p.correl if self.scat else None, self.scat, p.sel, p.mask0, p.x0, self.K0, T, B)
            self._g(p.x0, W['map.w'], C, self.K0, 1, rows, bias=W['map.b'], rowmask=p.mask0, out_f32=X)
        self._cap('correl', p.correl); self._cap('sel', p.sel); self._cap('mask0', p.mask0)
        self._cap('vid_map', X.view(B, T, C))
        # (2) early fusion: XAttNFusion (libs/modeling/fusion.py:56-66)
        self._fusion(p, X, B, T, text, kv_len, text_kv, text_ready, out_act=p.A1[0])
        self._cap('fusion', p.A1[0].view(B, T, C))
        # (3) video backbone: VideoTransformer.forward (libs/modeling/video_net.py:123-164)
        yield from self._backbone_steps(p, p.A1[0], C, B, T, window, halo_steps)
        # (4) heads with iterative refinement: fuse_and_predict (libs/modeling/model.py:442-471)
        hrows = B * p.Pp
        h, ld = self._tower(p, 'h1', self.head_layers, p.CAT, C2, C)
        cabi.head_out(h, ld, hrows, C, W['h1.out.w'], W['h1.out.b'], 1, 0, None, p.lv, p.logits1)
        self._cap('logits1', p.logits1.view(B, p.Pp))
        m0 = p.hmask[p.off[0]:]
        if self.fused_tcn and self.act_dtype == torch.bfloat16 and cabi.tcn_fused_supported(self.L, self.L):
            # one launch: expand -> L dilated residual layers -> conv_out (state in shared memory, mma.sync)
            cabi.tcn_fused(p.logits1, p.hmask, p.lv, W['r.in.w'], W['r.in.b'], W['r.wblob'], W['r.vblob'], self.L,
                           W['r.out.wb'], W['r.out.b'], R_REFINE, p.CAT, C2, C, B, scratch=p.R0)
        else:
            cabi.tcn_in(p.logits1, p.hmask, p.lv, W['r.in.w'], W['r.in.b'], R_REFINE, p.R0, B)
            cur, nxt = p.R0, p.R1
            for i in range(self.L):
                cabi.tcn_layer(cur, nxt, m0, p.Pp, W[f'r{i}.wd'], W[f'r{i}.bd'], W[f'r{i}.w1'], W[f'r{i}.b1'],
                               W[f'r{i}.lnw'], W[f'r{i}.lnb'], R_REFINE, 2 ** i, B, T)
                cur, nxt = nxt, cur
            cabi.tcn_out(cur, m0, p.Pp, W['r.out.w'], W['r.out.b'], R_REFINE, p.CAT, C2, C, p.lv, B)
        if self.fused_tcn and cabi.refine_pyramid_supported(self.L):
            cabi.refine_pyramid(p.CAT, C2, C, R_REFINE, p.hmask, p.lv, B)
        else:
            for l in range(1, self.L):
                cabi.refine_pool(p.CAT, C2, C, R_REFINE, p.hmask, p.lv, l, B)
        self._cap('cat', p.CAT.view(B, p.Pp, C2))
        h, ld = self._tower(p, 'h2', self.head_layers, p.CAT, C2, C2)
        cabi.head_out(h, ld, hrows, C2, W['h2.out.w'], W['h2.out.b'], 1, 0, None, p.lv, p.logits2)
        h, ld = self._tower(p, 'hr', self.reg_layers, p.CAT, C2, C2)
        cabi.head_out(h, ld, hrows, C2, W['hr.out.w'], W['hr.out.b'], 2, 1, W['hr.scales'], p.lv, p.offsets)
        return p

    # ------------------------------------------------------------------ post-processing
    def decode(self, p):
        """Evaluator._collect_segments for every query of the plan (libs/worker_v2.py:1131-1187)."""
        ev = self.opt['eval']
        cabi.decode(p.logits2, p.offsets, p.hmask, p.lv, p.B, True, float(ev['pre_nms_thresh']), p.topk,
                    float(ev['seg_len_thresh']), p.cand_segs, p.cand_scores, p.cand_idx, p.cand_count)
        return p.cand_segs, p.cand_scores, p.cand_count

    def nms(self, p, data=None, meta=None):
        """batched_nms (libs/nms/nms.py:106-148) for every query + seconds conversion
        (libs/worker_v2.py:1113-1122) when `data` is given."""
        nm = self.opt['nms']
        prm = cabi.NmsParams()
        prm.mode = {None: 0, 'nms': 1, 'soft_nms': 2}[nm['mode']]
        prm.iou_thresh, prm.sigma, prm.min_score = float(nm['iou_thresh']), float(nm['sigma']), float(nm['min_score'])
        prm.max_num_segs, prm.voting_thresh = int(nm['max_num_segs']), float(nm['voting_thresh'])
        if meta is not None:            # device-resident {vid_stride, clip_stride, clip_size / 2, fps, duration}
            prm.to_seconds = 1
            prm.video_meta = meta.data_ptr()
        elif data is not None:
            prm.to_seconds = 1
            prm.vid_stride = float(self.opt['model'].get('vid_stride', 1))
            prm.clip_stride = float(data['clip_stride'])
            prm.half_clip_size = float(0.5 * data['clip_size'])
            prm.fps, prm.duration = float(data['fps']), float(data['duration'])
        cabi.batched_nms(p.cand_segs, p.cand_scores, p.cand_count, p.B, p.topk, prm, p.out_segs, p.out_scores,
                         p.out_count, p.nms_ws)
        return p.out_segs, p.out_scores, p.out_count

    # ------------------------------------------------------------------ views in the reference's output format
    def level_views(self, p):
        """(fpn_logits_list, fpn_offsets_list, fpn_masks_list) as returned by the reference model
        (libs/modeling/model.py:561-565): per query, per level (1,T_l) / (1,T_l,2) / (1,T_l)."""
        lg = p.logits2.view(p.B, p.Pp)
        of = p.offsets.view(p.B, p.Pp, 2)
        mk = p.hmask.view(p.B, p.Pp).view(torch.bool)
        L, off, lens = self.L, p.off, p.lens
        logits = [tuple(lg[b:b + 1, off[l]:off[l] + lens[l]] for l in range(L)) for b in range(p.B)]
        offsets = [tuple(of[b:b + 1, off[l]:off[l] + lens[l]] for l in range(L)) for b in range(p.B)]
        masks = [tuple(mk[b:b + 1, off[l]:off[l] + lens[l]] for l in range(L)) for b in range(p.B)]
        return logits, offsets, masks


# ---------------------------------------------------------------------- stand-alone sub-graphs behind the builder registries
def _module_device(module):
    dev = next(module.parameters()).device
    if dev.type != 'cuda':
        raise RuntimeError(f'{type(module).__name__} runs on CUDA only: call .cuda() first (decaf_b200 has no CPU fallback)')
    return dev


class TextNetRunner(GrounderEngine):
    """The text-encoder part of the engine for a stand-alone `make_text_net(opt)` module (libs/modeling/text_net.py:158-188):
    same launches as GrounderEngine.encode_text_batch, weights taken from the module's own state dict."""

    def __init__(self, module, text_opt, act_dtype=torch.float32):
        self.dev = _module_device(module)
        self.act_dtype = act_dtype
        self.opt = {'model': {'text_net': dict(text_opt)}}
        self.Ct, self.Ctok, self.C = text_opt['embd_dim'], text_opt['in_dim'], text_opt['embd_dim']
        self.fusion_layers = 0
        self.gemm_impl = 0 if act_dtype == torch.bfloat16 else 1
        self.text_tc = act_dtype == torch.bfloat16 and bool(cabi.device_is_sm100())
        self.fused_text = False
        self._pe_cache, self._text_ws, self.lane = {}, {}, 0
        self._tblobs = self._tw16 = None
        W = {}
        self._pack_text(W, module.state_dict(), '', self.opt['model']['text_net'])
        self.W = W

    def encode(self, tokens, lens):
        if self.text_tc and self.Ctok % 8 == 0 and self.Ct % 32 == 0:
            return self._encode_text_tc(tokens, lens)[0]
        return self._encode_text_composed(tokens, lens)[0]


def run_text_net(module, text_opt, x, mask):
    """TextTransformer.forward(x (bs, C_tok, L), mask (bs, 1, L) or (bs, L)) -> (x (bs, C_t, L + 1), mask (bs, 1, L + 1))."""
    if mask.dim() == 2:
        mask = mask.unsqueeze(1)
    r = getattr(module, '_runner', None)
    if r is None:
        r = module._runner = TextNetRunner(module, text_opt, act_dtype=getattr(module, 'act_dtype', torch.float32))
    tok = x.float().transpose(1, 2).contiguous()
    lens = mask.reshape(mask.size(0), -1).sum(dim=1).to(torch.int32)
    tok = tok * mask.reshape(mask.size(0), -1, 1).to(tok.dtype)            # rows >= len must be zero
    out = r.encode(tok, lens)                                               # (bs, L + 1, C_t)
    return out.transpose(1, 2).contiguous(), torch.cat((mask[..., :1], mask), dim=-1)


def run_head(module, fpn, fpn_masks, n_out, fin, scales=None, act_dtype=None):
    """ClsHead.forward / RegHead.forward (libs/modeling/head.py:53-64, 95-108) for a stand-alone `make_head(opt)` module:
    fpn = tuple over levels of (bs, C, T_l), fpn_masks = tuple of (bs, 1, T_l) bool -> tuple of (bs, T_l) logits or
    (bs, T_l, 2) offsets and the squeezed masks.  Same kernels as the engine: all levels in one padded flat point layout,
    k3 conv -> LayerNorm -> ReLU towers on the tcgen05 GEMM (or its fp32 twin), final k3 conv + Scale + ReLU in decaf_head_out."""
    dev = _module_device(module)
    act_dtype = act_dtype or getattr(module, 'act_dtype', torch.bfloat16 if cabi.device_is_sm100() else torch.float32)
    B, C = fpn[0].size(0), fpn[0].size(1)
    lens = [int(x.size(-1)) for x in fpn]
    lv = cabi.make_levels(lens)
    Pp, hrows = lv.Pp, B * lv.Pp
    x = torch.zeros(B, Pp, C, dtype=act_dtype, device=dev)
    hmask = torch.zeros(B, Pp, dtype=torch.uint8, device=dev)
    for l, (f, m) in enumerate(zip(fpn, fpn_masks)):
        o = lv.off[l]
        x[:, o:o + lens[l]] = f.transpose(1, 2)
        hmask[:, o:o + lens[l]] = m.reshape(B, lens[l])
    sd = module.state_dict()
    conv = lambda w, act=True: (w.detach().permute(0, 2, 1).contiguous().to(dev, torch.float32)).to(act_dtype if act else torch.float32).contiguous()
    f32 = lambda t: t.detach().to(dev, torch.float32).contiguous()
    impl = 0 if act_dtype == torch.bfloat16 else 1
    fuse = act_dtype == torch.bfloat16 and bool(cabi.device_is_sm100()) and C <= 512 and hrows >= 64 and C % 8 == 0
    bufs = (torch.empty(hrows, C, dtype=act_dtype, device=dev), torch.empty(hrows, C, dtype=act_dtype, device=dev))
    tmp = None if fuse else torch.empty(hrows, C, device=dev)
    cur = x.view(hrows, C)
    n_layers = len(module.convs)
    for i in range(n_layers):
        w, nw, nb = conv(sd[f'convs.{i}.conv.weight']), f32(sd[f'norms.{i}.weight'].reshape(-1)), f32(sd[f'norms.{i}.bias'].reshape(-1))
        dst = bufs[i % 2]
        if fuse:
            cabi.gemm(cur, w, C, C, 1, hrows, taps=3, ln=True, ln_w=nw, ln_b=nb, act=cabi.ACT_RELU, rowmask=hmask.view(-1), out_act=dst, impl=impl)
        else:
            cabi.gemm(cur, w, C, C, 1, hrows, taps=3, out_f32=tmp, impl=impl)
            cabi.layernorm(tmp, C, 1, hrows, w=nw, b=nb, relu=True, rowmask=hmask.view(-1), out_act=dst)
        cur = dst
    out = torch.zeros(hrows, n_out, device=dev)
    sc = None if scales is None else torch.stack([f32(t.reshape(())) for t in scales]).contiguous()
    cabi.head_out(cur, C, hrows, C, conv(sd[f'{fin}.conv.weight'], act=False), f32(sd[f'{fin}.conv.bias']), n_out,
                  0 if scales is None else 1, sc, lv, out)
    out = out.view(B, Pp, n_out)
    outs, masks = tuple(), tuple()
    for l, m in enumerate(fpn_masks):
        o = lv.off[l]
        v = out[:, o:o + lens[l]]
        outs += ((v[..., 0] if n_out == 1 else v).contiguous(), )
        masks += (m.squeeze(1), )
    return outs, masks


class _PartialEngine(GrounderEngine):
    """Common scaffolding of the stand-alone sub-graph runners: the attributes GrounderEngine.plan / _encoder / _fusion read,
    without the rest of the model."""

    def _setup(self, module, C, n_heads, win, arch, act_dtype, Ct=None, max_seq_len=None, use_abs_pe=False):
        self.dev = _module_device(module)
        self.act_dtype = act_dtype
        self.gemm_impl = 0 if act_dtype == torch.bfloat16 else 1
        self.C, self.C2, self.Ct, self.Cin = C, C + R_REFINE, Ct or C, C
        self.n_heads, self.fusion_heads, self.win, self.arch, self.L = n_heads, n_heads, win, tuple(arch), arch[2]
        self.K0, self.linear_map, self.Ce_eff, self.Cs_eff = C, True, C, 0
        self.sn, self.sratio, self.msf, self.scat, self.sfonly, self.norm = 1, 0.0, False, False, False, False
        self.fusion_layers = 0
        self.opt = {'model': {'vid_net': {'max_seq_len': max_seq_len or 1, 'use_abs_pe': use_abs_pe}},
                    'eval': {'pre_nms_topk': 1}, 'nms': {'max_num_segs': 1}}
        self._plans, self._pe_cache, self._text_ws, self.lane, self.capture = {}, {}, {}, 0, None
        self.fuse_ln = act_dtype == torch.bfloat16 and bool(cabi.device_is_sm100())
        self.fused_tcn = True
        self.fused_ffn = act_dtype == torch.bfloat16 and bool(cabi.ffn_supported(C, cabi.BF16))
        self.ffn_min_rows = 16384
        self.fp32_tc = (act_dtype == torch.float32 and bool(cabi.device_is_sm100()) and os.environ.get('DECAF_FP32_TC', '1') != '0')
        self.gemm_impl = 0 if (act_dtype == torch.bfloat16 or self.fp32_tc) else 1
        self._w3, self._a3 = {}, {}
        self.W = {}


class VideoNetRunner(_PartialEngine):
    def __init__(self, module, kw, act_dtype):
        self._setup(module, kw['embd_dim'], kw['n_heads'], kw['mha_win_size'], kw['arch'], act_dtype,
                    max_seq_len=kw['max_seq_len'], use_abs_pe=kw['use_abs_pe'])
        self.in_dim = kw['in_dim']
        self._pack_video_net(self.W, module.state_dict(), '')

    def run(self, x, mask):
        B, K, T = x.shape
        assert T % (2 ** (self.L - 1)) == 0, f'T={T} must be divisible by 2^(levels-1)'
        p = self.plan(B, T)
        p.mask0.copy_(mask.reshape(B, T))
        cabi.build_masks(p.mask0, T, p.hmask, p.lv, B)
        src = x.float().transpose(1, 2).contiguous().view(B * T, K).to(self.act_dtype)
        fpn, masks = tuple(), tuple()
        hm = p.hmask.view(B, p.Pp)
        for level, X, cat, _ in self._backbone_steps(p, src, K, B, T, None, True):
            if cat is not None:                                     # a branch output = one FPN level
                fpn += (X.transpose(1, 2).contiguous(), )
                masks += (hm[:, p.off[level]:p.off[level] + X.shape[1]].bool().unsqueeze(1), )
        return fpn, masks


def run_video_net(module, kw, x, mask):
    """VideoTransformer.forward(x (bs, C_in, T), mask (bs, T) or (bs, 1, T)) -> (fpn, fpn_masks): tuples over the FPN levels of
    (bs, C, T_l) fp32 and (bs, 1, T_l) bool (libs/modeling/video_net.py:123-164)."""
    r = getattr(module, '_runner', None)
    if r is None:
        r = module._runner = VideoNetRunner(module, kw, getattr(module, 'act_dtype', None) or
                                            (torch.bfloat16 if cabi.device_is_sm100() else torch.float32))
    return r.run(x, mask)


class FusionRunner(_PartialEngine):
    def __init__(self, module, kw, act_dtype):
        self._setup(module, kw['vid_dim'], kw['n_heads'], 1, (0, 0, 1), act_dtype, Ct=kw['text_dim'])
        self.opt['model']['fusion'] = {'n_layers': kw['n_layers'], 'n_heads': kw['n_heads']}
        self.fusion_layers = kw['n_layers']
        self._pack_fusion(self.W, module.state_dict(), '')

    def run(self, q, q_mask, kv, kv_mask):
        B, C, T = q.shape
        p = self.plan(B, T)
        p.mask0.copy_(q_mask.reshape(B, T))
        X = p.XA
        X.view(B, T, C).copy_(q.float().transpose(1, 2))
        text = kv.float().transpose(1, 2).contiguous()              # (B, L1, Ct)
        kv_len = kv_mask.reshape(B, -1).sum(dim=1).to(torch.int32)
        out = torch.empty(B * T, C, device=self.dev)
        self._fusion(p, X, B, T, text, kv_len, None, None, out_f32=out)
        return out.view(B, T, C).transpose(1, 2).contiguous(), q_mask


def run_fusion(module, kw, vid, vid_masks, text, text_mask, text_size=None):
    """XAttNFusion.forward (libs/modeling/fusion.py:56-78): vid (bs, C, T) or a tuple of them, masks (bs, 1, T), text (bs, C_t, L),
    text_mask (bs, 1, L) -> the fused (ln_out-normalised) video features and masks.  The key mask must be a prefix mask (the
    first kv_len tokens valid), which is what the reference produces."""
    assert text_size is None, 'text_size (query repetition) is a training-time option'
    r = getattr(module, '_runner', None)
    if r is None:
        r = module._runner = FusionRunner(module, kw, getattr(module, 'act_dtype', None) or
                                          (torch.bfloat16 if cabi.device_is_sm100() else torch.float32))
    if not isinstance(vid, tuple):
        return r.run(vid, vid_masks, text, text_mask)
    out, out_masks = tuple(), tuple()
    for x, m in zip(vid, vid_masks):
        y, ym = r.run(x, m, text, text_mask)
        out += (y, )
        out_masks += (ym, )
    return out, out_masks

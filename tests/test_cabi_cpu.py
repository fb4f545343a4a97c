"""CPU: the C-ABI library loads and exports every symbol include/decaf_b200.h declares; host-side
logic (level geometry, config plumbing, state-dict layout) — no compute calls."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from decaf_b200 import _cabi
    hdr = open(os.path.join(ROOT, 'include', 'decaf_b200.h')).read()
    declared = set(re.findall(r'\b(decaf_[a-z0-9_]+)\s*\(', hdr))
    assert declared, 'no declarations parsed'
    for name in declared:
        assert hasattr(_cabi.lib, name), f'{name} declared in include/decaf_b200.h but not exported'
    assert set(_cabi.EXPORTED) == declared
    assert _cabi.version() >= 100


def test_levels_geometry():
    from decaf_b200 import _cabi
    lv = _cabi.make_levels([64, 32, 16, 8])
    assert lv.n_levels == 4 and lv.Pp == 64 + 32 + 16 + 8 + 4 + 1
    assert [lv.off[i] for i in range(4)] == [1, 66, 99, 116]


def test_state_dict_layout_matches_reference_fixture():
    """The weight containers expose exactly the reference's parameter names and shapes (recorded in
    the golden fixtures from the instantiated reference model)."""
    from decaf_b200 import synth
    from decaf_b200.worker_v2 import create_model
    from golden_util import CASES, GOLDEN_DIR
    for name, (kw, *_rest) in CASES.items():
        g = np.load(os.path.join(GOLDEN_DIR, f'{name}.npz'))
        shapes = {k: tuple(int(x) for x in s.split(',') if x) for k, s in zip(g['state_keys'], g['state_shapes'])}
        opt = synth.tiny_opt(**kw)
        model = create_model(opt)
        sd = model.state_dict()
        assert set(sd) == set(shapes), name
        assert all(tuple(sd[k].shape) == shapes[k] for k in sd), name
        model.load_state_dict(synth.fill_state_dict(shapes, 1))     # strict load works
        # the constructor mutates opt like the reference (libs/modeling/model.py:426-428)
        assert opt.model.cls_head.embd_dim == opt.model.vid_net.embd_dim + 32


def test_registries_and_errors():
    from decaf_b200 import synth
    from decaf_b200.modeling import make_fusion, make_head, make_text_net, make_video_net
    import torch
    opt = synth.tiny_opt()
    v = make_video_net(dict(opt.model.vid_net, in_dim=opt.model.vid_net.embd_dim))
    assert len(v.branch) == opt.model.vid_net.arch[2]
    assert make_text_net(opt.model.text_net).bkgd_token.shape == (opt.model.text_net.embd_dim, 1)
    assert len(make_fusion(opt.model.fusion).layers) == 2
    assert make_head(opt.model.cls_head).cls_head.conv.weight.shape[0] == 1
    with pytest.raises(KeyError):
        make_head(dict(name='nope'))
    with pytest.raises(RuntimeError):                                # weight containers never compute
        v(torch.zeros(1, 64, 8), torch.ones(1, 8, dtype=torch.bool))
    # no CUDA here: the product path must fail loudly, not fall back
    if not torch.cuda.is_available():
        from decaf_b200.worker_v2 import create_model
        m = create_model(synth.tiny_opt())
        with pytest.raises(RuntimeError):
            m.engine()


def test_evaluator_padding_rule():
    from decaf_b200 import synth
    import torch
    if not torch.cuda.is_available():
        pytest.skip('Evaluator construction allocates on the device')


def test_oracle_state_shapes_match_fixtures_and_mirror():
    """oracle/state_shapes.py (what bench.py's reference arm builds its weights from, without the product package) against
    the reference's own state-dict layout recorded in the golden fixtures, and against the product mirror on the bench /
    sweep configurations."""
    from decaf_b200 import synth
    from decaf_b200.worker_v2 import create_model
    from golden_util import CASES, GOLDEN_DIR
    from oracle.state_shapes import state_dict_shapes
    for name, (kw, *_rest) in CASES.items():
        g = np.load(os.path.join(GOLDEN_DIR, f'{name}.npz'))
        shapes = {k: tuple(int(x) for x in s.split(',') if x) for k, s in zip(g['state_keys'], g['state_shapes'])}
        assert state_dict_shapes(synth.tiny_opt(**kw)) == shapes, name
    for opt in (synth.nlq_opt(), synth.charades_opt(), synth.nlq_opt(embd_dim=512)):
        want = {k: tuple(v.shape) for k, v in create_model(opt.clone()).state_dict().items()}
        assert state_dict_shapes(opt) == want


def test_bench_reference_arm_does_not_load_the_product_library():
    """`bench.py --impl reference` must time the reference-side CPU path only: building its problem may not import the
    decaf_b200 package or load libdecaf_b200.so (checked in a fresh interpreter)."""
    import subprocess
    import sys
    code = ("import sys, bench; bench.VID_LEN = 64; bench.N_QUERY = 2; opt, sd, v = bench.make_problem(0, 1);"
            "from oracle import grounder_oracle, nms_oracle;"
            "assert not any(m == 'decaf_b200' or m.startswith('decaf_b200.') for m in sys.modules), sorted(sys.modules);"
            "assert 'libdecaf_b200' not in open('/proc/self/maps').read(); print('clean')")
    out = subprocess.run([sys.executable, '-c', code], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and 'clean' in out.stdout, out.stderr[-2000:]

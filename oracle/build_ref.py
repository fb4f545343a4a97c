"""ORACLE tooling (test infrastructure, NOT product code).

Compiles (a) the reference's own CPU NMS extension, unmodified, from where it lies under
/root/reference (libs/nms/src/nms_cpu.cpp -> oracle/_ref/nms_1d_cpu_vg.so; only possible in
the build container, the GPU box uses the prebuilt file that travels with the snapshot), and
(b) this repo's plain-C restatement oracle/nms_oracle.c -> oracle/libnms_oracle.so.

No reference source is copied into the repo: g++ reads it in place and only the binary lands
in oracle/_ref/ (git-ignored).
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = '/root/reference/libs/nms/src/nms_cpu.cpp'
REF_OUT = os.path.join(HERE, '_ref', 'nms_1d_cpu_vg.so')
C_SRC = os.path.join(HERE, 'nms_oracle.c')
C_OUT = os.path.join(HERE, 'libnms_oracle.so')


def _newer(src, out):
    return (not os.path.exists(out)) or os.path.getmtime(src) > os.path.getmtime(out)


def build_c_oracle(verbose=False):
    if _newer(C_SRC, C_OUT):
        # -ffp-contract=off: the reference extension is built without FMA contraction
        # (plain x86-64 g++ -O3 via torch's BuildExtension), keep the same float semantics.
        cmd = ['gcc', '-O2', '-ffp-contract=off', '-fPIC', '-shared', '-o', C_OUT, C_SRC, '-lm']
        if verbose:
            print(' '.join(cmd))
        subprocess.check_call(cmd)
    return C_OUT


def build_reference_nms(verbose=False):
    """Returns the path of the compiled reference extension, or None when neither the
    reference sources nor a prebuilt binary are available."""
    if not os.path.exists(REF_SRC):
        return REF_OUT if os.path.exists(REF_OUT) else None
    if not _newer(REF_SRC, REF_OUT):
        return REF_OUT
    import torch
    from torch.utils import cpp_extension
    os.makedirs(os.path.dirname(REF_OUT), exist_ok=True)
    inc = cpp_extension.include_paths()
    inc.append(sysconfig.get_paths()['include'])
    libdir = os.path.join(os.path.dirname(torch.__file__), 'lib')
    cmd = ['g++', '-O3', '-fPIC', '-shared', '-std=c++17', '-fopenmp',
           '-DTORCH_EXTENSION_NAME=nms_1d_cpu_vg', '-DTORCH_API_INCLUDE_EXTENSION_H',
           f'-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}']
    cmd += [f'-I{p}' for p in inc]
    cmd += [REF_SRC, '-o', REF_OUT, f'-L{libdir}', '-ltorch', '-ltorch_cpu', '-lc10',
            '-ltorch_python', f'-Wl,-rpath,{libdir}']
    if verbose:
        print(' '.join(cmd))
    subprocess.check_call(cmd)
    return REF_OUT


def load_reference_nms():
    """Import the compiled reference pybind module (``nms``, ``softnms``) or return None."""
    path = REF_OUT if os.path.exists(REF_OUT) else None
    if path is None:
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch symbols must be loaded first)
    spec = importlib.util.spec_from_file_location('nms_1d_cpu_vg', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == '__main__':
    print(build_c_oracle(verbose=True))
    print(build_reference_nms(verbose=True))
    sys.exit(0)

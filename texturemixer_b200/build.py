"""Build libtmx.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension:
the library is a plain C-ABI shared object, see include/tmx.h)."""
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, 'csrc')
LIB_PATH = os.path.join(PKG_DIR, 'libtmx.so')
SOURCES = ['context.cu', 'pointwise.cu', 'conv_ffma.cu', 'conv_tc.cu', 'conv_api.cu', 'perm_host.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '--use_fast_math=false', '-Xcompiler', '-fPIC,-O2,-fvisibility=default', '-shared']


def _nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found')


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(PKG_DIR, '..', 'include', 'tmx.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    if not force and not needs_build():
        return LIB_PATH
    flags = [f for f in NVCC_FLAGS if not f.startswith('--use_fast_math')]
    cmd = [_nvcc()] + flags + (['-Xptxas', '-v'] if verbose else []) + \
          [os.path.join(CSRC, s) for s in SOURCES] + ['-o', LIB_PATH]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError('nvcc failed building libtmx.so')
    if verbose:
        print(res.stdout + res.stderr)
    return LIB_PATH


if __name__ == '__main__':
    print(build_library(force=True, verbose='-v' in sys.argv))

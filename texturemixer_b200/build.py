"""Build libtmx.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension:
the library is a plain C-ABI shared object, see include/tmx.h)."""
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, 'csrc')
LIB_PATH = os.path.join(PKG_DIR, 'libtmx.so')
SOURCES = ['context.cu', 'pointwise.cu', 'conv_ffma.cu', 'conv_tc.cu', 'conv_api.cu', 'perm_host.cu', 'optim.cu', 'backward.cu', 'conv_wgrad.cu', 'gram.cu', 'conv_lin.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '--use_fast_math=false', '-Xcompiler', '-fPIC,-O2,-fvisibility=default', '-shared']


def _nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found')


OBJ_DIR = os.path.join(PKG_DIR, 'build')
HEADERS = [os.path.join(CSRC, 'common.cuh'), os.path.join(CSRC, 'tc_common.cuh'),
           os.path.join(PKG_DIR, '..', 'include', 'tmx.h')]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + HEADERS
    return _stale(LIB_PATH, deps)


def build_library(force=False, verbose=False):
    """Compile every csrc/*.cu for sm_100a (objects in parallel, cached by mtime) and link libtmx.so."""
    if not force and not needs_build():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(OBJ_DIR, exist_ok=True)
    cflags = [f for f in NVCC_FLAGS if not f.startswith('--use_fast_math') and f != '-shared']
    nvcc = _nvcc()

    def compile_one(src):
        path = os.path.join(CSRC, src)
        obj = os.path.join(OBJ_DIR, src.replace('.cu', '.o'))
        if not force and not _stale(obj, [path] + HEADERS):
            return obj, ''
        cmd = [nvcc] + cflags + (['-Xptxas', '-v'] if verbose else []) + ['-c', path, '-o', obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError('nvcc failed on %s' % src)
        return obj, res.stdout + res.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        results = list(ex.map(compile_one, SOURCES))
    if verbose:
        print(''.join(log for _, log in results))
    cmd = [nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-Xcompiler', '-fPIC'] + \
          [obj for obj, _ in results] + ['-o', LIB_PATH]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError('nvcc failed linking libtmx.so')
    return LIB_PATH


if __name__ == '__main__':
    print(build_library(force='-f' in sys.argv, verbose='-v' in sys.argv))

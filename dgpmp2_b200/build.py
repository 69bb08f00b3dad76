"""Build libdgpmp2_b200.so in-tree with nvcc for sm_100a (no torch involvement).

    python -m dgpmp2_b200.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_DIR = os.path.join(HERE, 'lib')
LIB_PATH = os.path.join(LIB_DIR, 'libdgpmp2_b200.so')
SOURCES = ['c_abi.cu']
DEPS = ['c_abi.cu', 'kernels.cuh', 'bcr.cuh', 'bcr_plan.cuh', 'factors.cuh', 'mp.cuh', 'hd.cuh', os.path.join('..', '..', 'include', 'dgpmp2_b200.h')]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-shared', '--cudart', 'static']


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return 'nvcc'


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


LAST_ACTION = None     # 'compiled' or 'reused' after build()


def build(force=False, verbose=False):
    """Compile the CUDA library if missing or older than its sources. Returns the .so path."""
    global LAST_ACTION
    if not force and not is_stale():
        LAST_ACTION = 'reused'
        return LIB_PATH
    LAST_ACTION = 'compiled'
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [_nvcc()] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-o', LIB_PATH] + \
          [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n%s\n%s' % (' '.join(cmd), res.stdout))
    if verbose:
        print(res.stdout)
    return LIB_PATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))

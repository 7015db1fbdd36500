"""Build the C-ABI shared library of molgym_b200.

  python -m molgym_b200.build            -> molgym_b200/lib/libmolgym_b200.so      (nvcc, sm_100a; the product)
  python -m molgym_b200.build --cusim    -> tests/cusim/_build/libmolgym_b200_cusim.so  (g++ + tests/cusim/cusim.h;
                                            kernel-logic tests on machines without a GPU; never loaded by the package)
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
LIB_DIR = os.path.join(HERE, 'lib')
LIB_PATH = os.path.join(LIB_DIR, 'libmolgym_b200.so')
CUSIM_DIR = os.path.join(ROOT, 'tests', 'cusim')
CUSIM_LIB = os.path.join(CUSIM_DIR, '_build', 'libmolgym_b200_cusim.so')
SOURCES = ['api.cu']


def _sources_mtime():
    newest = 0.0
    for d in (CSRC, os.path.join(ROOT, 'include'), CUSIM_DIR):
        for f in os.listdir(d):
            p = os.path.join(d, f)
            if os.path.isfile(p):
                newest = max(newest, os.path.getmtime(p))
    return newest


def _fresh(path):
    return os.path.exists(path) and os.path.getmtime(path) >= _sources_mtime()


def build_cuda(force=False, verbose=False):
    if not force and _fresh(LIB_PATH):
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '--use_fast_math=false',
           '-Xcompiler', '-fPIC', '-shared', '-I', os.path.join(ROOT, 'include')]
    cmd = [c for c in cmd if c != '--use_fast_math=false']
    if verbose:
        cmd += ['-Xptxas', '-v']
    cmd += [os.path.join(CSRC, s) for s in SOURCES] + ['-o', LIB_PATH]
    subprocess.run(cmd, check=True)
    return LIB_PATH


def build_cusim(force=False):
    if not force and _fresh(CUSIM_LIB):
        return CUSIM_LIB
    os.makedirs(os.path.dirname(CUSIM_LIB), exist_ok=True)
    cmd = ['g++', '-O2', '-std=c++17', '-fPIC', '-shared', '-DMGB_CUSIM', '-x', 'c++', '-I', CUSIM_DIR, '-I',
           os.path.join(ROOT, 'include'), '-Wno-unknown-pragmas']
    cmd += [os.path.join(CSRC, s) for s in SOURCES] + ['-o', CUSIM_LIB, '-lpthread']
    subprocess.run(cmd, check=True)
    return CUSIM_LIB


def packer_path():
    import sysconfig
    return os.path.join(LIB_DIR, '_mgb_packer' + sysconfig.get_config_var('EXT_SUFFIX'))


def build_packer(force=False):
    """CPython extension that flattens observation tuples (host side of the packing path)."""
    import sysconfig
    out = packer_path()
    src = os.path.join(CSRC, 'packer.c')
    if not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = ['gcc', '-O2', '-shared', '-fPIC', '-I', sysconfig.get_paths()['include'], src, '-o', out]
    subprocess.run(cmd, check=True)
    return out


if __name__ == '__main__':
    if '--cusim' in sys.argv:
        print(build_cusim(force=True))
    elif '--packer' in sys.argv:
        print(build_packer(force=True))
    else:
        print(build_cuda(force=True, verbose='-v' in sys.argv))

"""In-tree build of the CUDA library (sm_100a only).

``python -m gaboflow_b200.build`` or ``gaboflow_b200.build.build()`` compiles every ``csrc/*.cu`` into
``gaboflow_b200/libgabo_b200.so`` with nvcc (cross-compiles without a GPU).  The ``.so`` is git-ignored but
travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, 'csrc')
LIB_PATH = os.path.join(PKG_DIR, 'libgabo_b200.so')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-O3']


def _nvcc() -> str:
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found: the gaboflow_b200 CUDA library cannot be built')


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, '*.cuh')) + glob.glob(os.path.join(PKG_DIR, '..', 'include', '*.h'))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu -> libgabo_b200.so (objects in parallel, then one link)."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = _nvcc()
    objdir = os.path.join(PKG_DIR, 'build')
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, '-c', src, '-o', obj] + (['-Xptxas', '-v'] if verbose else [])
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}:\n{out}')
        if verbose:
            print(out)
    tmp = LIB_PATH + '.tmp'
    subprocess.run([nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-o', tmp, *objs, '-lcudart'], check=True)
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))

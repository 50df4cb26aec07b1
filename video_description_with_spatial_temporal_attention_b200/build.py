"""Compile ``csrc/*.cu`` into ``libstat_b200.so`` for sm_100a with nvcc (in-tree, so
the library travels with the repository snapshot)."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT = os.path.join(HERE, 'libstat_b200.so')
OBJ = os.path.join(HERE, 'build')
SOURCES = ('gemm_tf32x3.cu', 'att_step.cu', 'att_group.cu', 'recurrent.cu', 'optim.cu', 'backward.cu', 'step_fused.cu', 'cell_step.cu', 'stat_api.cu')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
         '-Xcompiler', '-fPIC']


def _nvcc():
    n = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.isfile(n):
        raise RuntimeError('nvcc not found')
    return n


def _digest():
    h = hashlib.sha256()
    root = os.path.dirname(HERE)
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    files.append(os.path.join(root, 'include', 'stat_b200.h'))
    for f in files:
        with open(f, 'rb') as fh:
            h.update(f.encode() + b'\0' + fh.read())
    h.update(' '.join(FLAGS).encode())
    return h.hexdigest()


def build_lib(force=False, verbose=False):
    """Returns the path of the library, compiling it when the sources changed."""
    stamp = os.path.join(OBJ, 'stamp')
    dig = _digest()
    if not force and os.path.isfile(OUT) and os.path.isfile(stamp):
        with open(stamp) as fh:
            if fh.read().strip() == dig:
                return OUT
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()

    def cc(src):
        obj = os.path.join(OBJ, src[:-3] + '.o')
        cmd = [nvcc] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', os.path.join(CSRC, src), '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed on %s:\n%s\n%s' % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(cc, SOURCES))
    r = subprocess.run([nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', OUT] + objs,
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
    with open(stamp, 'w') as fh:
        fh.write(dig)
    return OUT


if __name__ == '__main__':
    print(build_lib(force='--force' in sys.argv, verbose='-v' in sys.argv))

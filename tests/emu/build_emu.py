"""TEST INFRASTRUCTURE: compiles csrc/backward.cu with g++ against tests/emu/cuda_emu.h (CPU emulation of
the CUDA launch model) into tests/emu/_build/libstat_bw_emu.so, so that the kernels and the host
orchestration of the backward pass can be checked against the gradient oracle without a GPU.
The package never loads this library."""
from __future__ import annotations

import hashlib
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(ROOT, 'video_description_with_spatial_temporal_attention_b200', 'csrc', 'backward.cu')
OUT_DIR = os.path.join(HERE, '_build')
OUT = os.path.join(OUT_DIR, 'libstat_bw_emu.so')


def build(force=False):
    h = hashlib.sha256()
    for f in (SRC, os.path.join(HERE, 'cuda_emu.h'), os.path.join(ROOT, 'include', 'stat_b200.h')):
        with open(f, 'rb') as fh:
            h.update(fh.read())
    stamp = os.path.join(OUT_DIR, 'stamp')
    if not force and os.path.isfile(OUT) and os.path.isfile(stamp) and open(stamp).read() == h.hexdigest():
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ['g++', '-std=c++20', '-O1', '-g', '-fPIC', '-shared', '-pthread', '-DSTAT_EMU', '-I', HERE, '-x', 'c++', SRC,
           '-o', OUT]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('g++ failed:\n' + r.stderr)
    with open(stamp, 'w') as fh:
        fh.write(h.hexdigest())
    return OUT


if __name__ == '__main__':
    print(build(force=True))

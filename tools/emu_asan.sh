#!/bin/bash
# The backward pass (csrc/backward.cu) under the CPU launch emulation, compiled with AddressSanitizer: every kernel
# and the orchestration of stat_grad_shared, both STAT_BW_FAST modes, against the gradient oracle.  CPU only.
set -e
cd "$(dirname "$0")/.."
OUT=${TMPDIR:-/tmp}/stat_emu_asan
mkdir -p "$OUT"
g++ -std=c++20 -O1 -g -fPIC -shared -pthread -fsanitize=address -fno-omit-frame-pointer -DSTAT_EMU -I tests/emu \
    -x c++ video_description_with_spatial_temporal_attention_b200/csrc/backward.cu -o "$OUT/libstat_bw_emu.so"
cat > "$OUT/run.py" <<PY
import os, sys
sys.path.insert(0, os.getcwd())
from tests import test_backward_emu as t
from tests.emu import build_emu
from oracle import grad_oracle as go
build_emu.build = lambda force=False: "$OUT/libstat_bw_emu.so"
lib = t.load_emu()
for fast in (False, True):
    os.environ['STAT_BW_FAST'] = '1' if fast else '0'
    for gp in (False, True):
        o, params, batch = t._case(gp)
        g = t.run_emu(lib, o, params, batch, 0.7, 1e-4, flat=fast)
        t._compare(g, go.cost_and_grads(params, o, batch, alpha_c=0.7, decay_c=1e-4)[1])
        print('asan clean: fast=%s global_proj=%s' % (fast, gp), flush=True)
PY
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:verify_asan_link_order=0 python "$OUT/run.py"

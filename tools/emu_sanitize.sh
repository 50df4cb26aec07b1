#!/bin/bash
# The backward pass (csrc/backward.cu) under the CPU launch emulation (tests/emu), compiled with a sanitizer:
#   tools/emu_sanitize.sh address    out-of-bounds accesses of kernels and orchestration (compute-sanitizer memcheck's job)
#   tools/emu_sanitize.sh thread     data races between the emulated CUDA threads of a block, e.g. a missing
#                                    __syncthreads (racecheck's job); "selftest" as a second argument first proves
#                                    that an injected race (the barrier inside block_sum removed) IS reported
# Every kernel and the orchestration of stat_grad_shared, both STAT_BW_FAST modes, checked against the gradient
# oracle while the sanitizer watches.  CPU only; test infrastructure.
set -e
KIND=${1:-address}
cd "$(dirname "$0")/.."
OUT=${TMPDIR:-/tmp}/stat_emu_$KIND
mkdir -p "$OUT"
SRC=video_description_with_spatial_temporal_attention_b200/csrc/backward.cu
build() { g++ -std=c++20 -O1 -g -fPIC -shared -pthread -fsanitize=$KIND -fno-omit-frame-pointer -DSTAT_EMU -I tests/emu \
               -x c++ "$1" -o "$2"; }
cat > "$OUT/run.py" <<PY
import os, sys
sys.path.insert(0, os.getcwd())
from tests import test_backward_emu as t
from tests.emu import build_emu
from oracle import grad_oracle as go
build_emu.build = lambda force=False: sys.argv[1]
lib = t.load_emu()
for fast in (False, True):
    os.environ['STAT_BW_FAST'] = '1' if fast else '0'
    for gp in (False, True):
        o, params, batch = t._case(gp)
        g = t.run_emu(lib, o, params, batch, 0.7, 1e-4, flat=fast)
        t._compare(g, go.cost_and_grads(params, o, batch, alpha_c=0.7, decay_c=1e-4)[1])
        print('ran: fast=%s global_proj=%s' % (fast, gp), flush=True)
PY
RT=$(gcc -print-file-name=lib$([ "$KIND" = thread ] && echo tsan || echo asan).so)
export ASAN_OPTIONS=detect_leaks=0:verify_asan_link_order=0 TSAN_OPTIONS=halt_on_error=0
if [ "$KIND" = thread ] && [ "$2" = selftest ]; then
  python - "$SRC" "$OUT/racy.cu" <<'PY'
import sys
s = open(sys.argv[1]).read()
old = "  if (lane == 0) sh[w] = v;\n  __syncthreads();\n  float r = 0.f;"
assert old in s
open(sys.argv[2], 'w').write(s.replace(old, "  if (lane == 0) sh[w] = v;\n  float r = 0.f;", 1))
PY
  build "$OUT/racy.cu" "$OUT/libracy.so"
  n=$(LD_PRELOAD=$RT python "$OUT/run.py" "$OUT/libracy.so" 2>&1 | grep -c "ThreadSanitizer: data race" || true)
  echo "selftest: $n race reports with the barrier of block_sum removed (must be > 0)"
  [ "$n" -gt 0 ]
fi
build "$SRC" "$OUT/libstat_bw_emu.so"
LD_PRELOAD=$RT python "$OUT/run.py" "$OUT/libstat_bw_emu.so" > "$OUT/log.txt" 2>&1 || { tail -30 "$OUT/log.txt"; exit 1; }
grep "^ran:" "$OUT/log.txt"
n=$(grep -c "Sanitizer" "$OUT/log.txt" || true)
echo "$KIND sanitizer reports: $n"
[ "$n" -eq 0 ]

#!/bin/bash
# The caller-side elimination loop (tests/c/elim_loop.c) at K=1024 / L=1280: reference on the host CPU
# (scalar and AVX-512 builds) against liboblas_b200.so with one launch + synchronise per call (mode 0)
# and with the op queue (mode 2).  Numbers quoted in INTEGRATION.md.
set -e
cd "$(dirname "$0")/.."
gcc -std=c99 -O2 -D_DEFAULT_SOURCE -DOBLAS_B200_QUEUE -I include/compat tests/c/elim_loop.c -L oblas_b200 -loblas_b200 -Wl,-rpath,$PWD/oblas_b200 -o /tmp/elim_loop_b200
K=${1:-1024}; L=${2:-1280}
for exe in oracle/_ref/elim_loop_ref oracle/_ref/elim_loop_avx512; do
  [ -x $exe ] && { echo "== $exe"; $exe $K $L 2>&1; }
done
echo "== liboblas_b200.so, op queue (oblas_b200_set_async(2))"; /tmp/elim_loop_b200 $K $L 2 2>&1
echo "== liboblas_b200.so, op queue, second run"; /tmp/elim_loop_b200 $K $L 2 2>&1
if [ "${3:-}" = "sync" ]; then echo "== liboblas_b200.so, launch + synchronise per call"; timeout 300 /tmp/elim_loop_b200 $K $L 0 2>&1; fi

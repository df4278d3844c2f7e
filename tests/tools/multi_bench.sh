#!/bin/bash
# timing of several builds of the library on the same box: tests/tools/multi_bench.sh rounds lib1.so lib2.so ...
N=$1; shift
export MGCOD_BENCH_SKIP_SHAPE2=1 MGCOD_BENCH_SKIP_FILES=1
for i in $(seq $N); do
  for L in "$@"; do
    MGCOD_B200_LIB=$PWD/$L python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); p=d['phases_ms_per_step']; print('$L', round(d['ms_per_step'],2), ' '.join(f'{k}={v:.2f}' for k,v in p.items()))"
  done
done

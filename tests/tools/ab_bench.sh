#!/bin/bash
# A/B timing of two builds of the library on the same box: alternate runs, print ms_per_step and the phase split.
#   tests/tools/ab_bench.sh mgcod_b200/libmgcod_b200_A.so mgcod_b200/libmgcod_b200_B.so [rounds]
A=$1; B=$2; N=${3:-3}
export MGCOD_BENCH_SKIP_SHAPE2=1 MGCOD_BENCH_SKIP_FILES=1
for i in $(seq $N); do
  for L in "$A" "$B"; do
    MGCOD_B200_LIB=$PWD/$L python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); p=d['phases_ms_per_step']; print('$L', round(d['ms_per_step'],2), ' '.join(f'{k}={v:.2f}' for k,v in p.items()))"
  done
done

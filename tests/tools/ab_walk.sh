#!/bin/bash
# A/B of the two walk kernels in one build on the same box: MGCOD_B200_WALK=single (one position per lookup, k_walk)
# against the default pair walk (k_walk_pair).  Prints ms_per_step and the phase split of alternating runs.
#   tests/tools/ab_walk.sh [rounds]
N=${1:-2}
export MGCOD_BENCH_SKIP_SHAPE2=1 MGCOD_BENCH_SKIP_FILES=1
for i in $(seq $N); do
  for mode in single pair; do
    MGCOD_B200_WALK=$mode python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); p=d['phases_ms_per_step']; print('$mode', round(d['ms_per_step'],2), ' '.join(f'{k}={v:.2f}' for k,v in p.items()))"
  done
done

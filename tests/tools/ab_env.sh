#!/bin/bash
# A/B timing of one build under two settings of an environment variable (e.g. MGCOD_B200_CHAIN=warp against the default):
#   tests/tools/ab_env.sh MGCOD_B200_CHAIN warp [rounds]
VAR=$1; VAL=$2; N=${3:-2}
export MGCOD_BENCH_SKIP_SHAPE2=1 MGCOD_BENCH_SKIP_FILES=1
for i in $(seq $N); do
  for V in "" "$VAL"; do
    env $VAR=$V python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); p=d['phases_ms_per_step']; print('$VAR=$V', round(d['ms_per_step'],2), ' '.join(f'{k}={v:.2f}' for k,v in p.items()))"
  done
done

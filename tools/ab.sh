#!/bin/bash
# Same-box A/B of two library builds: tools/ab.sh <libA.so> <libB.so> [reps] [bench args...]
A=$1; B=$2; reps=${3:-3}; shift 3 2>/dev/null
for i in $(seq $reps); do
  for L in "$A" "$B"; do
    echo -n "$(basename $L): "
    LDIT_LIB_PATH=$(realpath $L) python bench.py --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], 'ms/step', d['value'], 'img/s; e2e', d['e2e']['ms_per_step'], 'ms; fc1', d['roofline']['kernel_ms'])"
  done
done

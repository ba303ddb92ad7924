#!/bin/bash
# A/B timing of one library under different environment settings (development tool):
#   tools/ab_env.sh c3 "4,2,2" build_ab/lib_x.so "" "SPUQ_NO_COGROUP=1"
cfg=$1; var=$2; lib=$3; shift 3
for round in 1 2; do
  for e in "$@"; do
    printf "[%s] r%d: " "$e" $round
    env $e SPUQ_LIB=$lib timeout 200 python tools/sweep.py --config $cfg --variants "$var" --reps 20 2>&1 | grep -o "[0-9.]* ms.*" | head -1
  done
done

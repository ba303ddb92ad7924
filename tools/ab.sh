#!/bin/bash
# A/B timing of library builds under build_ab/ on one box (development tool): tools/ab.sh c3 "4,2,2" base old ...
# build the variants first:  mkdir -p build_ab && make -C spuq_b200/csrc -B OUT=../../build_ab/lib_old.so EXTRA="-DSOME_SWITCH"
cfg=$1; var=$2; shift 2
for round in 1 2; do
  for n in "$@"; do
    printf "%s r%d: " $n $round
    SPUQ_LIB=build_ab/lib_$n.so timeout 200 python tools/sweep.py --config $cfg --variants "$var" --reps 20 2>&1 | grep -o "[0-9.]* ms.*" | head -1
  done
done

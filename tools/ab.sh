#!/bin/bash
# A/B of library variants on ONE box: tools/ab.sh "<command>" ab/libA.so ab/libB.so ...   (each variant twice, interleaved)
cmd="$1"; shift
for rep in 1 2; do
  for v in "$@"; do
    cp "$v" latticemontecarlo_b200/liblmc_b200.so
    echo "== $v (pass $rep)"
    bash -c "$cmd"
  done
done

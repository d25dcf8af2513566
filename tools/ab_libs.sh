#!/bin/bash
# A/B of library builds (make variant NAME=...): phase times [eval, qp update, ADMM, line search] at several batch sizes.
# usage: ab_libs.sh "<lib file names>" "<batches>"
for b in $2; do
  for lib in $1; do
    echo "$lib batch $b: $(python tools/prof_with_lib.py $lib --batch $b --steps 3 | tail -1)"
  done
done

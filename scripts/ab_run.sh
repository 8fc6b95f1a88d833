#!/bin/bash
# developer tool: A/B-test builds of the CUDA library (build/libs/*.so) on the same batch
# usage: scripts/ab_run.sh "<lib names>" "<prof_run args>" [env assignments...]
libs="$1"; shift; args="$1"; shift
for l in $libs; do
  echo "== lib=$l $args $*"
  env "$@" MPOA_LIB=$PWD/build/libs/$l.so MPOA_VERBOSE=1 timeout 300 python scripts/prof_run.py $args 2>&1 | tail -4 | cut -c1-330
done

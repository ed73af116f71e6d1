#!/bin/bash
# usage: gpu_ab_libs.sh <libA.so> <libB.so> <ab-spec>...   alternates A B A B in one call (box-to-box drift cancels)
O=gpurun_out/r02b
mkdir -p $O
A=$1; B=$2; shift 2
for rep in 1 2; do
  for L in $A $B; do
    echo "== $L (rep $rep)"
    RTW_AB_LIB=$L python tools/ab.py "$@" 2>$O/abl.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(' ', d['label'], d['size'], d['ms'], 'ms', d['sphere_tests'], 'sph/ray', d['node_tests'], 'nodes/ray')"
  done
done
tail -2 $O/abl.err

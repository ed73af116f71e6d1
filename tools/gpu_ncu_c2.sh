#!/bin/bash
# usage: gpu_ncu_c2.sh <tag> [ab-spec] [kernel regex]
O=${RTW_OUT:-gpurun_out/r02}
mkdir -p $O
SPEC=${2:-c2:1:3:1920x1080x64:1}
K=${3:-k_megakernel_flat}
python tools/ab.py "$SPEC" > $O/ncu_$1_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -o $O/prof_$1 -f python tools/ab.py "$SPEC" > $O/ncu_$1.log 2>&1
tail -2 $O/ncu_$1_plain.log; tail -3 $O/ncu_$1.log

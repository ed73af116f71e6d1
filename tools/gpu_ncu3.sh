#!/bin/bash
mkdir -p gpurun_out/r02
CMD="python tools/ab.py c4:8:500:1920x1080x16:2"
$CMD > gpurun_out/r02/ncu3_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_megakernel_bvh -s 2 -c 1 -o gpurun_out/r02/prof_bvh_c4 -f $CMD > gpurun_out/r02/ncu3.log 2>&1
tail -2 gpurun_out/r02/ncu3_plain.log; tail -3 gpurun_out/r02/ncu3.log

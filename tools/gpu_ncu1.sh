#!/bin/bash
mkdir -p gpurun_out/r02
CMD="python tools/ab.py s1_new:1:3:1920x1080x64:1"
$CMD > gpurun_out/r02/ncu1_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_megakernel_flat -s 2 -c 1 -o gpurun_out/r02/prof_flat2_s1 -f $CMD > gpurun_out/r02/ncu1.log 2>&1
tail -3 gpurun_out/r02/ncu1_plain.log; tail -5 gpurun_out/r02/ncu1.log
timeout 600 python -m pytest tests -m gpu -q -x --timeout 600 -k "production_primary or ties or edge_cases or trace_rays_random" > gpurun_out/r02/pytest_ab2.log 2>&1; tail -5 gpurun_out/r02/pytest_ab2.log

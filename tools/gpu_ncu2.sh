#!/bin/bash
mkdir -p gpurun_out/r02
CMD="python tools/ab.py cornell:6:3:600x600x100:1"
$CMD > gpurun_out/r02/ncu2_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_megakernel_flat -s 2 -c 1 -o gpurun_out/r02/prof_flat2_cornell -f $CMD > gpurun_out/r02/ncu2.log 2>&1
tail -2 gpurun_out/r02/ncu2_plain.log; tail -3 gpurun_out/r02/ncu2.log
timeout 600 python -m pytest tests -m gpu -q -x --timeout 600 -k "box_edges or ties" > gpurun_out/r02/pytest_ab4.log 2>&1; tail -8 gpurun_out/r02/pytest_ab4.log

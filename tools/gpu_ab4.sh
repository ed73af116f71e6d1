#!/bin/bash
mkdir -p gpurun_out/r02
python tools/ab.py 's1_k2:1:3:1920x1080x200:1' 's1_k3:1:3:1920x1080x200:1:RTW_FLAT_KERNEL=3' 's1_k1:1:3:1920x1080x200:1:RTW_FLAT_KERNEL=1' 's1_k2b:1:3:1920x1080x200:1' \
  'c3_k2:7:3:1920x1080x200:1' 'c3_k3:7:3:1920x1080x200:1:RTW_FLAT_KERNEL=3' 'c3_k1:7:3:1920x1080x200:1:RTW_FLAT_KERNEL=1' \
  'cornell_k2:6:3:600x600x200:1' 'cornell_k3:6:3:600x600x200:1:RTW_FLAT_KERNEL=3' > gpurun_out/r02/ab4.jsonl 2> gpurun_out/r02/ab4.err
cut -c1-150 gpurun_out/r02/ab4.jsonl; tail -3 gpurun_out/r02/ab4.err
python tools/bvh_sweep.py > gpurun_out/r02/bvh_sweep.jsonl 2> gpurun_out/r02/bvh_sweep.err; tail -2 gpurun_out/r02/bvh_sweep.err
timeout 300 python -m pytest tests -m gpu -q -x --timeout 600 -k "box_edges" > gpurun_out/r02/pytest_ab6.log 2>&1; tail -3 gpurun_out/r02/pytest_ab6.log

#!/bin/bash
mkdir -p gpurun_out/r02
python tools/ab.py 's1_old:1:3:1920x1080x200:1:RTW_FLAT_KERNEL=1' 's1_new:1:3:1920x1080x200:1' \
  'cornell_old:6:3:600x600x200:1:RTW_FLAT_KERNEL=1' 'cornell_new:6:3:600x600x200:1' \
  'c3_old:7:3:1920x1080x200:1:RTW_FLAT_KERNEL=1' 'c3_new:7:3:1920x1080x200:1' \
  's1small_old:1:3:600x400x50:1:RTW_FLAT_KERNEL=1' 's1small_new:1:3:600x400x50:1' > gpurun_out/r02/ab1.jsonl 2> gpurun_out/r02/ab1.err
cat gpurun_out/r02/ab1.jsonl; tail -3 gpurun_out/r02/ab1.err
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -k "image_parity or wavefront or unit or edge_cases or spp_split or flat_and_bvh or furnace or baseline_size or production_primary" > gpurun_out/r02/pytest_ab1.log 2>&1; tail -12 gpurun_out/r02/pytest_ab1.log

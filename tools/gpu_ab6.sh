#!/bin/bash
mkdir -p gpurun_out/r02
python tools/ab.py 'c4_exact:8:500:1920x1080x32:2:RTW_BVH_FMA=0' 'c4_fma:8:500:1920x1080x32:2' \
  'c2p_exact:1:11:1920x1080x64:2:RTW_BVH_FMA=0' 'c2p_fma:1:11:1920x1080x64:2' \
  's1_exact:1:3:1920x1080x100:2:RTW_BVH_FMA=0' 's1_fma:1:3:1920x1080x100:2' 'c4_exact2:8:500:1920x1080x32:2:RTW_BVH_FMA=0' 'c4_fma2:8:500:1920x1080x32:2' > gpurun_out/r02/ab6.jsonl 2> gpurun_out/r02/ab6.err
cut -c1-330 gpurun_out/r02/ab6.jsonl; tail -3 gpurun_out/r02/ab6.err
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -k "bvh or million or random_scene or primary or flat_and_bvh or ties or box_edges or negative or edge_cases" > gpurun_out/r02/pytest_ab8.log 2>&1; tail -8 gpurun_out/r02/pytest_ab8.log

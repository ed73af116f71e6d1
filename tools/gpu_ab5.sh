#!/bin/bash
mkdir -p gpurun_out/r02
python tools/ab.py 'c4_bin:8:500:1920x1080x32:2:RTW_BVH_WIDE=0' 'c4_wide:8:500:1920x1080x32:2' \
  'c2p_bin:1:11:1920x1080x64:2:RTW_BVH_WIDE=0' 'c2p_wide:1:11:1920x1080x64:2' \
  's1_bin:1:3:1920x1080x100:2:RTW_BVH_WIDE=0' 's1_wide:1:3:1920x1080x100:2' \
  'c4_sah_bin:8:500:1920x1080x32:2:RTW_BVH_WIDE=0,RTW_BVH_BUILDER=sah' 'c4_sah_wide:8:500:1920x1080x32:2:RTW_BVH_BUILDER=sah' > gpurun_out/r02/ab5.jsonl 2> gpurun_out/r02/ab5.err
cut -c1-330 gpurun_out/r02/ab5.jsonl; tail -3 gpurun_out/r02/ab5.err
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -k "bvh or million or random_scene or primary or image_parity or wavefront or flat_and_bvh or ties or box_edges or negative" > gpurun_out/r02/pytest_ab7.log 2>&1; tail -8 gpurun_out/r02/pytest_ab7.log

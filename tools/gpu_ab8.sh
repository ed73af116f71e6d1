#!/bin/bash
mkdir -p gpurun_out/r02
python tools/ab.py 's1_k1:1:3:1920x1080x500:1:RTW_FLAT_KERNEL=1' 's1_k3:1:3:1920x1080x500:1' 's1_k1b:1:3:1920x1080x500:1:RTW_FLAT_KERNEL=1' 's1_k3b:1:3:1920x1080x500:1' 's1_k2:1:3:1920x1080x500:1:RTW_FLAT_KERNEL=2' \
  'c3_k1:7:3:1920x1080x1000:1:RTW_FLAT_KERNEL=1' 'c3_k3:7:3:1920x1080x1000:1' 'cornell_k1_nobox:6:3:600x600x200:1:RTW_FLAT_KERNEL=1,RTW_BOX_PRIMS=0' 'cornell_k3:6:3:600x600x200:1' > gpurun_out/r02/ab8.jsonl 2> gpurun_out/r02/ab8.err
cut -c1-200 gpurun_out/r02/ab8.jsonl; tail -3 gpurun_out/r02/ab8.err

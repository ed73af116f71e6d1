#!/bin/bash
mkdir -p gpurun_out/r02
python tools/ab.py 's1_k3:1:3:1920x1080x500:1' 's1_k1:1:3:1920x1080x500:1:RTW_FLAT_KERNEL=1' 's1g5:1:5:1920x1080x100:1' > gpurun_out/r02/ab9.jsonl 2> gpurun_out/r02/ab9.err
cut -c1-300 gpurun_out/r02/ab9.jsonl; tail -3 gpurun_out/r02/ab9.err
timeout 600 python -m pytest tests -m gpu -q --timeout 600 -k "primary_hits or random_scene or image_parity or flat_and_bvh or negative" > gpurun_out/r02/pytest_ab10.log 2>&1; tail -4 gpurun_out/r02/pytest_ab10.log

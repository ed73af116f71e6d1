#!/bin/bash
mkdir -p gpurun_out/r02
python tools/ab.py 'cornell:6:3:600x600x200:1' 's1:1:3:1920x1080x200:1' 'c3:7:3:1920x1080x200:1' > gpurun_out/r02/ab3.jsonl 2> gpurun_out/r02/ab3.err
cat gpurun_out/r02/ab3.jsonl; tail -3 gpurun_out/r02/ab3.err
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -k "not million and not many_prims and not baseline_resolution" > gpurun_out/r02/pytest_ab5.log 2>&1; tail -8 gpurun_out/r02/pytest_ab5.log

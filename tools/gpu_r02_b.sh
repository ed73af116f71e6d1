#!/bin/bash
O=gpurun_out/r02b
mkdir -p $O
python tools/upload_trace.py 1 3 3 > $O/upload.log 2>&1
python tools/ab.py 'c2:1:3:1920x1080x500:1' 'c1b:1:3:600x400x50:0' 'cornell:6:3:600x600x200:0' 'c3:7:3:1920x1080x1000:0' 's2:2:3:1920x1080x200:0' 'g5:1:5:1920x1080x100:1' > $O/ab.jsonl 2> $O/ab.err
cut -c1-400 $O/ab.jsonl; tail -3 $O/ab.err; tail -8 $O/upload.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $O/pytest_gpu.log 2>&1; tail -5 $O/pytest_gpu.log

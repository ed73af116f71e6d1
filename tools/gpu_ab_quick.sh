#!/bin/bash
O=gpurun_out/r02b
mkdir -p $O
python tools/ab.py 'c2:1:3:1920x1080x500:1' 'cornell:6:3:600x600x200:0' 'c3:7:3:1920x1080x1000:0' "$@" > $O/abq.jsonl 2> $O/abq.err
cut -c1-330 $O/abq.jsonl; tail -3 $O/abq.err

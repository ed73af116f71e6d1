#!/bin/bash
O=gpurun_out/r02g
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "lbvh or device_built or many_prims or million" > $O/test.log 2>&1; echo "test rc $?"; tail -5 $O/test.log
RTW_UPLOAD_TRACE=1 python tools/ab.py 'c4_auto:8:500:1920x1080x32:2' 'c4_p0:8:500:1920x1080x32:2:RTW_LBVH_POW=0' 'c4_p4:8:500:1920x1080x32:2:RTW_LBVH_POW=4' \
  'f150_auto:8:150:1920x1080x32:2' 'f150_p0:8:150:1920x1080x32:2:RTW_LBVH_POW=0' > $O/ab2.jsonl 2> $O/ab2.err
python -c "
import sys, json
for l in open('$O/ab2.jsonl'):
    d = json.loads(l); print(d['label'], d['ms'], d['opts'], d['node_tests'], d['sphere_tests'])"
grep -n "candidates\|device build\|^upload\|total" $O/ab2.err | head -40

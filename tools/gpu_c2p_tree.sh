#!/bin/bash
O=gpurun_out/r02j
mkdir -p $O
S='1:11:1920x1080x50:2'
python tools/ab.py "sah:$S" "lbvh:$S:RTW_BVH_BUILDER=lbvh" "lbvh_p0:$S:RTW_BVH_BUILDER=lbvh,RTW_LBVH_POW=0" "sah_lm1:$S:RTW_BVH_LEAF_MAX=1" "sah_lm2:$S:RTW_BVH_LEAF_MAX=2" \
  "sah_lm8:$S:RTW_BVH_LEAF_MAX=8" "c2_sah:1:3:1920x1080x50:2" "c2_lbvh:1:3:1920x1080x50:2:RTW_BVH_BUILDER=lbvh" > $O/ab.jsonl 2> $O/ab.err
python -c "
import sys, json
for l in open('$O/ab.jsonl'):
    d = json.loads(l); print(d['label'], d['ms'], d['opts'], d['node_tests'], d['sphere_tests'])"
tail -3 $O/ab.err

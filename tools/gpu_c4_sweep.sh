#!/bin/bash
O=gpurun_out/r02h
mkdir -p $O
S='8:500:1920x1080x32:2'
python tools/ab.py "base:$S" "lm2:$S:RTW_BVH_LEAF_MAX=2" "lm3:$S:RTW_BVH_LEAF_MAX=3" "lm6:$S:RTW_BVH_LEAF_MAX=6" "lm8:$S:RTW_BVH_LEAF_MAX=8" \
  "t20:$S:RTW_BVH_THRESH=20" "t28:$S:RTW_BVH_THRESH=28" "t32:$S:RTW_BVH_THRESH=32" "s2:$S:RTW_BVH_STEPS=2" "s4:$S:RTW_BVH_STEPS=4" "s6:$S:RTW_BVH_STEPS=6" \
  "l2:$S:RTW_BVH_LEAF=2" "l6:$S:RTW_BVH_LEAF=6" "l8:$S:RTW_BVH_LEAF=8" "t28s4:$S:RTW_BVH_THRESH=28,RTW_BVH_STEPS=4" "t20s2:$S:RTW_BVH_THRESH=20,RTW_BVH_STEPS=2" \
  "p8:$S:RTW_LBVH_POW=8" "p3:$S:RTW_LBVH_POW=3" "base2:$S" > $O/ab.jsonl 2> $O/ab.err
python -c "
import sys, json
for l in open('$O/ab.jsonl'):
    d = json.loads(l); print(d['label'], d['ms'], d['opts'], d['node_tests'], d['sphere_tests'])"
tail -3 $O/ab.err

#!/bin/bash
O=gpurun_out/r02k
mkdir -p $O
python tools/ab.py "c4:8:500:1920x1080x32:2" "c4_lm1:8:500:1920x1080x32:2:RTW_BVH_LEAF_MAX=1" "cor:6:3:600x600x200:2" "cor_lm1:6:3:600x600x200:2:RTW_BVH_LEAF_MAX=1" \
  "c2:1:3:1920x1080x100:2" "c2_lm1:1:3:1920x1080x100:2:RTW_BVH_LEAF_MAX=1" "c2p:1:11:1920x1080x100:2" "c2p_lm1:1:11:1920x1080x100:2:RTW_BVH_LEAF_MAX=1" \
  "f150:8:150:1920x1080x64:2" "f150_lm1:8:150:1920x1080x64:2:RTW_BVH_LEAF_MAX=1" "s1g30:1:30:1920x1080x50:2" "s1g30_lm1:1:30:1920x1080x50:2:RTW_BVH_LEAF_MAX=1" > $O/ab.jsonl 2> $O/ab.err
python -c "
import sys, json
for l in open('$O/ab.jsonl'):
    d = json.loads(l); print(d['label'], d['prims'], d['ms'], d['opts'], d['node_tests'], d['sphere_tests'], d['rect_tests'])"
tail -3 $O/ab.err

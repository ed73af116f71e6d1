#!/bin/bash
# round 2: first entries of the BVH traversal stack in shared memory (RTW_BVH_KERNEL=4: 16 entries, 5: 24) — parity + A/B
O=gpurun_out/r02m
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "bvh_schedules" > $O/test.log 2>&1; echo "test rc $?"; tail -3 $O/test.log
python tools/ab.py 'c2p:1:11:1920x1080x100:2' 'c2p_s16:1:11:1920x1080x100:2:RTW_BVH_KERNEL=4' 'c2p_s24:1:11:1920x1080x100:2:RTW_BVH_KERNEL=5' \
  'c4:8:500:1920x1080x32:2' 'c4_s16:8:500:1920x1080x32:2:RTW_BVH_KERNEL=4' 'c4_s24:8:500:1920x1080x32:2:RTW_BVH_KERNEL=5' \
  'c2:1:3:1920x1080x100:2' 'c2_s16:1:3:1920x1080x100:2:RTW_BVH_KERNEL=4' 'cor:6:3:600x600x200:2' 'cor_s16:6:3:600x600x200:2:RTW_BVH_KERNEL=4' \
  's1g30:1:30:1920x1080x50:2' 's1g30_s16:1:30:1920x1080x50:2:RTW_BVH_KERNEL=4' \
  'c2p_b:1:11:1920x1080x100:2' 'c2p_s16_b:1:11:1920x1080x100:2:RTW_BVH_KERNEL=4' 'c4_b:8:500:1920x1080x32:2' 'c4_s16_b:8:500:1920x1080x32:2:RTW_BVH_KERNEL=4' > $O/ab.jsonl 2> $O/ab.err
python -c "
import sys, json
for l in open('$O/ab.jsonl'):
    d = json.loads(l); print(d['label'], d['prims'], d['ms'], d['opts'], d['node_tests'])"
tail -3 $O/ab.err

#!/bin/bash
# round 2: the BVH ray-queue schedule — parity test, racecheck on a small frame, A/B against the state machine
O=${RTW_OUT:-gpurun_out/r02c}
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "bvh_schedules" > $O/test.log 2>&1; echo "test rc $?"; tail -5 $O/test.log
timeout 600 python tools/ab.py \
  'c2p_sm:1:11:1920x1080x100:2' 'c2p_spec:1:11:1920x1080x100:2:RTW_BVH_KERNEL=3' \
  'c2p_spec_l8:1:11:1920x1080x100:2:RTW_BVH_KERNEL=3,RTW_BVH_LEAF=8' 'c2p_spec_l12:1:11:1920x1080x100:2:RTW_BVH_KERNEL=3,RTW_BVH_LEAF=12' \
  'c2p_spec_l16:1:11:1920x1080x100:2:RTW_BVH_KERNEL=3,RTW_BVH_LEAF=16' 'c2p_spec_l8s4:1:11:1920x1080x100:2:RTW_BVH_KERNEL=3,RTW_BVH_LEAF=8,RTW_BVH_STEPS=4' \
  'c2p_spec_l8s2:1:11:1920x1080x100:2:RTW_BVH_KERNEL=3,RTW_BVH_LEAF=8,RTW_BVH_STEPS=2' \
  'c4_sm:8:500:1920x1080x32:2' 'c4_spec:8:500:1920x1080x32:2:RTW_BVH_KERNEL=3' 'c4_spec_l8:8:500:1920x1080x32:2:RTW_BVH_KERNEL=3,RTW_BVH_LEAF=8' \
  'c4_spec_l12:8:500:1920x1080x32:2:RTW_BVH_KERNEL=3,RTW_BVH_LEAF=12' 'c4_spec_l16:8:500:1920x1080x32:2:RTW_BVH_KERNEL=3,RTW_BVH_LEAF=16' \
  'c4_spec_l8s4:8:500:1920x1080x32:2:RTW_BVH_KERNEL=3,RTW_BVH_LEAF=8,RTW_BVH_STEPS=4' \
  'c2_sm:1:3:1920x1080x100:2' 'c2_spec:1:3:1920x1080x100:2:RTW_BVH_KERNEL=3' \
  'cor_sm:6:3:600x600x200:2' 'cor_spec:6:3:600x600x200:2:RTW_BVH_KERNEL=3' \
  > $O/ab.jsonl 2> $O/ab.err
cut -c1-400 $O/ab.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['label'], d['ms'], d['opts'], d['node_tests'])
"; tail -3 $O/ab.err

#!/bin/bash
# round 2: stack entries carry their entry distance (RTW_BVH_TSTACK) + largest L1 split for the BVH kernels — parity tests, A/B of two builds
O=gpurun_out/r02e
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_unit.py -x -q -m gpu -k "bvh or BVH or ties or schedules or lbvh or million or large or wavefront or primary or probe" > $O/test.log 2>&1; echo "test rc $?"; tail -4 $O/test.log
SPECS="c2p:1:11:1920x1080x100:2 c4:8:500:1920x1080x32:2 c2:1:3:1920x1080x100:2 cor:6:3:600x600x200:2"
for rep in 1 2; do
  echo "== before (rep $rep)"; RTW_AB_LIB=_ab/lib_before_tstack.so python tools/ab.py $SPECS 2>>$O/ab.err | tee -a $O/ab_before.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(' ', d['label'], d['ms'], 'ms', d['node_tests'], 'nodes/ray')"
  echo "== tstack + max L1 (rep $rep)"; python tools/ab.py $SPECS 2>>$O/ab.err | tee -a $O/ab_after.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(' ', d['label'], d['ms'], 'ms', d['node_tests'], 'nodes/ray')"
done
echo "== tstack, default carveout"; python tools/ab.py 'c2p:1:11:1920x1080x100:2:RTW_BVH_CARVEOUT=-1' 'c4:8:500:1920x1080x32:2:RTW_BVH_CARVEOUT=-1' 2>>$O/ab.err | tee -a $O/ab_after_default_carveout.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(' ', d['label'], d['ms'], 'ms', d['node_tests'], 'nodes/ray')"
tail -3 $O/ab.err

for cfg in "20 2 8" "16 2 8" "24 2 8" "20 1 8" "20 3 8" "20 4 8" "20 2 4" "20 2 12" "20 2 16" "24 3 12" "16 3 8"; do
  set -- $cfg
  a=$(RTW_BVH_THRESH=$1 RTW_BVH_STEPS=$2 RTW_BVH_LEAF=$3 timeout 100 python tools/quick_bench.py 50 1 11 2 | tail -1 | awk '{print $3}')
  b=$(RTW_BVH_THRESH=$1 RTW_BVH_STEPS=$2 RTW_BVH_LEAF=$3 timeout 100 python tools/quick_bench.py 32 8 100 2 | tail -1 | awk '{print $3}')
  echo "thresh $1 steps $2 leaf $3 : grid11 $a ms, field100 $b ms"
done

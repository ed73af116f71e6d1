#!/bin/bash
# usage: gpu_multi.sh N   (N GPUs on one box)
N=$1
mkdir -p gpurun_out/r02
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests -m gpu -q --timeout 600 -k "render_multi" > gpurun_out/r02/pytest_multi_$N.log 2>&1; tail -3 gpurun_out/r02/pytest_multi_$N.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02/bench_$N.json 2> gpurun_out/r02/bench_$N.err; echo "rc $?" >> gpurun_out/r02/bench_$N.err
tail -3 gpurun_out/r02/bench_$N.err; python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r02/bench_$N.json") if l.startswith("{")][-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"])
    print(json.dumps(d["strong"], indent=1))
except Exception as e:
    print("no line", e)
PY

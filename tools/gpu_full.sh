#!/bin/bash
mkdir -p gpurun_out/r02
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02/build.log 2>&1
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02/pytest_gpu.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02/smoke.log 2>&1; echo "smoke rc $?" >> gpurun_out/r02/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02/bench.json 2> gpurun_out/r02/bench.err; echo "bench rc $?" >> gpurun_out/r02/bench.err
timeout 900 python tools/run_configs.py --no-cpu > gpurun_out/r02/configs.jsonl 2> gpurun_out/r02/configs.err
tail -4 gpurun_out/r02/pytest_gpu.log; tail -2 gpurun_out/r02/smoke.log; head -c 400 gpurun_out/r02/bench.json; tail -2 gpurun_out/r02/bench.err

#!/bin/bash
# round-2 GPU call 1: parity suite, bench line, per-config table
mkdir -p gpurun_out/r02
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/r02/pytest_gpu.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02/bench.json 2> gpurun_out/r02/bench.err; echo "bench rc $?" >> gpurun_out/r02/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02/bench_ref.json 2> gpurun_out/r02/bench_ref.err
timeout 900 python tools/run_configs.py --no-cpu > gpurun_out/r02/configs.jsonl 2> gpurun_out/r02/configs.err
tail -3 gpurun_out/r02/pytest_gpu.log; cat gpurun_out/r02/bench.json | head -c 1500

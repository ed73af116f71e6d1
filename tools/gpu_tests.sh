#!/bin/bash
mkdir -p gpurun_out/r02
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02/build.log 2>&1
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02/pytest_gpu.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02/pytest_gpu.log
tail -15 gpurun_out/r02/pytest_gpu.log

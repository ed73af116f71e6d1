#!/bin/bash
# final round-2 GPU call: parity suite, smoke, bench line (+ reference arm), per-config table, launch list
O=${RTW_OUT:-gpurun_out/r02}
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc $?" >> $O/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc $?" >> $O/bench.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 900 python tools/run_configs.py --no-cpu > $O/configs.jsonl 2> $O/configs.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 > $O/ncu_launches.log 2>&1
tail -3 $O/pytest_gpu.log; tail -2 $O/smoke.log; head -c 400 $O/bench.json; echo; tail -2 $O/bench.err; cut -c1-160 $O/configs.jsonl

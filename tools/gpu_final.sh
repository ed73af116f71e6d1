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
# ncu capture of the BVH kernel on C4 with the cost-chosen Morton grid (the plain run of the same command first)
python tools/ab.py 'c4:8:500:1920x1080x16:0' > $O/ncu_c4_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_megakernel_bvh -s 2 -c 1 -o $O/prof_bvh_c4 -f python tools/ab.py 'c4:8:500:1920x1080x16:0' > $O/ncu_c4.log 2>&1
tail -1 $O/ncu_c4_plain.log | cut -c1-300

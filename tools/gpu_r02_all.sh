#!/bin/bash
# round-2 GPU call: parity suite, bench line (+ reference arm), per-config table, launch list, ncu full captures
O=gpurun_out/r02
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc $?" >> $O/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc $?" >> $O/bench.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 900 python tools/run_configs.py --no-cpu > $O/configs.jsonl 2> $O/configs.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 > $O/ncu_launches.log 2>&1
cap() { # name kernel-regex ab-spec
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -o $O/prof_$1 -f python tools/ab.py "$3" > $O/ncu_$1.log 2>&1; tail -2 $O/ncu_$1.log
}
cap flat_c2 k_megakernel_flat "c2:1:3:1920x1080x64:1"
cap flat_cornell k_megakernel_flat "cornell:6:3:600x600x100:0"
cap flat_c3 k_megakernel_flat "c3:7:3:1920x1080x64:0"
cap bvh_c4 k_megakernel_bvh "c4:8:500:1920x1080x16:0"
cap bvh_c2p k_megakernel_bvh "c2p:1:11:1920x1080x32:0"
tail -3 $O/pytest_gpu.log; tail -2 $O/smoke.log; head -c 600 $O/bench.json; tail -2 $O/bench.err; cat $O/configs.jsonl | cut -c1-200

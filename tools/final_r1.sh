#!/bin/bash
# Round-end measurement pass on one B200 (outputs under gpurun_out/final/, copied into profiles/ afterwards).
set -u
O=gpurun_out/final; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.txt 2>&1
python bench.py > $O/bench_1gpu.log 2>&1; tail -1 $O/bench_1gpu.log > $O/bench_r1_1gpu.json
python bench.py --net full --no-cpu-baseline 2>/dev/null | tail -1 > $O/bench_full_r1_1gpu.json
python bench.py --workload train --no-cpu-baseline 2>/dev/null | tail -1 > $O/bench_train_r1_1gpu.json
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > $O/bench_reference_r1.json
VMB_PDL=0 timeout 300 python tools/step_profile.py 2>&1 | grep -v Warn | tail -30 > $O/step_profile.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
for s in pin96 pout96; do
  PB_SHAPE=$s VMB_PIXLIN_TC=2 timeout 200 ncu --set full --clock-control none --import-source on -k regex:pixlin_tc -s 2 -c 1 -o $O/ncu_pixlin_tc_$s -f python tools/pixlin_one.py > /dev/null 2>&1
done
timeout 200 python tools/pixlin_bench.py > $O/pixlin_bench.txt 2>&1
ls -la $O
cat $O/pytest_gpu.txt $O/smoke.txt; cut -c1-400 $O/bench_r1_1gpu.json; cut -c1-200 $O/bench_full_r1_1gpu.json; cut -c1-300 $O/bench_train_r1_1gpu.json; cut -c1-300 $O/bench_reference_r1.json

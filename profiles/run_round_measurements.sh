#!/bin/bash
# Round measurements on ONE B200 (run from the repo root under gpurun); writes into gpurun_out/, which
# `python profiles/make_summaries.py` then turns into the tracked summaries under profiles/.
#   gpurun --timeout 1500 -- 'bash profiles/run_round_measurements.sh > gpurun_out/r1_all.log 2>&1'
set -x
O=gpurun_out
python -m pytest tests -x -q -m gpu > $O/pytest_gpu_r1.log 2>&1; echo pytest_rc=$?; tail -2 $O/pytest_gpu_r1.log
python -c 'import __graft_entry__ as g; g.smoke()' > $O/smoke_r1.log 2>&1; echo smoke_rc=$?
python bench.py > $O/bench_r1.json 2> $O/bench_r1.err; echo bench_rc=$?
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_r1_reference.json 2>> $O/bench_r1.err; echo ref_rc=$?
# launch list and one full capture of the dominant kernel, each after the same command exited 0 without ncu
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r1.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_launches.log 2>&1; echo launches_rc=$?
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:simulate_kernel -s 1 -c 1 -f -o $O/prof_r1_arm_table python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_full.log 2>&1; echo full_rc=$?
for wl in se3_narrow_passage se2_arena arm_free arm_elbow se3_highres; do
  python bench.py --workload $wl --steps 5 --warmup 3 > $O/bench_r1_$wl.json 2>> $O/bench_r1.err
done
python bench.py --particles 1048576 --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_r1_arm_1m.json 2>> $O/bench_r1.err
python tests/gpu_perf_env_builder.py > $O/env_builder_r1.log 2>&1
python tests/gpu_perf_check_config.py > $O/check_config_perf.log 2>&1
python tests/gpu_big_parity.py > $O/big_parity.log 2>&1
python tests/gpu_perf.py > $O/gpu_perf_r1.log 2>&1
echo done

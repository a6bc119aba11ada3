#!/bin/bash
# Round-2 measurements on ONE B200 (run from the repo root under gpurun); writes into gpurun_out/ (scratch), from which
# profiles/ncu_summary.py and profiles/make_r2_summaries.py make the tracked summaries under profiles/.
#   gpurun --timeout 2400 -- 'bash profiles/run_round2_measurements.sh > gpurun_out/r2_all.log 2>&1'
set -x
O=gpurun_out
python -c 'import __graft_entry__ as g; g.smoke()' > $O/smoke_r2.log 2>&1; echo smoke_rc=$?
python bench.py > $O/bench_r2.json 2> $O/bench_r2.err; echo bench_rc=$?
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_r2_reference.json 2>> $O/bench_r2.err; echo ref_rc=$?
for wl in se3_narrow_passage se2_arena arm_free arm_elbow se3_highres; do
  python bench.py --workload $wl --steps 5 --warmup 3 --config5-particles 0 > $O/bench_r2_$wl.json 2>> $O/bench_r2.err
done
python bench.py --workload se2_arena --particles 128 --steps 20 --warmup 5 --config5-particles 0 > $O/bench_r2_se2_128.json 2>> $O/bench_r2.err
python bench.py --workload se3_narrow_passage --particles 16384 --steps 10 --warmup 3 --config5-particles 0 > $O/bench_r2_se3_16384.json 2>> $O/bench_r2.err
python bench.py --workload arm_table --particles 2368 --steps 10 --warmup 3 --config5-particles 0 --no-cpu-baseline > $O/bench_r2_arm_2368.json 2>> $O/bench_r2.err
# launch list and full captures of the dominant kernel, each after the same command exited 0 without ncu
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --config5-particles 0"
$B > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r2.csv $B > $O/ncu_launches_r2.log 2>&1; echo launches_rc=$?
$B > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:simulate_kernel -s 1 -c 1 -f -o $O/prof_r2_arm_table $B > $O/ncu_full_r2.log 2>&1; echo full_rc=$?
B2="$B --workload se3_narrow_passage --particles 16384"
$B2 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:simulate_kernel -s 1 -c 1 -f -o $O/prof_r2_se3_narrow $B2 > $O/ncu_full_r2_se3.log 2>&1; echo full2_rc=$?
B4="$B --workload se3_highres"
$B4 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:simulate_kernel -s 1 -c 1 -f -o $O/prof_r2_se3_highres $B4 > $O/ncu_full_r2_highres.log 2>&1; echo full4_rc=$?
# config 4 with the link-level culling of check_env on (default) and off: DRAM bytes and duration of the simulate kernel
for c in 1 0; do
  FKS_CULL=$c $B4 > /dev/null 2>&1 && FKS_CULL=$c ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors.sum --clock-control none -k regex:simulate_kernel -s 1 -c 1 --csv --log-file $O/dram_r2_highres_cull$c.csv $B4 > /dev/null 2>&1; echo cull${c}_rc=$?
done
python tests/gpu_perf.py > $O/gpu_perf_r2.log 2>&1
echo done

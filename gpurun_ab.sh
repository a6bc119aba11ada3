set -x
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r1.log 2>&1; echo pytest_rc=$?; tail -2 gpurun_out/pytest_gpu_r1.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r1.log 2>&1; echo smoke_rc=$?
python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; echo bench_rc=$?
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_reference.json 2>> gpurun_out/bench_r1.err; echo ref_rc=$?
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo launches_rc=$?
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain_bench2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:simulate_kernel -s 1 -c 1 -o gpurun_out/prof_r1_arm_table python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo full_rc=$?
for wl in se3_narrow_passage se2_arena arm_free arm_elbow se3_highres; do python bench.py --workload $wl --steps 5 --warmup 3 > gpurun_out/bench_r1_$wl.json 2>> gpurun_out/bench_r1.err; done
python bench.py --workload arm_table --particles 1048576 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r1_arm_1m.json 2>> gpurun_out/bench_r1.err

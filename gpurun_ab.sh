timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
export FKSGPU_LIBRARY=$PWD/fast_kinematic_simulator_b200/libfksgpu_timers.so
timeout 300 python tests/gpu_perf.py 2>&1 | grep -vE "^$|n=   2368|n=  16384|n=    128" | grep -A2 "n=  65536"

#!/bin/bash
# developer aid: builds fast_kinematic_simulator_b200/libfksgpu_timers.so with per-phase clock counters
set -e
cd "$(dirname "$0")/fast_kinematic_simulator_b200/csrc"
make -j4 > /dev/null
mkdir -p build_t
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fopenmp -I../../include -I. -DFKS_PHASE_TIMERS -c fks_kernels.cu -o build_t/fks_kernels.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fopenmp -o ../libfksgpu_timers.so build_t/fks_kernels.o build/fks_api.o build/fks_env_builder.o build/environment_builder.o -lgomp

# usage: ab.sh "<variants>" "<workload n> ..."   (run on the GPU box from the repo root)
for v in $1; do
  lib=$PWD/fast_kinematic_simulator_b200/libfksgpu${v:+_$v}.so
  [ "$v" = "base" ] && lib=$PWD/fast_kinematic_simulator_b200/libfksgpu.so
  echo "== $v"
  for wl in "${@:2}"; do FKSGPU_LIBRARY=$lib python tests/gpu_perf.py $wl 2>&1 | grep -v "^$"; done
done

# developer aid: A/B of the warps-per-CTA cap (variants built by ./build_variant.sh wNN -DFKS_MAX_WARPS=NN)
for w in $1; do
  lib=$PWD/fast_kinematic_simulator_b200/libfksgpu_w$w.so
  [ "$w" = "32" ] && lib=$PWD/fast_kinematic_simulator_b200/libfksgpu.so
  echo "== $w warps per CTA"
  for wl in "${@:2}"; do FKS_WARPS_PER_BLOCK=$w FKSGPU_LIBRARY=$lib python tests/gpu_perf.py $wl 2>&1 | grep -v "^$"; done
done

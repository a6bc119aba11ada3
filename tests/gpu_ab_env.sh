# developer aid: A/B over values of one environment variable:  gpu_ab_env.sh VAR "v1 v2 ..." "<workload n>" ...
var=$1; vals=$2; shift 2
for v in $vals; do
  echo "== $var=$v"
  for wl in "$@"; do env $var=$v python tests/gpu_perf.py $wl 2>&1 | grep -v "^$"; done
done

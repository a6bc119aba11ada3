for w in 16 8 4 2; do echo "== warps per block $w"; FKS_WARPS_PER_BLOCK=$w timeout 300 python tests/gpu_perf.py 2>&1 | grep -vE "^$|n=   2368|n=  16384|n=    128"; done

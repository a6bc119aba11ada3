for m in 0 1 2 0 1; do echo "== FKS_CULL=$m"; FKS_CULL=$m timeout 300 python tests/gpu_perf.py 2>&1 | grep -vE "^$|n=   2368|n=  16384|n=    128|phases|per call"; done

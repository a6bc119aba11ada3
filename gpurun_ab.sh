timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python tests/gpu_perf.py 2>&1 | grep -vE "^$|n=   2368|n=  16384|n=    128|phases|per call"

timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python tests/gpu_perf.py 2>&1 | grep -vE "^$|n=   2368|n=  16384|n=    128|phases|per call"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench arm_table', d['ms_per_step'], d['value'])"

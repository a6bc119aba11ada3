timeout 600 python tests/gpu_perf_check_config.py

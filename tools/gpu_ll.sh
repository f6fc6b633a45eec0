timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 300 python tools/bench_configs.py c1 c4 --iters 30 2>&1 | cut -c1-200
timeout 300 python tools/perf_probe.py 1000 1000 1000 5000 32 20 2>&1 | grep -E "run|phase"
timeout 300 python tools/perf_probe.py 4096 4096 128 8192 64 10 2>&1 | grep -E "run|phase"

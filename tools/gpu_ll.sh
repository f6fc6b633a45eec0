timeout 900 python -m pytest tests -m gpu -q -x -k "mttkrp or config2 or dimension_tree or golden or script6" 2>&1 | tail -5
timeout 300 python tools/perf_probe.py 4096 4096 256 8192 64 5 2>&1 | grep -E "mttkrp|run"
timeout 300 python tools/perf_probe.py 1000 1000 1000 5000 32 10 2>&1 | grep -E "mttkrp|run"
timeout 300 python tools/perf_probe.py 2048 2048 512 1024 8 5 2>&1 | grep -E "mttkrp|run"

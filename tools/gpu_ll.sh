timeout 600 python -m pytest tests -m gpu -q -k "tf32 or mttkrp_all or config2" 2>&1 | tail -12
PROBE_PREC=1 timeout 300 python tools/perf_probe.py 1000 1000 1000 5000 32 10 2>&1 | grep -E "mttkrp|run"
PROBE_PREC=1 timeout 300 python tools/perf_probe.py 4096 4096 256 8192 64 5 2>&1 | grep -E "mttkrp|run"
PROBE_PREC=1 timeout 300 python tools/perf_probe.py 2048 2048 512 1024 16 5 2>&1 | grep -E "mttkrp|run"

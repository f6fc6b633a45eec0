timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
timeout 300 python tools/bench_configs.py c1 c4 --iters 30 2>&1 | cut -c1-300

timeout 900 python -m pytest tests -m gpu -q -k "prox_tv or prox_matches or prox_large or mode1 or golden" 2>&1 | tail -4
timeout 900 python tools/bench_configs.py c5 --iters 10 2>&1 | tail -4 | cut -c1-190

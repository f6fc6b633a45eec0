timeout 900 python tools/bench_configs.py c5 --iters 10 2>&1 | tail -4 | cut -c1-330

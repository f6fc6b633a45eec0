timeout 600 python tools/create_probe.py 1000 1000 500 2>&1 | tail -4

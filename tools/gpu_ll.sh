timeout 900 python -m pytest tests -m gpu -q -k "release_all_device" 2>&1 | tail -8

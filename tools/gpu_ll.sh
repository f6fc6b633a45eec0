timeout 900 python -m pytest tests -m gpu -q -k "nvecs or init_front" 2>&1 | tail -5

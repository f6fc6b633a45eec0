timeout 900 python -m pytest tests -m gpu -q -k "linear_coupl or script14 or par2" 2>&1 | tail -25

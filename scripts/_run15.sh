python -m pytest tests -m gpu -q -x -k "lat_period or sym" 2>&1 | tail -15

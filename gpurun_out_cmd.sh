python -m pytest tests -x -q -m gpu -k "full_size" 2>&1 | tail -5
python -m pytest tests -x -q -m gpu --durations=3 2>&1 | tail -8

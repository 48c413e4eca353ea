python -m pytest tests -x -q -m gpu 2>&1 | tail -1
python scripts/batch_times.py 16 2>&1 | tail -4 | head -3

python -m pytest tests -x -q -m gpu 2>&1 | tail -12
python bench.py --workload 8o --batch 64 --steps 2 --no-cpu-baseline 2>&1 | tail -5 | cut -c1-600

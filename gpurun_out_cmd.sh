python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for r in 0 1; do echo "== rects $r"; TCC_B200_TILE_RECTS=$r python scripts/batch_times.py 16 2>&1 | tail -4 | head -3; done

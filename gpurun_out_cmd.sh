python -m pytest tests -x -q -m gpu 2>&1 | tail -2
TCC_B200_PANEL_CLASSES=3 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for g in 1 2 4; do echo "== classes $g"; TCC_B200_PANEL_CLASSES=$g python scripts/batch_times.py 16 2>&1 | tail -4 | head -3; done

python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for th in 128 256; do for te in 6144 13000; do echo "== threads $th tile $te"; TCC_B200_BATCH_THREADS=$th TCC_B200_TILE_ELEMS=$te python scripts/batch_times.py 16 2>&1 | tail -8; done; done

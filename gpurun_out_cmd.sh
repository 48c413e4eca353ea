python -m pytest tests -x -q -m gpu 2>&1 | tail -1
TCC_B200_TILE_ELEMS=400 TCC_B200_BATCH_THREADS=64 python -m pytest tests -x -q -m gpu -k "not full_size" 2>&1 | tail -1

echo "== T=32"; python scripts/sigma_time.py 16
for t in 16 64; do cp scripts/libtcc_sab$t.so tencirchem_b200/lib/libtcc_b200.so; echo "== T=$t"; python scripts/sigma_time.py 16; done

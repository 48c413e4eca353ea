python scripts/measure_fp64_peak.py | tee gpurun_out/fp64_peak.json

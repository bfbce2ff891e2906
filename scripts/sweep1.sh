set -x
python -m pytest tests/test_gpu_planner.py -x -q -m gpu -k "launch_geometry" 2>&1 | tail -3
for L in 0 2032 2064 2128 2256 1128 1064; do python scripts/perf_probe.py c3 2048 5000 1 $L 128; done
for L in 2064 2128; do python scripts/perf_probe.py c3 8192 5000 1 $L 128; done
for L in 0 2032 2064 2128 128 64; do python scripts/perf_probe.py c2 4096 5000 1 $L 128; done

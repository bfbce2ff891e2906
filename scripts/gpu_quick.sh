set -x
timeout 900 python -m pytest tests/test_gpu_planner.py -x -q -m gpu 2>&1 | tail -2
python scripts/perf_probe.py c3 16384 5000 1 0 64
python scripts/perf_probe.py c2 4096 5000 1 0 64
python scripts/perf_probe.py c3 2048 5000 1 0 64

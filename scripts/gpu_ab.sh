set -x
timeout 600 python -m pytest tests/test_gpu_planner.py tests/test_dropin.py tests/test_fnd_loop.py -x -q -m gpu 2>&1 | tail -3
python scripts/perf_probe.py c2 4096 5000 1 0 64 | cut -c1-160
python scripts/perf_probe.py c3 16384 5000 1 0 64 | cut -c1-160
python scripts/perf_probe.py c2 4096 5000 1 0 64 | cut -c1-160

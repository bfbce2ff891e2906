set -x
timeout 300 python -m pytest tests/test_gpu_planner.py -x -q -m gpu 2>&1 | tail -2
RRTFND_LIB=$PWD/rrt_star_fnd_algorithm_b200/_variants/lib_phase.so timeout 100 python scripts/phase_probe.py c2 4096 2>&1 | tail -14 | cut -c1-120
python scripts/perf_probe.py c2 4096 5000 1 0 64 | cut -c1-120
python scripts/perf_probe.py c3 16384 5000 1 0 64 | cut -c1-120

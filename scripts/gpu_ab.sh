set -x
timeout 300 python -m pytest tests/test_gpu_planner.py -x -q -m gpu 2>&1 | tail -2
for v in 2_2 8_4 24_8 64_8; do
  echo "== $v"; RRTFND_LIB=$PWD/rrt_star_fnd_algorithm_b200/_variants/lib_phase_$v.so timeout 100 python scripts/phase_probe.py c3 16384 2>&1 | grep -E "ext/s|annulus|nearest" | cut -c1-120
done
RRTFND_LIB=$PWD/rrt_star_fnd_algorithm_b200/_variants/lib_phase_24_8.so timeout 100 python scripts/phase_probe.py c2 4096 2>&1 | grep -E "ext/s|annulus" | cut -c1-120
python scripts/perf_probe.py c3 16384 5000 1 0 64 | cut -c1-120
python scripts/perf_probe.py c2 4096 5000 1 0 64 | cut -c1-120

# A/B of the compile-time variants prepared at the end of round 1 (none of them has run on a GPU yet). Build them first,
# in the build container:
#   scripts/build_variant.sh rank_near -DRRTFND_RANK_IN_NEAR=1
#   scripts/build_variant.sh split -DRRTFND_SPLIT_CANDIDATES=1
#   scripts/build_variant.sh smem_tables -DRRTFND_SMEM_TABLES=1
#   scripts/build_variant.sh lane_hint -DRRTFND_LANE_MIN_HINT=1
#   scripts/build_variant.sh libm_pow -DRRTFND_LIBM_POW=1
# then: gpurun --timeout 900 -- 'bash scripts/gpu_variants.sh'
set -x
V=$PWD/rrt_star_fnd_algorithm_b200/_variants
timeout 120 python -m pytest tests/test_libm_pow.py -x -q -m gpu 2>&1 | tail -2            # the -0 / subnormal fix of rrtfnd_pow_libm
python scripts/perf_probe.py c3 16384 5000 1 0 64 | cut -c1-140                           # baseline of this box
python scripts/perf_probe.py c2 4096 5000 1 0 64 | cut -c1-140
for v in rank_near split lane_hint smem_tables; do
  [ -f $V/$v.so ] || continue
  echo "== $v"
  RRTFND_LIB=$V/$v.so timeout 300 python -m pytest tests/test_gpu_planner.py tests/test_fnd_loop.py -x -q -m gpu 2>&1 | tail -2
  RRTFND_LIB=$V/$v.so python scripts/perf_probe.py c3 16384 5000 1 0 64 | cut -c1-140
  RRTFND_LIB=$V/$v.so python scripts/perf_probe.py c2 4096 5000 1 0 64 | cut -c1-140
done
if [ -f $V/libm_pow.so ]; then
  echo "== libm_pow (oracle in libm mode; the goldens come from the reference, i.e. from libm pow, either way)"
  RRTFND_LIB=$V/libm_pow.so RRTFND_ORACLE_LIBM_POW=1 timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
  RRTFND_LIB=$V/libm_pow.so python scripts/perf_probe.py c3 16384 5000 1 0 64 | cut -c1-140
  RRTFND_LIB=$V/libm_pow.so python scripts/perf_probe.py c2 4096 5000 1 0 64 | cut -c1-140
fi

# Round 2, GPU session 1: parity of the default build (new headline-path tests), A/B of the prepared variants, phase clocks.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_r2_s1.sh'
set -x
mkdir -p gpurun_out
V=$PWD/rrt_star_fnd_algorithm_b200/_variants
O=gpurun_out/s1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O.env.txt; nproc >> $O.env.txt
timeout 1200 python -m pytest tests -x -q -m gpu --durations=8 > $O.tests.txt 2>&1; tail -n 15 $O.tests.txt
probe() { python scripts/perf_probe.py "$@" 2>&1 | tail -n 1 | cut -c1-330; }
{
for v in default base fused rank_near split lane_hint smem_tables minw20; do
  L=$V/$v.so; [ $v = default ] && L=$PWD/rrt_star_fnd_algorithm_b200/librrtfnd.so
  [ -f $L ] || continue
  echo "== $v"
  RRTFND_LIB=$L timeout 300 python -m pytest tests/test_gpu_planner.py -x -q -m gpu 2>&1 | tail -n 2
  [ $v = fused ] && RRTFND_LIB=$L timeout 600 python -m pytest tests/test_headline_paths.py tests/test_fnd_loop.py tests/test_dropin.py -x -q -m gpu 2>&1 | tail -n 5
  RRTFND_LIB=$L probe c3 16384 5000 1 0 64
  RRTFND_LIB=$L probe c2 4096 5000 1 0 64
done
echo "== default, carve-out 100 % (L1 = 28 KB)"; RRTFND_CARVEOUT=100 probe c3 16384 5000 1 0 64
echo "== default, carve-out 75 %"; RRTFND_CARVEOUT=75 probe c3 16384 5000 1 0 64
echo "== default, 4-stage ring"; probe c3 16384 5000 1 4161 64
echo "== default, plain loads"; probe c3 16384 5000 1 4113 64
echo "== default, 2048 planners (strong-scaling shard of 8)"; probe c3 2048 5000 1 0 64
echo "== default, 4096 planners"; probe c3 4096 5000 1 0 64
echo "== phase clocks c3"; RRTFND_LIB=$V/phase.so python scripts/phase_probe.py c3 16384 5000 2>&1 | tail -n 14
echo "== phase clocks c2"; RRTFND_LIB=$V/phase.so python scripts/phase_probe.py c2 4096 5000 2>&1 | tail -n 14
echo "== libm_pow"
RRTFND_LIB=$V/libm_pow.so RRTFND_ORACLE_LIBM_POW=1 timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 4
RRTFND_LIB=$V/libm_pow.so probe c3 16384 5000 1 0 64
RRTFND_LIB=$V/libm_pow.so probe c2 4096 5000 1 0 64
} > $O.ab.txt 2>&1
cat $O.ab.txt

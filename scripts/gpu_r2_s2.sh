# Round 2, GPU session 2: occupancy x scan-path matrix, walk-record / parallel-choose-parent variants, phase clocks, libm pow.
#   gpurun --timeout 900 -- 'bash scripts/gpu_r2_s2.sh'      (every step has its own timeout)
mkdir -p gpurun_out
V=$PWD/rrt_star_fnd_algorithm_b200/_variants
D=$PWD/rrt_star_fnd_algorithm_b200/librrtfnd.so
O=gpurun_out/s2
probe() { timeout 150 python scripts/perf_probe.py "$@" 2>&1 | tail -n 1 | cut -c1-200; }
{
echo "== fp64 issue rates"; timeout 60 rrt_star_fnd_algorithm_b200/_tools/fp64_rate
echo "== full GPU suite, default build"; timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 6
echo "== phase clocks c3"; RRTFND_LIB=$V/phase.so timeout 150 python scripts/phase_probe.py c3 16384 5000 2>&1 | tail -n 14
echo "== default ring"; probe c3 16384 5000 1 0 64
echo "== default plain loads"; probe c3 16384 5000 1 4113 64
for v in minw20 minw24; do
  echo "== $v parity"; RRTFND_LIB=$V/$v.so timeout 200 python -m pytest tests/test_gpu_planner.py -x -q -m gpu 2>&1 | tail -n 1
  echo "== $v ring"; RRTFND_LIB=$V/$v.so probe c3 16384 5000 1 0 64
  echo "== $v plain loads"; RRTFND_LIB=$V/$v.so probe c3 16384 5000 1 4113 64
  echo "== $v c2 (plain loads by default)"; RRTFND_LIB=$V/$v.so probe c2 4096 5000 1 0 64
done
for v in walkrec cppar walkrec_cppar; do
  echo "== $v parity"; RRTFND_LIB=$V/$v.so timeout 300 python -m pytest tests/test_gpu_planner.py tests/test_headline_paths.py -x -q -m gpu 2>&1 | tail -n 1
  echo "== $v c3"; RRTFND_LIB=$V/$v.so probe c3 16384 5000 1 0 64
  echo "== $v c2"; RRTFND_LIB=$V/$v.so probe c2 4096 5000 1 0 64
done
echo "== default c3 2048 / 4096 planners (strong-scaling shards)"; probe c3 2048 5000 1 0 64; probe c3 4096 5000 1 0 64
echo "== libm_pow: whole GPU suite with the oracle in libm mode"
RRTFND_LIB=$V/libm_pow.so RRTFND_ORACLE_LIBM_POW=1 timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 4
RRTFND_LIB=$V/libm_pow.so probe c3 16384 5000 1 0 64
RRTFND_LIB=$V/libm_pow.so probe c2 4096 5000 1 0 64
echo "== phase clocks c2"; RRTFND_LIB=$V/phase.so timeout 150 python scripts/phase_probe.py c2 4096 5000 2>&1 | tail -n 14
} > $O.txt 2>&1
cat $O.txt

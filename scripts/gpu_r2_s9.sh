# Round 2, GPU sessions 9 and 10: a kernel change against the previous kernel (prev.so = HEAD) on the same box, parity first.
mkdir -p gpurun_out
O=gpurun_out/${SESSION:-s9}
V=$PWD/rrt_star_fnd_algorithm_b200/_variants
probe() { timeout 150 python scripts/perf_probe.py "$@" 2>&1 | tail -n 1 | cut -c1-200; }
{
echo "== parity: planner + headline + loop tests, default build"; timeout 600 python -m pytest tests/test_gpu_planner.py tests/test_headline_paths.py tests/test_fnd_loop.py tests/test_dropin.py -x -q -m gpu 2>&1 | tail -n 3
echo "== c3 16384: previous / new / previous / new"
RRTFND_LIB=$V/prev.so probe c3 16384 5000 1 0 64; probe c3 16384 5000 1 0 64; RRTFND_LIB=$V/prev.so probe c3 16384 5000 1 0 64; probe c3 16384 5000 1 0 64
echo "== c2 4096: previous / new"; RRTFND_LIB=$V/prev.so probe c2 4096 5000 1 0 64; probe c2 4096 5000 1 0 64
echo "== c3 2048: previous / new"; RRTFND_LIB=$V/prev.so probe c3 2048 5000 1 0 64; probe c3 2048 5000 1 0 64
for it in 600 5000; do echo "== phase clocks c3 x $it iterations (one slice), new kernel"; RRTFND_SLICES=1 RRTFND_LIB=$V/phase.so timeout 150 python scripts/phase_probe.py c3 16384 $it 2>&1 | tail -n 14; done
echo "== c4 100k"; timeout 300 python scripts/c4_run.py 100000 1 2>&1 | tail -n 1 | cut -c1-300
} > $O.txt 2>&1
cat $O.txt

# Round 2, GPU session 8: the adaptive, key-ordered annulus search against the previous kernel (same box), parity first.
mkdir -p gpurun_out
O=gpurun_out/s8
V=$PWD/rrt_star_fnd_algorithm_b200/_variants
probe() { timeout 150 python scripts/perf_probe.py "$@" 2>&1 | tail -n 1 | cut -c1-200; }
{
echo "== full GPU suite, default build"; timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 6
echo "== c3 16384: previous kernel / new / previous / new"
RRTFND_LIB=$V/oldann.so probe c3 16384 5000 1 0 64; probe c3 16384 5000 1 0 64; RRTFND_LIB=$V/oldann.so probe c3 16384 5000 1 0 64; probe c3 16384 5000 1 0 64
echo "== c2 4096: previous / new"; RRTFND_LIB=$V/oldann.so probe c2 4096 5000 1 0 64; probe c2 4096 5000 1 0 64
echo "== c3 2048: previous / new"; RRTFND_LIB=$V/oldann.so probe c3 2048 5000 1 0 64; probe c3 2048 5000 1 0 64
echo "== c3 16384 new, slice shapes 2 / 3"; RRTFND_SLICE_SHAPE=2 probe c3 16384 5000 1 0 64; RRTFND_SLICE_SHAPE=3 probe c3 16384 5000 1 0 64
for it in 600 5000; do echo "== phase clocks c3 x $it iterations (one slice), new kernel"; RRTFND_SLICES=1 RRTFND_LIB=$V/phase.so timeout 150 python scripts/phase_probe.py c3 16384 $it 2>&1 | tail -n 14; done
echo "== c4 100k (node grid uses the same search)"; timeout 300 python scripts/c4_run.py 100000 1 2>&1 | tail -n 1 | cut -c1-400
} > $O.txt 2>&1
cat $O.txt

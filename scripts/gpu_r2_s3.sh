# Round 2, GPU session 3: new default (parallel choose-parent, sliced overlapping launches), slice counts, bench lines.
mkdir -p gpurun_out
O=gpurun_out/s3
probe() { timeout 150 python scripts/perf_probe.py "$@" 2>&1 | tail -n 1 | cut -c1-200; }
{
echo "== full GPU suite, default build"; timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 6
for n in 1 8 4 15; do echo "== c3 16384, RRTFND_SLICES=$n"; RRTFND_SLICES=$n probe c3 16384 5000 1 0 64; done
echo "== c3 16384 default (8 slices by rule)"; probe c3 16384 5000 1 0 64
for n in 1 4 8; do echo "== c2 4096, RRTFND_SLICES=$n"; RRTFND_SLICES=$n probe c2 4096 5000 1 0 64; done
echo "== c3 2048 default / 4-stage ring"; probe c3 2048 5000 1 0 64; probe c3 2048 5000 1 4161 64
echo "== c3 4096 slices 1 / 4"; RRTFND_SLICES=1 probe c3 4096 5000 1 0 64; RRTFND_SLICES=4 probe c3 4096 5000 1 0 64
echo "== bench c3 (default command)"; timeout 400 python bench.py 2> $O.bench_c3.err | tee $O.bench_c3.json | cut -c1-1500
echo "== bench c5"; timeout 400 python bench.py --workload c5 2> $O.bench_c5.err | tee $O.bench_c5.json | cut -c1-1800
echo "== bench c2"; timeout 300 python bench.py --workload c2 --cpu-seconds 8 2> $O.bench_c2.err | tee $O.bench_c2.json | cut -c1-1200
} > $O.txt 2>&1
tail -c 3000 $O.bench_c3.err $O.bench_c5.err $O.bench_c2.err >> $O.txt 2>&1
cat $O.txt

# Round 2, GPU session 6: validation of the rectangle-free lean instances and the childless bitmap; bench lines.
mkdir -p gpurun_out
O=gpurun_out/s6
probe() { timeout 150 python scripts/perf_probe.py "$@" 2>&1 | tail -n 1 | cut -c1-200; }
{
echo "== full GPU suite, default build"; timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 6
echo "== c3 16384 / c2 4096 / c2r 4096 / c3 2048"; probe c3 16384 5000 1 0 64; probe c2 4096 5000 1 0 64; probe c2r 4096 5000 1 0 64; probe c3 2048 5000 1 0 64
echo "== bench c3 (default command)"; timeout 400 python bench.py 2> $O.bench_c3.err | tee $O.bench_c3.json | cut -c1-300
echo "== bench c5"; timeout 400 python bench.py --workload c5 2> $O.bench_c5.err | tee $O.bench_c5.json | cut -c1-300
echo "== bench c2"; timeout 300 python bench.py --workload c2 --cpu-seconds 8 2> $O.bench_c2.err | tee $O.bench_c2.json | cut -c1-300
echo "== bench c2r"; timeout 300 python bench.py --workload c2r --cpu-seconds 8 2> $O.bench_c2r.err | tee $O.bench_c2r.json | cut -c1-300
echo "== bench --impl reference (c3)"; timeout 400 python bench.py --impl reference --steps 2 --warmup 1 2> $O.bench_ref.err | tee $O.bench_ref.json | cut -c1-300
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
} > $O.txt 2>&1
tail -c 1500 $O.bench_c3.err $O.bench_c5.err $O.bench_c2.err $O.bench_c2r.err >> $O.txt 2>&1
cat $O.txt

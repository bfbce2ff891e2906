# Round 2, last GPU session: the shipped build (root filter, slope bound 1024) - whole GPU suite, smoke, the bench lines.
mkdir -p gpurun_out
O=gpurun_out/fin3
{
echo "== full GPU suite"; timeout 200 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 3
echo "== smoke"; timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
echo "== bench c3 (default command)"; timeout 200 python bench.py --cpu-seconds 8 2> $O.bench_c3.err | tee $O.bench_c3.json | cut -c1-200
echo "== bench c3, 2,048 planners per GPU (the N = 8 shard)"; timeout 100 python bench.py --planners 2048 --no-cpu-baseline 2> $O.bench_c3_2048.err | tee $O.bench_c3_2048.json | cut -c1-200
echo "== bench c5"; timeout 150 python bench.py --workload c5 --no-cpu-baseline 2> $O.bench_c5.err | tee $O.bench_c5.json | cut -c1-200
echo "== bench c2"; timeout 100 python bench.py --workload c2 --no-cpu-baseline 2> $O.bench_c2.err | tee $O.bench_c2.json | cut -c1-200
echo "== bench c2r"; timeout 100 python bench.py --workload c2r --no-cpu-baseline 2> $O.bench_c2r.err | tee $O.bench_c2r.json | cut -c1-200
} > $O.txt 2>&1
cat $O.txt

# Round 2, last GPU session: the shipped build (root filter on by default) - whole GPU suite, smoke, the bench lines, launch list.
mkdir -p gpurun_out
O=gpurun_out/fin2
{
echo "== full GPU suite"; timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 3
echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
echo "== bench c3 (default command)"; timeout 300 python bench.py 2> $O.bench_c3.err | tee $O.bench_c3.json | cut -c1-200
echo "== bench c5"; timeout 200 python bench.py --workload c5 --cpu-seconds 5 2> $O.bench_c5.err | tee $O.bench_c5.json | cut -c1-200
echo "== bench c2"; timeout 150 python bench.py --workload c2 --cpu-seconds 5 2> $O.bench_c2.err | tee $O.bench_c2.json | cut -c1-200
echo "== bench c3, 2,048 planners per GPU (the N = 8 shard)"; timeout 150 python bench.py --planners 2048 --no-cpu-baseline 2> $O.bench_c3_2048.err | tee $O.bench_c3_2048.json | cut -c1-200
echo "== bench c2r"; timeout 150 python bench.py --workload c2r --cpu-seconds 5 2> $O.bench_c2r.err | tee $O.bench_c2r.json | cut -c1-200
echo "== ncu launch list of the default bench command"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 40 --csv --log-file $O.launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O.ncu_launches.log 2>&1; tail -n 1 $O.ncu_launches.log | cut -c1-120
} > $O.txt 2>&1
tail -c 400 $O.bench_c3.err $O.bench_c5.err $O.bench_c2.err $O.bench_c2r.err $O.bench_c3_2048.err >> $O.txt 2>&1
cat $O.txt

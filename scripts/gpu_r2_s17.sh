# Round 2, GPU session 17: the root filter (csrc/root_filter.h: circle tests decided without the square root and the divisions
# when certain) against HEAD's kernel on one box, parity first. Variants: rf2 = root filter, rf2p = root + point filter,
# rf = root filter with the per-segment products hoisted. Ordered by priority: the call may be cut short by the round's budget.
mkdir -p gpurun_out
O=gpurun_out/s17
V=$PWD/rrt_star_fnd_algorithm_b200/_variants
probe() { timeout 120 python scripts/perf_probe.py "$@" 2>&1 | tail -n 1 | cut -c1-200; }
{
echo "== rf2p parity: collision ops (700k borderline segments) + planner goldens"; RRTFND_LIB=$V/rf2p.so timeout 300 python -m pytest tests/test_cull.py tests/test_gpu_planner.py -x -q -m gpu 2>&1 | tail -n 2
echo "== c3 16384: HEAD / rf2 / rf2p / HEAD"
probe c3 16384 5000 1 0 64; RRTFND_LIB=$V/rf2.so probe c3 16384 5000 1 0 64; RRTFND_LIB=$V/rf2p.so probe c3 16384 5000 1 0 64; probe c3 16384 5000 1 0 64
echo "== whole GPU suite, rf2p"; RRTFND_LIB=$V/rf2p.so timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 3
echo "== c2 4096: HEAD / rf2 / rf2p"; probe c2 4096 5000 1 0 64; RRTFND_LIB=$V/rf2.so probe c2 4096 5000 1 0 64; RRTFND_LIB=$V/rf2p.so probe c2 4096 5000 1 0 64
echo "== c3 2048: HEAD / rf2p"; probe c3 2048 5000 1 0 64; RRTFND_LIB=$V/rf2p.so probe c3 2048 5000 1 0 64
echo "== bench c3 (default command), rf2p"; RRTFND_LIB=$V/rf2p.so timeout 400 python bench.py 2> $O.bench_c3.err | tee $O.bench_c3.json | cut -c1-200
echo "== bench c2, rf2p"; RRTFND_LIB=$V/rf2p.so timeout 200 python bench.py --workload c2 --cpu-seconds 5 2> $O.bench_c2.err | tee $O.bench_c2.json | cut -c1-200
echo "== bench c5, rf2p"; RRTFND_LIB=$V/rf2p.so timeout 300 python bench.py --workload c5 --cpu-seconds 5 2> $O.bench_c5.err | tee $O.bench_c5.json | cut -c1-200
echo "== bench c3 at 2,048 planners, rf2p"; RRTFND_LIB=$V/rf2p.so timeout 200 python bench.py --planners 2048 --no-cpu-baseline 2> $O.bench_c3_2048.err | tee $O.bench_c3_2048.json | cut -c1-200
echo "== c3 16384: rf (hoisted); rf2p with culling cells 6 / 12 (auto = 8.94)"
RRTFND_LIB=$V/rf.so probe c3 16384 5000 1 0 64; RRTFND_LIB=$V/rf2p.so probe c3 16384 5000 1 0 64 6; RRTFND_LIB=$V/rf2p.so probe c3 16384 5000 1 0 64 12
echo "== phase clocks c3, rf2p"; RRTFND_SLICES=1 RRTFND_LIB=$V/phase_rf2p.so timeout 150 python scripts/phase_probe.py c3 16384 5000 2>&1 | tail -n 14
echo "== ncu launch list of the default bench command, rf2p"
RRTFND_LIB=$V/rf2p.so timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 40 --csv --log-file $O.launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O.ncu_launches.log 2>&1; tail -n 1 $O.ncu_launches.log | cut -c1-120
} > $O.txt 2>&1
tail -c 600 $O.bench_c3.err $O.bench_c2.err $O.bench_c5.err $O.bench_c3_2048.err >> $O.txt 2>&1
cat $O.txt

# Round 2, GPU session 4: validation of the new defaults, config 4 probes, ncu evidence, compute-sanitizer.
mkdir -p gpurun_out
O=gpurun_out/s4
V=$PWD/rrt_star_fnd_algorithm_b200/_variants
probe() { timeout 150 python scripts/perf_probe.py "$@" 2>&1 | tail -n 1 | cut -c1-200; }
{
echo "== full GPU suite, default build"; timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 4
echo "== c3 2048: 1 slice (default rule) / forced 4"; probe c3 2048 5000 1 0 64; RRTFND_SLICES=4 probe c3 2048 5000 1 0 64
echo "== c3 8192 default"; probe c3 8192 5000 1 0 64
echo "== c2 4096 default (8 slices)"; probe c2 4096 5000 1 0 64
echo "== c2r 4096 default (rectangles)"; probe c2r 4096 5000 1 0 64
echo "== c4: one tree, 30k / 100k nodes; 148 replicas at 100k"
timeout 200 python scripts/c4_run.py 30000 1 2>&1 | tail -n 1
timeout 300 python scripts/c4_run.py 100000 1 2>&1 | tail -n 1
timeout 300 python scripts/c4_run.py 100000 148 2>&1 | tail -n 1
echo "== c4 phase shares at 100k (instrumented build)"; RRTFND_LIB=$V/phase.so timeout 300 python scripts/c4_run.py 100000 1 2>&1 | tail -n 1
echo "== long parity: node grid + walk records, 200,000 nodes vs the oracle"
RRTFND_LONG_TESTS=1 timeout 900 python -m pytest tests/test_gpu_planner.py -x -q -m gpu -k large_tree --durations=1 2>&1 | tail -n 4
echo "== compute-sanitizer memcheck: two goldens + sliced launch + one replanning loop"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_planner.py tests/test_fnd_loop.py -x -q -m gpu -k "word_buffer_matches_reference_golden and (main_star or float_fn) or sliced and float_fn or tick_by_tick and device-loop+edges" 2>&1 | tail -n 12
echo "== compute-sanitizer racecheck: one golden"
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_planner.py -x -q -m gpu -k "word_buffer_matches_reference_golden and float_fn" 2>&1 | tail -n 12
echo "== ncu launch list of the default bench command"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file $O.launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O.ncu_launches.log 2>&1; tail -n 2 $O.ncu_launches.log | cut -c1-300
echo "== ncu --set full, plan_kernel, 2,048 planners (one slice)"
RRTFND_SLICES=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:plan_kernel -c 1 -o $O.prof_c3_2048 python bench.py --planners 2048 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O.ncu_full.log 2>&1; tail -n 2 $O.ncu_full.log | cut -c1-300
ncu -i $O.prof_c3_2048.ncu-rep --page raw --csv > $O.prof_c3_2048.raw.csv 2>/dev/null
} > $O.txt 2>&1
cat $O.txt

# Round 2, GPU session 5: config 4 to 10^6 nodes, rectangle-dispatch overhead, collision FP64 roofline, ncu on the default launch.
mkdir -p gpurun_out
O=gpurun_out/s5
V=$PWD/rrt_star_fnd_algorithm_b200/_variants
probe() { timeout 150 python scripts/perf_probe.py "$@" 2>&1 | tail -n 1 | cut -c1-200; }
{
echo "== parity: planner + node grid tests"; timeout 300 python -m pytest tests/test_gpu_planner.py -x -q -m gpu 2>&1 | tail -n 2
echo "== c3 16384: default / without the rectangle dispatch / default again"
probe c3 16384 5000 1 0 64; RRTFND_LIB=$V/norect.so probe c3 16384 5000 1 0 64; probe c3 16384 5000 1 0 64
echo "== c2 4096: default / without the rectangle dispatch"; probe c2 4096 5000 1 0 64; RRTFND_LIB=$V/norect.so probe c2 4096 5000 1 0 64
echo "== collision operators vs the FP64 issue rate"; timeout 200 python scripts/collision_roofline.py 2>&1 | tail -n 1
echo "== c4 at 100k with the large chain cache (+ phase shares)"
timeout 300 python scripts/c4_run.py 100000 1 2>&1 | tail -n 1
RRTFND_LIB=$V/phase.so timeout 300 python scripts/c4_run.py 100000 1 2>&1 | tail -n 1
echo "== c4 at 300k"; timeout 600 python scripts/c4_run.py 300000 1 2>&1 | tail -n 1
echo "== c4 at 1,000,000 nodes"; timeout 1700 python scripts/c4_run.py 1000000 1 2>&1 | tail -n 1
echo "== ncu (reduced sections) on the slices of the default 16,384-planner launch"
timeout 900 ncu --section SpeedOfLight --section Occupancy --section WarpStateStats --section MemoryWorkloadAnalysis --section LaunchStats --clock-control none -k regex:plan_kernel -c 8 --csv --page raw --log-file $O.ncu_default_slices.csv python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $O.ncu_default.log 2>&1; tail -n 1 $O.ncu_default.log | cut -c1-200
} > $O.txt 2>&1
cat $O.txt

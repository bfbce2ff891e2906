set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_planner.py -x -q -m gpu 2>&1 | tail -2
for W in c3 c2; do
P=2048; [ $W = c2 ] && P=4096
CMD="python bench.py --workload $W --planners $P --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain_$W.log 2>&1 && timeout 600 ncu --section SourceCounters --section WarpStateStats --section LaunchStats --section Occupancy --clock-control none --import-source on -k regex:plan_kernel -c 1 -f -o gpurun_out/src_$W $CMD > gpurun_out/ncu_src_$W.log 2>&1
tail -2 gpurun_out/ncu_src_$W.log; cut -c1-200 gpurun_out/plain_$W.log | tail -1
done

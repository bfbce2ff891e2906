set -x
mkdir -p gpurun_out
for W in c3 c2; do
P=2048; [ $W = c2 ] && P=4096
CMD="python bench.py --workload $W --planners $P --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain_$W.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:plan_kernel -c 1 -f -o gpurun_out/prof_$W $CMD > gpurun_out/ncu_$W.log 2>&1
tail -2 gpurun_out/ncu_$W.log; cut -c1-300 gpurun_out/plain_$W.log | tail -1
done

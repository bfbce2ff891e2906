set -x
mkdir -p gpurun_out
CMD="python bench.py --planners 4736 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain_final.log 2>&1 && timeout 1100 ncu --set full --clock-control none --import-source on -k regex:plan_kernel -c 1 -f -o gpurun_out/prof_final_c3 $CMD > gpurun_out/ncu_final.log 2>&1
tail -2 gpurun_out/ncu_final.log; cut -c1-250 gpurun_out/plain_final.log | tail -1

set -x
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 3000 gpurun_out/bench_c3.json
python bench.py --workload c2 --steps 3 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 1500 gpurun_out/bench_c2.json
CMD="python bench.py --planners 592 --iters 1000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:plan_kernel -c 1 -o gpurun_out/prof_c3 $CMD > gpurun_out/ncu2.log 2>&1
tail -5 gpurun_out/ncu2.log

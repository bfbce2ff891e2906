set -x
mkdir -p gpurun_out
( time python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err ) 2>&1 | grep real; tail -c 3200 gpurun_out/bench_c3.json
( time python bench.py --workload c2 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err ) 2>&1 | grep real; tail -c 600 gpurun_out/bench_c2.json
( time python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | grep real; tail -c 900 gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1; cut -c1-210 gpurun_out/launches_final.csv | tail -9
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain_final.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:plan_kernel -c 1 -f -o gpurun_out/prof_final_c3 $CMD > gpurun_out/ncu_final.log 2>&1
tail -2 gpurun_out/ncu_final.log

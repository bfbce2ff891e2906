set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time timeout 300 python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err ) 2>&1 | grep real; tail -c 3300 gpurun_out/bench_c3.json | cut -c1-900
( time timeout 200 python bench.py --workload c2 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err ) 2>&1 | grep real; tail -c 3000 gpurun_out/bench_c2.json | cut -c1-300
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1; cut -c1-230 gpurun_out/launches_final.csv | tail -2
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:plan_kernel -c 1 --csv --log-file gpurun_out/dram_c3_default.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_dram.log 2>&1; cut -c1-260 gpurun_out/dram_c3_default.csv | tail -3
CMD="python bench.py --planners 2048 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain_batch.log 2>&1 && timeout 400 ncu --set full --clock-control none --import-source on -k regex:plan_kernel -c 1 -f -o gpurun_out/prof_batch_c3 $CMD > gpurun_out/ncu_batch.log 2>&1
tail -2 gpurun_out/ncu_batch.log; cut -c1-200 gpurun_out/plain_batch.log | tail -1

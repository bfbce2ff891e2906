set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time timeout 300 python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err ) 2>&1 | grep real; tail -c 3300 gpurun_out/bench_c3.json | cut -c1-3300
( time timeout 200 python bench.py --workload c2 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err ) 2>&1 | grep real; tail -c 3000 gpurun_out/bench_c2.json | cut -c1-600
( time timeout 200 python bench.py --workload c5 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err ) 2>&1 | grep real; tail -c 3000 gpurun_out/bench_c5.json | cut -c1-400
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1; cut -c1-230 gpurun_out/launches_final.csv | tail -4
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:plan_kernel -c 1 --csv --log-file gpurun_out/dram_c3_default.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_dram.log 2>&1; cut -c1-260 gpurun_out/dram_c3_default.csv | tail -4
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c5.csv python bench.py --workload c5 --planners 512 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches_c5.log 2>&1; wc -l gpurun_out/launches_c5.csv

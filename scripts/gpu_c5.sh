set -x
mkdir -p gpurun_out
BENCH_DUMP_AFTER=100 timeout 120 python bench.py --workload c5 --steps 2 --warmup 1 --cpu-seconds 8 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
tail -c 3800 gpurun_out/bench_c5.json; grep -v "^\s*$" gpurun_out/bench_c5.err | tail -30 | cut -c1-300

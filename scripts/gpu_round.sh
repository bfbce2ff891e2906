# one GPU call: parity tests, then short benches, then ncu of the planner kernel (run from the repo root on the GPU box)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 2500 gpurun_out/bench_c3.json; tail -5 gpurun_out/bench_c3.err
timeout 600 python bench.py --workload c2 --steps 3 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 2500 gpurun_out/bench_c2.json; tail -5 gpurun_out/bench_c2.err
if [ -n "$NCU" ]; then
for W in c3 c2; do
CMD="python bench.py --workload $W --planners 2048 --iters 1000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain_$W.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:plan_kernel -c 1 -f -o gpurun_out/prof_$W $CMD > gpurun_out/ncu_$W.log 2>&1
tail -3 gpurun_out/ncu_$W.log
done
fi

set -x
mkdir -p gpurun_out
timeout 80 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --planners 4096 --steps 1 --warmup 3 > gpurun_out/bench_c3_n2.json 2> gpurun_out/bench_c3_n2.err; tail -c 1500 gpurun_out/bench_c3_n2.json; grep -v "^\s*$" gpurun_out/bench_c3_n2.err | grep -iv "OMP_NUM\|\*\*\*\*" | tail -5 | cut -c1-300

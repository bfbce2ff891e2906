set -x
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_c3_n2.json 2> gpurun_out/bench_c3_n2.err; tail -c 1800 gpurun_out/bench_c3_n2.json; grep -v "^\s*$" gpurun_out/bench_c3_n2.err | tail -5 | cut -c1-300

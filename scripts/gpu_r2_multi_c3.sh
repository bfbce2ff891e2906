# c3 only on N GPUs (final kernel): bash scripts/gpu_r2_multi_c3.sh <N>     (gpurun --gpus N)
N=$1
mkdir -p gpurun_out
O=gpurun_out/f$N
timeout 300 python -m pytest tests/test_headline_paths.py tests/test_gpu_planner.py -x -q -m gpu 2>&1 | tail -n 2
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 3 --warmup 3 > $O.c3.json 2> $O.c3.err; tail -c 400 $O.c3.err; grep "^{" $O.c3.json | cut -c1-300

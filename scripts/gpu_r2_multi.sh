# Round 2, multi-GPU session: bash scripts/gpu_r2_multi.sh <N>     (gpurun --gpus N)
N=$1
mkdir -p gpurun_out
O=gpurun_out/m$N
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
nvidia-smi --query-gpu=index,name --format=csv,noheader > $O.env.txt
run 29511 --steps 3 --warmup 3 > $O.c3.json 2> $O.c3.err; tail -c 600 $O.c3.err
run 29512 --workload c5 --steps 3 --warmup 3 > $O.c5.json 2> $O.c5.err; tail -c 600 $O.c5.err
run 29513 --workload c2 --steps 3 --warmup 3 --no-e2e > $O.c2.json 2> $O.c2.err; tail -c 300 $O.c2.err
run 29514 --impl reference --steps 1 --warmup 0 > $O.ref.json 2> $O.ref.err; tail -c 300 $O.ref.err
for f in c3 c5 c2 ref; do echo "== $f"; cut -c1-400 $O.$f.json; done

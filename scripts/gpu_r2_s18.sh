# Round 2, GPU session 18: the slope bound of the culled segments (RRTFND_SLOPE_MAX; 256 so far). Steeper segments test the whole
# table on one lane while the chunk's other lanes wait; a larger bound makes them rarer and the culling lists longer.
mkdir -p gpurun_out
O=gpurun_out/s18
probe() { timeout 120 python scripts/perf_probe.py "$@" 2>&1 | tail -n 1 | cut -c1-200; }
{
echo "== c3 16384: slope_max 256 / 1024 / 2048 / 4096 / 256"
for mu in 256 1024 2048 4096 256; do RRTFND_SLOPE_MAX=$mu probe c3 16384 5000 1 0 64; done
echo "== whole GPU suite at 4096"; RRTFND_SLOPE_MAX=4096 timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 3
echo "== c3 2048: 256 / 1024 / 2048 / 4096"; for mu in 256 1024 2048 4096; do RRTFND_SLOPE_MAX=$mu probe c3 2048 5000 1 0 64; done
echo "== c2 4096: 256 / 1024 / 2048 / 4096"; for mu in 256 1024 2048 4096; do RRTFND_SLOPE_MAX=$mu probe c2 4096 5000 1 0 64; done
echo "== collision operators + planner goldens at 1024"; RRTFND_SLOPE_MAX=1024 timeout 200 python -m pytest tests/test_cull.py tests/test_gpu_planner.py -x -q -m gpu 2>&1 | tail -n 2
echo "== collision operators + planner goldens at 2048"; RRTFND_SLOPE_MAX=2048 timeout 200 python -m pytest tests/test_cull.py tests/test_gpu_planner.py -x -q -m gpu 2>&1 | tail -n 2
echo "== bench c5 at 256 / 4096"; for mu in 256 4096; do RRTFND_SLOPE_MAX=$mu timeout 200 python bench.py --workload c5 --no-cpu-baseline 2>/dev/null | cut -c1-120; done
} > $O.txt 2>&1
cat $O.txt

set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_fnd.py tests/test_fnd_loop.py -x -q -m gpu 2>&1 | tail -3
timeout 100 python scripts/c5_debug.py 4096 2>&1 | grep -E "planned|tick (1|8|16|24):|done" | cut -c1-260
timeout 200 python bench.py --workload c5 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; tail -c 3000 gpurun_out/bench_c5.json | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['replanning'], d['e2e']['value'])"

set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -12
for MW in 16 24 32; do
L=$((1+16*2+256*MW))
python scripts/perf_probe.py c3 16384 5000 1 $L 64
python scripts/perf_probe.py c2 4736 5000 1 $L 64
done
python scripts/perf_probe.py c3 2048 5000 1 $((1+16*2+256*16)) 64
python scripts/perf_probe.py c2 16384 5000 1 $((2+16*2+256*32)) 64

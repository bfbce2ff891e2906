set -x
for MIN in 128 512; do for CELL in 0 2.5 10; do NODE_GRID=1 NG_MIN=$MIN NG_CELL=$CELL python scripts/perf_probe.py c3 2048 5000 1 0 64; done; done
python scripts/perf_probe.py c3 2048 5000 1 0 64

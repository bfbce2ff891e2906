set -x
python scripts/perf_probe.py c3 16384 5000 1 0 64
python scripts/perf_probe.py c2 4736 5000 1 0 64
python scripts/perf_probe.py c3 16384 5000 1 $((1+16*1+256*16)) 64

set -x
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
# launch = wpb + 16*stages + 256*minw
for L in 0 $((1+16*4+256*16)) $((1+16*2+256*16)) $((1+16*4+256*24)) $((1+16*2+256*32)); do python scripts/perf_probe.py c3 2048 5000 1 $L 64; done
for C in 4 6 12; do python scripts/perf_probe.py c3 2048 5000 1 0 64 $C; done
python scripts/perf_probe.py c3 2048 5000 0 0 64
for L in 0 $((1+16*2+256*16)) $((1+16*2+256*24)) $((1+16*2+256*32)) $((1+16*4+256*32)) $((2+16*2+256*32)); do python scripts/perf_probe.py c2 4096 5000 1 $L 64; done
for C in 3 5 10; do python scripts/perf_probe.py c2 4096 5000 1 0 64 $C; done
python scripts/perf_probe.py c2 4096 5000 0 0 64
python scripts/perf_probe.py c3 16384 5000 1 0 64
python scripts/perf_probe.py c2 16384 5000 1 0 64

# Round 2, closing measurement: per-phase clocks of the shipped build (instrumented variant of the same sources), one slice.
mkdir -p gpurun_out
{ echo "== phase clocks c3, shipped build"; RRTFND_SLICES=1 RRTFND_LIB=$PWD/rrt_star_fnd_algorithm_b200/_variants/phase.so timeout 80 python scripts/phase_probe.py c3 16384 5000 2>&1 | tail -n 14; } > gpurun_out/fin5.txt 2>&1
cat gpurun_out/fin5.txt

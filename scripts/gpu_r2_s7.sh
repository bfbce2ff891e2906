# Round 2, GPU session 7: where the first slice's time goes (phase clocks of short runs), slice shapes.
mkdir -p gpurun_out
O=gpurun_out/s7
V=$PWD/rrt_star_fnd_algorithm_b200/_variants
probe() { timeout 150 python scripts/perf_probe.py "$@" 2>&1 | tail -n 1 | cut -c1-200; }
{
for it in 600 1767 5000; do echo "== phase clocks c3, 16384 planners x $it iterations (one slice)"; RRTFND_SLICES=1 RRTFND_LIB=$V/phase.so timeout 150 python scripts/phase_probe.py c3 16384 $it 2>&1 | tail -n 14; done
echo "== c3 16384 default"; probe c3 16384 5000 1 0 64
for sh in 1 2 3; do echo "== c3 16384, RRTFND_SLICE_SHAPE=$sh"; RRTFND_SLICE_SHAPE=$sh probe c3 16384 5000 1 0 64; done
echo "== c3 16384, 12 slices shape 2"; RRTFND_SLICES=12 RRTFND_SLICE_SHAPE=2 probe c3 16384 5000 1 0 64
echo "== c3 16384 default again"; probe c3 16384 5000 1 0 64
echo "== c2 4096 default / shape 2"; probe c2 4096 5000 1 0 64; RRTFND_SLICE_SHAPE=2 probe c2 4096 5000 1 0 64
} > $O.txt 2>&1
cat $O.txt

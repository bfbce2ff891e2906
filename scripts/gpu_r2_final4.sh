# Round 2, closing GPU call: launch list (durations, DRAM bytes) of the shipped build's default bench command; two more slope bounds.
mkdir -p gpurun_out
O=gpurun_out/fin4
probe() { timeout 100 python scripts/perf_probe.py "$@" 2>&1 | tail -n 1 | cut -c1-200; }
{
echo "== ncu launch list of the default bench command"
timeout 150 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 30 --csv --log-file $O.launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O.ncu_launches.log 2>&1; tail -n 1 $O.ncu_launches.log | cut -c1-120
echo "== c3 16384: slope_max 512 / 1024 / 1536"; for mu in 512 1024 1536; do RRTFND_SLOPE_MAX=$mu probe c3 16384 5000 1 0 64; done
} > $O.txt 2>&1
cat $O.txt

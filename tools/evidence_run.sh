set -x
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2a_gputests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench_cfg2.json 2> gpurun_out/r2a_bench_cfg2.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches_cfg2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --launch-skip 165 -c 60 -f -o /tmp/r2a_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu -i /tmp/r2a_full.ncu-rep --page raw --csv > gpurun_out/r2a_full_raw.csv 2>/dev/null
for w in cfg3:1.0 cfg4:0.5 cfg1:1.0 cfg5cap:0.1; do
  n=${w%%:*}; s=${w##*:}
  timeout 600 python bench.py --workload $n --scale $s --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_$n.json 2> gpurun_out/r2a_bench_$n.err
done
ls -la gpurun_out

# Round-2 evidence, one GPU box: bench lines of every configuration, launch list and full ncu metrics of a warm cfg2 pass,
# the drop-in CLI end to end, the full-length cfg5 chain.  Outputs under gpurun_out/ (tools/make_profile_summary.py copies
# the kept ones to profiles/).
set -x
T=${1:-r2b}
python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_cfg2.json 2> gpurun_out/${T}_bench_cfg2.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${T}_launches_cfg2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
# the fourth resident pass (after three warm-up passes): skip = launches before the fourth k_validate, count = launches of one pass
read SKIP CNT <<< $(python - <<PY
import csv
rows=[r for r in csv.reader(l for l in open("gpurun_out/${T}_launches_cfg2.csv") if not l.startswith("=="))]
h=rows[0]; i=h.index("Kernel Name")
names=[r[i].split("(")[0] for r in rows[1:]]
starts=[k for k,n in enumerate(names) if n=="k_validate"]
print(starts[3], starts[4]-starts[3])
PY
)
ncu --set full --clock-control none --launch-skip $SKIP -c $CNT -f -o /tmp/${T}_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu -i /tmp/${T}_full.ncu-rep --page raw --csv > gpurun_out/${T}_full_raw.csv 2>/dev/null
for w in cfg3:1.0 cfg4:0.5 cfg1:1.0 cfg5cap:0.1 zipf2:0.2; do
  n=${w%%:*}; s=${w##*:}
  timeout 600 python bench.py --workload $n --scale $s --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/${T}_bench_$n.json 2> gpurun_out/${T}_bench_$n.err
done
timeout 600 python tools/run_cfg5_full.py 0.02 > gpurun_out/${T}_cfg5_full.json 2> gpurun_out/${T}_cfg5_full.err
timeout 900 python tools/cli_e2e.py --workload cfg2 --scale 1.0 --repeat 2 > gpurun_out/${T}_cli_e2e_cfg2.json 2> gpurun_out/${T}_cli_e2e_cfg2.err
ls -la gpurun_out

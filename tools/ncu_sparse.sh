set -x
T=${1:-r2d}
ncu --set full --clock-control none --import-source on --kernel-name regex:'k_cluster_sparse' -c 1 -f -o /tmp/${T}_cfg1 python bench.py --workload cfg1 --scale 1.0 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ncu -i /tmp/${T}_cfg1.ncu-rep --page raw --csv > gpurun_out/${T}_cfg1_raw.csv 2>/dev/null
ncu -i /tmp/${T}_cfg1.ncu-rep --page source --csv --kernel-name regex:k_cluster_sparse > gpurun_out/${T}_cfg1_sparse_src.csv 2>/dev/null

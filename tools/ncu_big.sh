# ncu captures of the large-chain cluster-editing kernels: cfg4 sample (k_cluster_big) and cfg1 (k_cluster_sparse)
set -x
T=${1:-r2c}
ncu --set full --clock-control none --import-source on --kernel-name regex:'k_cluster_big|k_pair_scores|k_read_rates|k_thread' -c 8 -f -o /tmp/${T}_cfg4 python bench.py --workload cfg4 --scale 0.05 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ncu -i /tmp/${T}_cfg4.ncu-rep --page raw --csv > gpurun_out/${T}_cfg4_raw.csv 2>/dev/null
ncu -i /tmp/${T}_cfg4.ncu-rep --page source --csv --kernel-name regex:k_cluster_big > gpurun_out/${T}_cfg4_big_src.csv 2>/dev/null
ncu --set full --clock-control none --import-source on --kernel-name regex:'k_cluster_sparse|k_sp_' -c 9 -f -o /tmp/${T}_cfg1 python bench.py --workload cfg1 --scale 1.0 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ncu -i /tmp/${T}_cfg1.ncu-rep --page raw --csv > gpurun_out/${T}_cfg1_raw.csv 2>/dev/null
ncu -i /tmp/${T}_cfg1.ncu-rep --page source --csv --kernel-name regex:k_cluster_sparse > gpurun_out/${T}_cfg1_sparse_src.csv 2>/dev/null
ls -la gpurun_out | grep ${T}

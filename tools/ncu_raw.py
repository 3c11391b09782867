"""Print selected metrics of an `ncu --page raw --csv` export, one column per captured launch."""
import csv
import sys
WANT = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'launch__shared_mem_per_block_dynamic', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
extra = sys.argv[2:]
for w in WANT + extra:
    idx = [i for i, h in enumerate(hdr) if h == w]
    if idx:
        print(w, units[idx[0]], [r[idx[0]] for r in data])
for i, h in enumerate(hdr):
    if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and 'not_issued' not in h:
        vals = [float(r[i]) for r in data]
        if max(vals) > 0.2:
            print(h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), [round(v, 2) for v in vals])

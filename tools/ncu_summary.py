#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export: the metrics the roofline discussion needs."""
import csv, sys
KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warp_latency_issue_stalled_barrier_per_warp_active.pct',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum', 'lts__t_sectors_op_red.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_red.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum',
        'sm__pipe_tensor_op_imma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'sm__cycles_elapsed.max']
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    print('==', path)
    units = dict(zip(hdr, rows[1])) if len(rows) > 1 else {}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print('kernel:', d.get('Kernel Name', '')[:110])
        for k in KEYS:
            if k in d and d[k] != '':
                print('   %-85s %s %s' % (k, d[k], units.get(k, '')))
        extra = [k for k in hdr if 'stalled' in k and 'per_issue_active' in k and k not in KEYS]
        for k in extra:
            try:
                if float(d[k]) > 0.3:
                    print('   %-85s %s' % (k, d[k]))
            except Exception:
                pass

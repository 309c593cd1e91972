"""Summarise an .ncu-rep (raw page) into the metrics that matter for the roofline."""
import csv, subprocess, sys, json
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'launch__grid_size', 'launch__block_size', 'sass__inst_executed_local_loads',
        'sass__inst_executed_local_stores', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block_dynamic',
        'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum']
def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {}
        for i, h in enumerate(hdr):
            if h in WANT or h == 'Kernel Name':
                d[h] = vals[i] + (' ' + units[i] if units[i] else '')
            if 'average_warps_issue_stalled' in h and h.endswith('per_issue_active.ratio'):
                try:
                    if float(vals[i]) > 0.2: d[h.replace('smsp__average_warps_issue_stalled_', 'stall_').replace('_per_issue_active.ratio', '')] = vals[i]
                except ValueError: pass
        res.append(d)
    print(json.dumps(res, indent=1))
main(sys.argv[1])

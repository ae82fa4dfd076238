"""Text summary of an .ncu-rep (ncu --set full): the metrics the roofline discussion in DESIGN.md uses, per kernel launch.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/rNN_x_ncu.txt"""
import csv
import io
import re
import subprocess
import sys

WANT = re.compile(
    r'^(gpu__time_duration\.sum|launch__(registers_per_thread|grid_size|block_size|occupancy_limit_\w+|waves_per_multiprocessor|shared_mem_per_block_\w+)'
    r'|sm__warps_active\.avg\.pct_of_peak_sustained_active|sm__throughput\.avg\.pct_of_peak_sustained_elapsed'
    r'|sm__inst_executed\.sum(\.per_cycle_elapsed)?|sm__issue_active\.avg\.pct_of_peak_sustained_elapsed|smsp__inst_executed\.sum'
    r'|sm__pipe_fp64_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)|sm__inst_executed_pipe_(fp64|fma|fmaheavy|alu|xu|lsu|uniform|tensor\w*)\.(avg|sum)\.pct_of_peak_sustained_active'
    r'|sm__pipe_tensor\w*cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)|sm__inst_executed_pipe_\w+\.sum'
    r'|dram__bytes_(read|write)\.sum|dram__throughput\.avg\.pct_of_peak_sustained_elapsed|lts__t_bytes\.sum|lts__t_sector_hit_rate\.pct'
    r'|l1tex__data_bank_conflicts_pipe_lsu\w*\.sum|l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum|smsp__inst_executed_op_(local|shared|global)_(ld|st)\.sum'
    r'|sass__inst_executed_local_(loads|stores)|smsp__thread_inst_executed_per_inst_executed\.ratio|smsp__average_warps?_issue_stalled_\w+_per_issue_active\.ratio'
    r'|smsp__average_warp_latency_issue_stalled_\w+\.ratio|smsp__warp_issue_stalled_\w+_per_warp_active\.pct|sm__cycles_elapsed\.max|sm__cycles_active\.avg'
    r'|smsp__sass_thread_inst_executed_op_(dfma|dmul|dadd|ffma|fmul|fadd)_pred_on\.sum|smsp__inst_executed_pipe_\w+\.sum)$')


def main():
    out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"== {d.get('Kernel Name', '?')}  grid {d.get('Grid Size', '?')} block {d.get('Block Size', '?')}")
        for k, u in zip(hdr, units):
            if WANT.match(k) and d.get(k, '') not in ('', 'n/a'):
                print(f'  {k} = {d[k]} {u}')


if __name__ == '__main__':
    main()

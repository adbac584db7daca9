"""Summarise an .ncu-rep: key metrics + top stalled source lines.  python tools/ncu_summary.py file.ncu-rep [ntop]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio']
for wn in want:
    idx = [i for i, h in enumerate(hdr) if h == wn]
    if idx:
        print('%-85s %s' % (wn, [r[idx[0]] for r in rows[1:]]))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = None
data = []
for r in rows:
    if len(r) > 3 and r[0] == 'Address':
        if h is not None:
            break
        h = r
        continue
    if h is not None and len(r) >= 8 and r[0].startswith('0x'):
        data.append(r)
if h:
    isamp = h.index('# Samples'); isrc = h.index('Source'); iex = h.index('Instructions Executed')
    tot = sum(int(r[isamp] or 0) for r in data)
    print('total samples', tot)
    for r in sorted(data, key=lambda r: -int(r[isamp] or 0))[:ntop]:
        print('%6s %5.1f%% ex=%-9s %s' % (r[isamp], 100.0 * int(r[isamp]) / max(1, tot), r[iex], r[isrc][:110]))

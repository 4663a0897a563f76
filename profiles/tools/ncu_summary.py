"""Summarise an .ncu-rep: key raw metrics + hottest SASS regions. Usage: ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__cycles_elapsed.avg', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__average_warp_latency_issue_stalled_long_scoreboard.pct',
        'lts__t_sector_hit_rate.pct', 'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_xu.sum', 'sm__inst_executed_pipe_alu.sum',
        'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_fp64.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed']
for i, h in enumerate(hdr):
    if h in want:
        print(f'{h:75s} {units[i]:12s} {vals[i]}')
stall = [(float(vals[i]), h) for i, h in enumerate(hdr) if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')]
for v, h in sorted(stall, reverse=True)[:8]:
    print(f'  stall {h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]:30s} {v:.2f}')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = rows[2:]
isrc = hdr.index('Source'); iex = hdr.index('Instructions Executed'); ith = hdr.index('Thread Instructions Executed'); ist = hdr.index('Warp Stall Sampling (All Samples)')
tot = sum(int(r[iex]) for r in data); tst = sum(int(r[ist]) for r in data)
print('total warp instr', tot, 'SASS lines', len(data), 'stall samples', tst)
step = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for a in range(0, len(data), step):
    ex = sum(int(r[iex]) for r in data[a:a+step]); th = sum(int(r[ith]) for r in data[a:a+step]); st = sum(int(r[ist]) for r in data[a:a+step])
    if ex > 0.01 * tot or st > 0.01 * tst:
        print(f'{a:5d} instr {100*ex/tot:5.1f}% thr/instr {th/max(ex,1):5.1f} stall {100*st/max(tst,1):5.1f}% | {data[a][isrc].strip()[:60]}')

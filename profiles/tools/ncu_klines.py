"""Per-source-line profile of one kernel of a multi-kernel report: ncu_klines.py rep kernel [top] [inst]"""
import csv, subprocess, io, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
byinst = len(sys.argv) > 4
out = subprocess.run(['ncu','-i',rep,'--kernel-name',kern,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(out)))
hdr=None; lines=[]
for r in rows:
    if r and r[0]=='Line No': hdr=r; continue
    if hdr is None or len(r)<10: continue
    if r[0]!='': lines.append(r)
iS=hdr.index('Warp Stall Sampling (All Samples)'); iE=hdr.index('Instructions Executed'); iT=hdr.index('Thread Instructions Executed')
def num(x):
    try: return float(x)
    except: return 0.0
tE=sum(num(r[iE]) for r in lines); tS=sum(num(r[iS]) for r in lines)
print('warp instr', tE, 'stall samples', tS)
key=(lambda r:num(r[iE])) if byinst else (lambda r:num(r[iS]))
for r in sorted(lines,key=key,reverse=True)[:top]:
    print(f'{r[0]:>5s} inst {100*num(r[iE])/tE:5.1f}% stall {100*num(r[iS])/tS:5.1f}% thr {num(r[iT])/max(num(r[iE]),1):5.1f} | {r[1].strip()[:110]}')
raw = subprocess.run(['ncu','-i',rep,'--kernel-name',kern,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw))); h=rows[0]; d=dict(zip(h,rows[2]))
for k in ['gpu__time_duration.sum','launch__registers_per_thread','launch__grid_size','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','lts__t_bytes.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__inst_executed.sum']:
    print(k, d.get(k))

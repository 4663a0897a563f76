"""Copy the measurements of the last gpurun calls from gpurun_out/ (scratch) into profiles/ (tracked):
bench lines, the ncu launch lists, and per-kernel summaries of the `ncu --set full` captures
(key raw metrics + hottest CUDA source lines).

    python profiles/tools/mk_profiles.py [tag]                      assemble profiles/ from gpurun_out/
    python profiles/tools/mk_profiles.py --summarise REP OUT [KEY]  (on the GPU box) one .ncu-rep -> text summary OUT;
                                                             KEY: record the dominant kernel's DRAM traffic in
                                                             gpurun_out/summ/traffic.json under that workload key
"""
import csv, io, json, shutil, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent.parent
G, P = ROOT / 'gpurun_out', ROOT / 'profiles'
SUMM = len(sys.argv) > 1 and sys.argv[1] == '--summarise'
tag = sys.argv[1] if (len(sys.argv) > 1 and not SUMM) else 'r01'
P.mkdir(exist_ok=True)

def num(x):
    try: return float(x)
    except Exception: return 0.0

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.sum', 'sm__cycles_elapsed.avg',
        'lts__t_sector_hit_rate.pct', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_tensor.sum',
        'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_xu.sum']

def kernels(rep):
    raw = subprocess.run(['ncu', '-i', str(rep), '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    return rows[0], rows[1], rows[2:]

def summarise(rep, out, traffic_key=None, traffic=None):
    hdr, units, rows = kernels(rep)
    lines = [f'# ncu --set full --clock-control none --import-source on; report {rep.name}', '']
    for r in rows:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        name = d['Kernel Name']
        lines.append(f'## {name}')
        for k in WANT:
            if k in d: lines.append(f'{k:75s} {u[k]:16s} {d[k]}')
        st = sorted(((num(d[h]), h) for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')), reverse=True)[:8]
        for v, h in st:
            lines.append(f'  stall {h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]:30s} {v:.2f}')
        if traffic_key and traffic is not None and ('k_rr_points' in name or 'k_ts_flux2' in name):
            mult = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}
            traffic[traffic_key] = num(d['dram__bytes_read.sum']) * mult.get(u['dram__bytes_read.sum'], 1) + \
                                   num(d['dram__bytes_write.sum']) * mult.get(u['dram__bytes_write.sum'], 1)
        # hottest source lines of this kernel
        short = name.split('(')[0].split('<')[0].split()[-1]
        src = subprocess.run(['ncu', '-i', str(rep), '--kernel-name', short, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                             capture_output=True, text=True).stdout
        h2 = None; ls = []
        for rr in csv.reader(io.StringIO(src)):
            if rr and rr[0] == 'Line No': h2 = rr; continue
            if h2 is None or len(rr) < 10 or rr[0] == '': continue
            ls.append(rr)
        if h2:
            iS = h2.index('Warp Stall Sampling (All Samples)'); iE = h2.index('Instructions Executed'); iT = h2.index('Thread Instructions Executed')
            tE = sum(num(x[iE]) for x in ls) or 1; tS = sum(num(x[iS]) for x in ls) or 1
            lines.append(f'  -- hottest source lines (of {tE:.0f} warp instructions, {tS:.0f} stall samples)')
            for x in sorted(ls, key=lambda x: num(x[iS]) / tS + num(x[iE]) / tE, reverse=True)[:18]:
                lines.append(f'  {x[0]:>5s} inst {100*num(x[iE])/tE:5.1f}% stall {100*num(x[iS])/tS:5.1f}% thr/inst {num(x[iT])/max(num(x[iE]),1):5.1f} | {x[1].strip()[:100]}')
        lines.append('')
    out.write_text('\n'.join(lines))

if SUMM:
    rep, out = Path(sys.argv[2]), Path(sys.argv[3])
    key = sys.argv[4] if len(sys.argv) > 4 else None
    out.parent.mkdir(parents=True, exist_ok=True)
    tf = out.parent / 'traffic.json'
    traffic = json.loads(tf.read_text()) if tf.exists() else {}
    summarise(rep, out, key, traffic)
    tf.write_text(json.dumps(traffic, indent=1) + '\n')
    sys.exit(0)

traffic = {}
tf = P / 'traffic.json'
if tf.exists():
    traffic = json.loads(tf.read_text())
for f in sorted(G.glob('bench_*.json')):
    lines = [l for l in f.read_text().splitlines() if l.startswith('{')]
    if lines:
        (P / f'{tag}_{f.name}').write_text(lines[-1] + '\n')
for f in sorted(G.glob('launches_*.csv')):
    shutil.copy(f, P / f'{tag}_{f.name}')
S = G / 'summ'
if S.exists():
    for f in sorted(S.glob('*.txt')):
        shutil.copy(f, P / f'{tag}_ncu_{f.name}')
    if (S / 'traffic.json').exists():
        traffic.update(json.loads((S / 'traffic.json').read_text()))
tf.write_text(json.dumps(traffic, indent=1) + '\n')
print(sorted(p.name for p in P.iterdir()))

"""Instruction mix by SASS opcode of an .ncu-rep (first kernel). Usage: ncu_opmix.py file.ncu-rep [topN]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
mix = collections.Counter(); thr = collections.Counter()
for r in rows:
    if r and 'Source' in r and 'Instructions Executed' in r:
        hdr = r; iSrc = hdr.index('Source'); iE = hdr.index('Instructions Executed'); iT = hdr.index('Thread Instructions Executed'); continue
    if hdr is None or len(r) <= iE: continue
    try: e = float(r[iE]); t = float(r[iT])
    except: continue
    toks = r[iSrc].split()
    if not toks: continue
    op = toks[1] if toks[0].startswith('@') and len(toks) > 1 else toks[0]
    op = op.split('.')[0]
    mix[op] += e; thr[op] += t
tot = sum(mix.values())
print('total', tot)
for op, e in mix.most_common(top):
    print(f'{op:12s} {100*e/tot:5.1f}%  thr/inst {thr[op]/max(e,1):5.1f}')

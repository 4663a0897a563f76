"""Per-CUDA-source-line summary of an .ncu-rep (needs -lineinfo). Usage: ncu_lines.py file.ncu-rep [topN]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; lines = []
for r in rows:
    if r and r[0] == 'Line No':
        hdr = r; continue
    if hdr is None or len(r) < 10: continue
    if r[0] != '':   # a CUDA source line aggregate
        lines.append(r)
iS = hdr.index('Warp Stall Sampling (All Samples)'); iE = hdr.index('Instructions Executed'); iT = hdr.index('Thread Instructions Executed')
def num(x):
    try: return float(x)
    except: return 0.0
tE = sum(num(r[iE]) for r in lines); tS = sum(num(r[iS]) for r in lines)
print(f'total warp instr {tE:.0f}  stall samples {tS:.0f}')
key = (lambda r: num(r[iE])) if (len(sys.argv) > 3 and sys.argv[3] == "inst") else (lambda r: num(r[iS]))
for r in sorted(lines, key=key, reverse=True)[:top]:
    e = num(r[iE]); s = num(r[iS]); t = num(r[iT])
    print(f'{r[0]:>5s} inst {100*e/tE:5.1f}% stall {100*s/tS:5.1f}% thr {t/max(e,1):5.1f} | {r[1].strip()[:100]}')

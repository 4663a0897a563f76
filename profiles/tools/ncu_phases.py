import csv, subprocess, sys, io
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr=None; lines=[]
for r in rows:
    if r and r[0]=='Line No': hdr=r; continue
    if hdr is None or len(r)<10: continue
    if r[0] != '': lines.append(r)
iS = hdr.index('Warp Stall Sampling (All Samples)'); iE = hdr.index('Instructions Executed')
def num(x):
    try: return float(x)
    except: return 0.0
tE=sum(num(r[iE]) for r in lines); tS=sum(num(r[iS]) for r in lines)
# cumulative by line number (file-agnostic: prints line->share so ranges can be read off)
acc={}
for r in lines:
    acc.setdefault(r[0],[0,0,r[1].strip()[:70]])
    acc[r[0]][0]+=num(r[iE]); acc[r[0]][1]+=num(r[iS])
for ln in sorted(acc, key=lambda x:int(x)):
    e,s,t=acc[ln]
    if e/tE>0.002 or s/tS>0.004: print(f'{ln:>5s} inst {100*e/tE:5.2f}% stall {100*s/tS:5.2f}% | {t}')

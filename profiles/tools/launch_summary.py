import csv, sys
from collections import OrderedDict
rows=list(csv.reader(open(sys.argv[1])))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hdr]; ik=h.index('Kernel Name'); iv=h.index('Metric Value')
d=OrderedDict()
for r in rows[hdr+1:]:
    if len(r)>iv:
        try: d.setdefault(r[ik].split('(')[0][:70],[]).append(float(r[iv].replace(',','')))
        except: pass
for k,v in d.items(): print(f'{k:70s} n={len(v):3d} mean={sum(v)/len(v)/1000:9.2f} us  last={v[-1]/1000:9.2f}')

ncu --metrics gpu__time_duration.sum,launch__grid_size,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -s 20 -c 15 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_c2.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); im=hdr.index('Metric Name'); iv=hdr.index('Metric Value'); iid=hdr.index('ID')
from collections import OrderedDict
d=OrderedDict()
for r in rows[1:]:
    d.setdefault((r[iid], r[ik][:40]), {})[r[im]]=r[iv]
for k,v in d.items(): print(k, v)
PY

mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
for w in c2 c5 c3; do
  timeout 300 python bench.py --steps 50 --warmup 3 --workload $w --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print(d['config']['workload'][:12], 'value %.3e'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'kernel_ms %.3f setup_ms %.3f'%(r['kernel_ms'], r['setup_ms']), 'frac', r.get('frac'))"
done 2>&1 | tee gpurun_out/try.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
grep -v "^==" gpurun_out/launches_c2.csv | python -c "
import csv,sys
for r in list(csv.reader(sys.stdin))[-4:]: print(r[4][:40], r[-1])"
ncu --set full --clock-control none --import-source on -k regex:'k_rr_orbit|k_rr_ldm|k_bin' -s 9 -c 3 -o gpurun_out/prof_setup_c2 -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full_s.log 2>&1

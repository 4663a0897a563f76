mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 --workload c4 --no-cpu > gpurun_out/bench_c4.json 2>gpurun_out/bench_c4.err; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_c4.json').read()); r=d['roofline']
print('c4 value %.4e ms/step %.4f kernel_ms %.4f setup_ms %.4f frac %s e2e %.4e' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['setup_ms'], r.get('frac'), d['e2e']['value']))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_c4.csv python bench.py --steps 2 --warmup 3 --no-cpu --workload c4 > gpurun_out/ncu_launch_c4.log 2>&1
python scratch/launch_summary.py gpurun_out/launches_c4.csv

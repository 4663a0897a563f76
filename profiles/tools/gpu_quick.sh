# parity tests + short bench lines for the given workloads (default c2 c3 c5)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
for w in ${@:-c2 c3 c5}; do
python bench.py --steps 20 --warmup 3 --workload $w --no-cpu > gpurun_out/bench_$w.json 2>gpurun_out/bench_$w.err; python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_$w.json') if l.startswith('{')][-1]); r=d['roofline']
print('$w value %.4e ms/step %.4f kernel_ms %.4f setup_ms %.4f frac %s e2e %.4e' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['setup_ms'], r.get('frac'), d['e2e']['value']))
PY
done

mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
for prec in fp64 fp32; do for w in c2 c5 c3 c1; do
  timeout 300 python bench.py --steps 50 --warmup 3 --workload $w --no-cpu --precision $prec 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$prec', d['config']['workload'][:12], 'value %.3e'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'kernel_ms %.4f setup_ms %.4f'%(r['kernel_ms'], r['setup_ms']), 'frac', r.get('frac'), 'e2e %.3e'%d['e2e']['value'])"
done; done 2>&1 | tee gpurun_out/try.log

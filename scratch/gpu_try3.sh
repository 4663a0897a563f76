mkdir -p gpurun_out
for dbg in 0 32 16 48; do
  echo "== PTB_POINTS_DEBUG=$dbg"
  for w in c2 c3; do
    PTB_POINTS_DEBUG=$dbg timeout 300 python bench.py --steps 50 --warmup 3 --workload $w --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print(d['config']['workload'][:12], 'value %.3e'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'kernel_ms %.3f setup_ms %.3f'%(r['kernel_ms'], r['setup_ms']), 'frac', r.get('frac'))"
  done
done 2>&1 | tee gpurun_out/try3.log

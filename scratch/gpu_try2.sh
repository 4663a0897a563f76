mkdir -p gpurun_out
for ipw in 1 4 8 16; do for mib in 16 32; do
  echo "== ITEMS_PER_WARP=$ipw MIN_ITEM_BLOCKS=$mib"
  for w in c2 c5 c3; do
    PTB_ITEMS_PER_WARP=$ipw PTB_MIN_ITEM_BLOCKS=$mib timeout 300 python bench.py --steps 30 --warmup 3 --workload $w --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print(d['config']['workload'][:12], 'value %.3e'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'kernel_ms %.3f setup_ms %.3f'%(r['kernel_ms'], r['setup_ms']), 'frac', r.get('frac'))"
  done
done; done 2>&1 | tee gpurun_out/try2.log

cd pytransit_b200/csrc
for mb in 2 3 4; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -cudart static -DPT_MINB=$mb -o ../libptb200.so ptb200.cu || exit 1
  cd ../..
  echo "== PT_MINB=$mb"
  for w in c2 c5 c3; do
    python bench.py --steps 30 --warmup 3 --workload $w --no-cpu | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print(d['config']['workload'][:12], 'value %.3e'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'kernel_ms %.3f setup_ms %.3f'%(r['kernel_ms'], r['setup_ms']), 'frac', r.get('frac'))"
  done
  cd pytransit_b200/csrc
done

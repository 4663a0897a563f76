mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
cd pytransit_b200/csrc
for mb in ${MBS:-2 3}; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -cudart static -DPT_MINB=$mb -o ../libptb200.so ptb200.cu || exit 1
  cd ../..
  echo "== PT_MINB=$mb"
  for w in c2 c5 c3; do
    timeout 300 python bench.py --steps 30 --warmup 3 --workload $w --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print(d['config']['workload'][:12], 'value %.3e'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'kernel_ms %.3f setup_ms %.3f'%(r['kernel_ms'], r['setup_ms']), 'frac', r.get('frac'))"
  done
  cd pytransit_b200/csrc
done 2>&1 | tee ../../gpurun_out/try.log

mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
for w in c2 c5 c3; do
  timeout 300 python bench.py --steps 50 --warmup 3 --workload $w --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print(d['config']['workload'][:12], 'value %.3e'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'kernel_ms %.3f setup_ms %.3f'%(r['kernel_ms'], r['setup_ms']), 'frac', r.get('frac'))"
done 2>&1 | tee gpurun_out/try.log
timeout 600 python bench.py --steps 5 --warmup 3 --workload c4 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -3 gpurun_out/bench_c4.err; cat gpurun_out/bench_c4.json
timeout 300 python bench.py --steps 200 --warmup 5 --workload c1 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; tail -3 gpurun_out/bench_c1.err; cat gpurun_out/bench_c1.json
ncu --set full --clock-control none --import-source on -k regex:'k_ts_|k_ldtk' -s 4 -c 4 -o gpurun_out/prof_ts_c4 -f python bench.py --steps 1 --warmup 1 --no-cpu --workload c4 > gpurun_out/ncu_full4.log 2>&1
tail -2 gpurun_out/ncu_full4.log

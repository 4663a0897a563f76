mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --workload c4 --no-cpu > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -3 gpurun_out/bench_c4.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_c4.json').read()); r=d['roofline']
print('value %.3e'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'kernel_ms %.4f setup_ms %.4f'%(r['kernel_ms'], r['setup_ms']), 'frac', r.get('frac'), 'e2e', d['e2e'])"
ncu --set full --clock-control none --import-source on -k regex:'k_ts_|k_ldtk' -s 5 -c 5 -o gpurun_out/prof_ts_c4 -f python bench.py --steps 1 --warmup 1 --no-cpu --workload c4 > gpurun_out/ncu_full4.log 2>&1
tail -2 gpurun_out/ncu_full4.log

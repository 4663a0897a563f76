mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
python bench.py --steps 50 --warmup 3 --workload c5 --no-cpu > gpurun_out/bench_c5.json
python bench.py --steps 10 --warmup 3 --workload c3 --no-cpu > gpurun_out/bench_c3.json
python bench.py --steps 10 --warmup 3 --workload c4 --no-cpu > gpurun_out/bench_c4.json
python bench.py --steps 200 --warmup 3 --workload c1 --no-cpu > gpurun_out/bench_c1.json
python bench.py --steps 100 --warmup 3 --precision fp32 --no-cpu > gpurun_out/bench_c2_fp32.json
python bench.py --steps 100 --warmup 3 --host-result copy --no-cpu > gpurun_out/bench_c2_copy.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json
ncu --set full --clock-control none --import-source on -k regex:k_host_delta -s 3 -c 1 -o gpurun_out/prof_delta_c2 -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_delta_c2.log 2>&1
mkdir -p gpurun_out/summ; python profiles/tools/mk_profiles.py --summarise gpurun_out/prof_delta_c2.ncu-rep gpurun_out/summ/delta_c2.txt; rm -f gpurun_out/prof_delta_c2.ncu-rep
python - <<'PY'
import json
for w in ('c2','c2_fp32','c2_copy','c3','c4','c5','c1','ref'):
    d=json.loads([l for l in open(f'gpurun_out/bench_{w}.json') if l.startswith('{')][-1]); e=d.get('e2e') or {}
    print(w, 'value %.4e ms/step %.4f e2e %.4e d2h %s' % (d['value'], d['ms_per_step'], e.get('value') or 0, e.get('d2h_bytes_per_step')))
PY

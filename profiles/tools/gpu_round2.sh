# Round-2 measurement pass (1 GPU): parity tests, bench lines for every config, launch lists and ncu --set full captures
# summarised on the box (only the C3 report travels back).   bash profiles/tools/gpu_round2.sh
mkdir -p gpurun_out/summ
rm -f gpurun_out/*.ncu-rep gpurun_out/summ/* gpurun_out/bench_*.json gpurun_out/launches_*.csv
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -2 gpurun_out/bench_c2.err
python bench.py --steps 50 --warmup 3 --workload c5 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
python bench.py --steps 10 --warmup 3 --workload c3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
python bench.py --steps 10 --warmup 3 --workload c4 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
python bench.py --steps 10 --warmup 3 --workload c4 --precision fp32 --no-cpu > gpurun_out/bench_c4_fp32.json 2> gpurun_out/bench_c4_fp32.err
python bench.py --steps 200 --warmup 3 --workload c1 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
python bench.py --steps 100 --warmup 3 --precision fp32 --no-cpu > gpurun_out/bench_c2_fp32.json 2> gpurun_out/bench_c2_fp32.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
for w in c2 c3 c4 c5; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/launches_$w.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-numba --no-counters --no-collective --workload $w > gpurun_out/ncu_launch_$w.log 2>&1
done
cap() {  # name regex skip count workload traffic_key keep
ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o gpurun_out/prof_$1 -f python bench.py --steps 2 --warmup 3 --no-cpu --no-numba --no-counters --no-collective --workload $5 > gpurun_out/ncu_$1.log 2>&1
python profiles/tools/mk_profiles.py --summarise gpurun_out/prof_$1.ncu-rep gpurun_out/summ/$1.txt $6
if [ "$7" != keep ]; then rm -f gpurun_out/prof_$1.ncu-rep; fi
}
cap points_c2 k_rr_points 3 1 c2 c2 drop
cap points_c3 k_rr_points 1 1 c3 c3 keep
cap points_c5 k_rr_points 3 1 c5 c5 drop
cap setup_c2 'k_rr_orbit|k_rr_ldm|k_bin' 9 3 c2 "" drop
cap ts_c4 'k_ts_|k_ldtk' 5 5 c4 c4 drop
cap delta_c2 k_host_delta 3 1 c2 "" drop
python - <<'PY'
import json
for w in ('c2','c2_fp32','c3','c4','c4_fp32','c5','c1','ref'):
    try:
        d=json.loads([l for l in open(f'gpurun_out/bench_{w}.json') if l.startswith('{')][-1]); e=d.get('e2e') or {}; r=d.get('roofline') or {}
        print(w, 'value %.4e ms/step %.4f kernel_ms %s frac %s e2e %.4e' % (d['value'], d['ms_per_step'], r.get('kernel_ms'), r.get('frac'), e.get('value') or 0))
    except Exception as ex:
        print(w, 'FAILED', ex)
PY
ls gpurun_out/summ; du -sh gpurun_out

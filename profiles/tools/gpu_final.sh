# Full measurement pass: parity tests, bench lines for every config, launch lists and ncu --set full captures
# (summarised on the box: the .ncu-rep files are too big to travel back, only two are kept).
mkdir -p gpurun_out/summ
rm -f gpurun_out/*.ncu-rep gpurun_out/summ/*
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -3 gpurun_out/bench_c2.err; cut -c1-400 gpurun_out/bench_c2.json
python bench.py --steps 50 --warmup 3 --workload c5 --no-cpu > gpurun_out/bench_c5.json
python bench.py --steps 10 --warmup 3 --workload c3 --no-cpu > gpurun_out/bench_c3.json
python bench.py --steps 10 --warmup 3 --workload c4 --no-cpu > gpurun_out/bench_c4.json
python bench.py --steps 200 --warmup 3 --workload c1 --no-cpu > gpurun_out/bench_c1.json
python bench.py --steps 100 --warmup 3 --precision fp32 --no-cpu > gpurun_out/bench_c2_fp32.json
python bench.py --steps 100 --warmup 3 --host-result copy --no-cpu > gpurun_out/bench_c2_copy.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json
for w in c2 c4 c5; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_$w.csv python bench.py --steps 2 --warmup 3 --no-cpu --workload $w > gpurun_out/ncu_launch_$w.log 2>&1
done
cap() {  # name regex skip count workload traffic_key keep
ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o gpurun_out/prof_$1 -f python bench.py --steps 2 --warmup 3 --no-cpu --workload $5 > gpurun_out/ncu_$1.log 2>&1
python profiles/tools/mk_profiles.py --summarise gpurun_out/prof_$1.ncu-rep gpurun_out/summ/$1.txt $6
if [ "$7" != keep ]; then rm -f gpurun_out/prof_$1.ncu-rep; fi
}
cap points_c2 k_rr_points 3 1 c2 c2 keep
cap points_c3 k_rr_points 1 1 c3 c3 keep
cap points_c5 k_rr_points 3 1 c5 "" drop
cap setup_c2 'k_rr_orbit|k_rr_ldm|k_bin' 9 3 c2 "" drop
cap ts_c4 'k_ts_|k_ldtk' 6 6 c4 c4 drop
cap delta_c2 k_host_delta 3 1 c2 "" drop
ls -la gpurun_out gpurun_out/summ | tail -40; du -sh gpurun_out

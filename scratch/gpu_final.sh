# Full measurement pass: parity tests, bench lines for every config, launch lists and ncu --set full captures.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -3 gpurun_out/bench_c2.err; cut -c1-400 gpurun_out/bench_c2.json
python bench.py --steps 50 --warmup 3 --workload c5 --no-cpu > gpurun_out/bench_c5.json
python bench.py --steps 10 --warmup 3 --workload c3 --no-cpu > gpurun_out/bench_c3.json
python bench.py --steps 10 --warmup 3 --workload c4 --no-cpu > gpurun_out/bench_c4.json
python bench.py --steps 200 --warmup 3 --workload c1 --no-cpu > gpurun_out/bench_c1.json
python bench.py --steps 100 --warmup 3 --precision fp32 --no-cpu > gpurun_out/bench_c2_fp32.json
python bench.py --steps 100 --warmup 3 --host-result copy --no-cpu > gpurun_out/bench_c2_copy.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json
for w in c2 c4; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_$w.csv python bench.py --steps 2 --warmup 3 --no-cpu --workload $w > gpurun_out/ncu_launch_$w.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:k_rr_points -s 3 -c 1 -o gpurun_out/prof_points_c2 -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rr_points -s 1 -c 1 -o gpurun_out/prof_points_c3 -f python bench.py --steps 1 --warmup 1 --no-cpu --workload c3 > gpurun_out/ncu_full3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rr_points -s 3 -c 1 -o gpurun_out/prof_points_c5 -f python bench.py --steps 2 --warmup 3 --no-cpu --workload c5 > gpurun_out/ncu_full5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_rr_orbit|k_rr_ldm|k_bin' -s 9 -c 3 -o gpurun_out/prof_setup_c2 -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full_s.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_ts_|k_ldtk' -s 6 -c 6 -o gpurun_out/prof_ts_c4 -f python bench.py --steps 1 --warmup 1 --no-cpu --workload c4 > gpurun_out/ncu_full4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_host_delta -s 3 -c 1 -o gpurun_out/prof_delta_c2 -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full_d.log 2>&1
ls -la gpurun_out | tail -30


mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
python bench.py --steps 100 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -3 gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json
python bench.py --steps 20 --warmup 3 --workload c5 --no-cpu | tee gpurun_out/bench_c5.json
python bench.py --steps 5 --warmup 3 --workload c3 --no-cpu | tee gpurun_out/bench_c3.json
python bench.py --impl reference --steps 3 --warmup 1 | tee gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rr_points -s 3 -c 1 -o gpurun_out/prof_points_c2 -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rr_points -s 1 -c 1 -o gpurun_out/prof_points_c3 -f python bench.py --steps 1 --warmup 1 --no-cpu --workload c3 > gpurun_out/ncu_full3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_rr_orbit|k_rr_ldm|k_bin' -s 12 -c 4 -o gpurun_out/prof_setup_c2 -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full_s.log 2>&1
ls -la gpurun_out

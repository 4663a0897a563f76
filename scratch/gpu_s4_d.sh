mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_rr_points -s 1 -c 1 -o gpurun_out/prof_points_c3 -f python bench.py --steps 1 --warmup 1 --no-cpu --workload c3 > gpurun_out/ncu_full3.log 2>&1
tail -2 gpurun_out/ncu_full3.log
ncu --set full --clock-control none --import-source on -k regex:k_rr_points -s 3 -c 1 -o gpurun_out/prof_points_c5 -f python bench.py --steps 2 --warmup 3 --no-cpu --workload c5 > gpurun_out/ncu_full5.log 2>&1
tail -2 gpurun_out/ncu_full5.log

ncu --set full --clock-control none --import-source on -k regex:k_rr_ldm -s 3 -c 1 -o gpurun_out/prof_ldm_c2 -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rr_points -s 3 -c 1 -o gpurun_out/prof_points_c2_v3 -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_b.log 2>&1
ls -la gpurun_out/*.ncu-rep

mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_ts_setup|k_ts_ldm|k_ldtk_profiles_slab' -s 3 -c 3 -o gpurun_out/prof_ts_c4 -f python bench.py --steps 1 --warmup 1 --no-cpu --workload c4 > gpurun_out/ncu_full4.log 2>&1
tail -2 gpurun_out/ncu_full4.log

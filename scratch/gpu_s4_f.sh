mkdir -p gpurun_out
for w in c4 c1 c2; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_$w.csv python bench.py --steps 2 --warmup 3 --no-cpu --workload $w > gpurun_out/ncu_launch_$w.log 2>&1
done
ls -la gpurun_out/launches_*.csv

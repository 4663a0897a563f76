set -x
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -8
python bench.py --steps 100 --warmup 3 --no-cpu | tee gpurun_out/bench_c2.json
python bench.py --steps 20 --warmup 3 --workload c5 --no-cpu | tee gpurun_out/bench_c5.json
python bench.py --steps 5 --warmup 3 --workload c3 --no-cpu | tee gpurun_out/bench_c3.json

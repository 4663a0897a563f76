mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
python bench.py --steps 100 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -3 gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json
python bench.py --steps 20 --warmup 3 --workload c5 --no-cpu | tee gpurun_out/bench_c5.json
python bench.py --steps 5 --warmup 3 --workload c3 --no-cpu | tee gpurun_out/bench_c3.json
python bench.py --steps 5 --warmup 3 --workload c4 --no-cpu | tee gpurun_out/bench_c4.json
python bench.py --steps 200 --warmup 3 --workload c1 --no-cpu | tee gpurun_out/bench_c1.json
python bench.py --steps 100 --warmup 3 --precision fp32 --no-cpu | tee gpurun_out/bench_c2_fp32.json
python bench.py --impl reference --steps 3 --warmup 1 | tee gpurun_out/bench_ref.json
nproc; lscpu | grep -E "Model name|^CPU\(s\)"

mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
nvidia-smi -L
for w in c2 c5; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 3 --workload $w > gpurun_out/bench_${w}_n2.json 2> gpurun_out/bench_${w}_n2.err
tail -3 gpurun_out/bench_${w}_n2.err; cat gpurun_out/bench_${w}_n2.json
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 | tee gpurun_out/bench_ref_n2.json
python bench.py --steps 50 --warmup 3 --workload c5 --no-cpu | tee gpurun_out/bench_c5.json
python bench.py --steps 10 --warmup 3 --workload c3 --no-cpu | tee gpurun_out/bench_c3.json

mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
for g in peer nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 --workload c5 --gather $g > gpurun_out/bench_c5_n2_$g.json 2> gpurun_out/bench_c5_n2_$g.err
tail -3 gpurun_out/bench_c5_n2_$g.err | cut -c1-300; cut -c1-330 gpurun_out/bench_c5_n2_$g.json
done

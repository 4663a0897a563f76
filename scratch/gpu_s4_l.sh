mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 tests/mgpu_peer_gather_check.py 2>&1 | grep "peer gather\|Error" | head -3
for g in peer nccl; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 200 --warmup 5 --workload c5 --gather $g 2>gpurun_out/bench_c5_n2_$g.err | grep '^{' > gpurun_out/bench_c5_n2_$g.json
cut -c1-260 gpurun_out/bench_c5_n2_$g.json
done

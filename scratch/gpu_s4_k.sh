mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 8 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 100 --warmup 3 2>gpurun_out/bench_c2_n$n.err | grep '^{' > gpurun_out/bench_c2_n$n.json
tail -2 gpurun_out/bench_c2_n$n.err | cut -c1-300; cut -c1-260 gpurun_out/bench_c2_n$n.json
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 100 --warmup 5 --workload c5 2>gpurun_out/bench_c5_n8.err | grep '^{' > gpurun_out/bench_c5_n8.json
tail -2 gpurun_out/bench_c5_n8.err | cut -c1-300; cut -c1-260 gpurun_out/bench_c5_n8.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --steps 100 --warmup 5 --workload c5 --gather nccl 2>gpurun_out/bench_c5_n8_nccl.err | grep '^{' > gpurun_out/bench_c5_n8_nccl.json
cut -c1-260 gpurun_out/bench_c5_n8_nccl.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29535 tests/mgpu_peer_gather_check.py 2>&1 | grep "peer gather\|Error" | head -3

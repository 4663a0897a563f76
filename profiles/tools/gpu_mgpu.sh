mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "peer_memory" 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 5 --workload c5 2>gpurun_out/bench_c5_n2_peer.err | grep '^{' > gpurun_out/bench_c5_n2_peer.json; cut -c1-250 gpurun_out/bench_c5_n2_peer.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 100 --warmup 5 2>gpurun_out/bench_c2_n2.err | grep '^{' > gpurun_out/bench_c2_n2.json; cut -c1-250 gpurun_out/bench_c2_n2.json

# Multi-GPU lines (run with gpurun --gpus N -- 'bash profiles/tools/gpu_mgpu.sh N'): the peer-gather check, then the
# default bench line (C2 headline + the C5-shard `collective` block) under torchrun, one rank per GPU.
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "peer_memory or allgather" 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_peer_gather_check.py 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 100 --warmup 5 2>gpurun_out/bench_c2_n$N.err | grep '^{' > gpurun_out/bench_c2_n$N.json; cut -c1-400 gpurun_out/bench_c2_n$N.json
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_c2_n$N.json').read()); c=d.get('collective') or {}
print('N=$N value %.4e ms %.4f e2e %.4e e2e_copy %.4e' % (d['value'], d['ms_per_step'], (d.get('e2e') or {}).get('value',0), (d.get('e2e_copy') or {}).get('value',0)))
print('collective', {k:c.get(k) for k in ('value','ms_per_step','local_only_ms_per_step','gather_cost_ms_per_step','efficiency_vs_local_only','bit_identical_to_nccl_all_gather','variants')})
PY

# Build-time variants of the supersampled kernel, timed on C3 (run on the GPU box: bash profiles/tools/ss_variants.sh)
b() { python bench.py --steps 10 --warmup 3 --workload $1 --no-cpu --no-numba --no-counters --no-collective 2> gpurun_out/$1.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 $2', d['value'], d['ms_per_step'], d['roofline'].get('kernel_ms'), d['roofline']['frac'])"; }
b c3 default
for v in "$@"; do PTB_NVCC_EXTRA="$v" python -m pytransit_b200.build --force > /dev/null; b c3 "$v"; done

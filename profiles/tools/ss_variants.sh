b() { python bench.py --steps 10 --warmup 3 --workload $1 --no-cpu --no-numba --no-counters --no-collective 2> gpurun_out/$1.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 $2', d['value'], d['ms_per_step'], d['roofline'].get('kernel_ms'), d['roofline']['frac'])"; }
b c3 straight
PTB_NVCC_EXTRA="-DSS_BRANCHY_TAIL=1" python -m pytransit_b200.build --force > /dev/null; b c3 branchy
PTB_NVCC_EXTRA="-DSS_QCAP_=512" python -m pytransit_b200.build --force > /dev/null; b c3 qcap512
PTB_NVCC_EXTRA="-DSS_QCAP_=128" python -m pytransit_b200.build --force > /dev/null; b c3 qcap128

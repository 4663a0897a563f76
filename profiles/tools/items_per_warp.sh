b() { python bench.py --steps 10 --warmup 3 --workload $1 --no-cpu --no-numba --no-counters --no-collective 2> gpurun_out/$1.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 $2', d['ms_per_step'], d['roofline'].get('kernel_ms'))"; }
for ipw in 8 16 32 64; do PTB_ITEMS_PER_WARP=$ipw b c3 ipw$ipw; done
for ipw in 8 16 32; do PTB_ITEMS_PER_WARP=$ipw b c5 ipw$ipw; PTB_ITEMS_PER_WARP=$ipw b c2 ipw$ipw; done

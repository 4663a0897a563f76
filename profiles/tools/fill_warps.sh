# (belongs to the warp-role experiment recorded in DESIGN.md section 8; the PTB_FILL_WARPS switch is not in the tree)
# Sweep of the fill-role warp count of k_rr_points (flux mode) on C2 (run on the GPU box)
b() { python bench.py --steps 50 --warmup 3 --workload $1 $3 --no-cpu --no-numba --no-counters --no-collective 2> gpurun_out/$1.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 $3 $2', d['ms_per_step'], d['roofline'].get('kernel_ms'), d['roofline']['frac'])"; }
for fw in 0 1 2 3 4; do PTB_FILL_WARPS=$fw b c2 fw$fw; done
for fw in 0 1 2; do PTB_FILL_WARPS=$fw b c2 fw$fw "--precision fp32"; done
for fw in 0 1; do PTB_FILL_WARPS=$fw b c1 fw$fw; done

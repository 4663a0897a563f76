# C3 kernel-time of build variants (PTB_NVCC_EXTRA), one line each:  bash profiles/tools/gpu_c3_variants.sh "<flags1>" "<flags2>" ...
mkdir -p gpurun_out
for v in "$@"; do
  PTB_NVCC_EXTRA="$v" python -m pytransit_b200.build --force > /dev/null 2>&1 || { echo "build failed: $v"; continue; }
  timeout 300 python bench.py --workload c3 --steps 10 --no-cpu --no-collective > gpurun_out/var.json 2> gpurun_out/var.err
  python - "$v" <<'PY'
import json, sys
try:
    d = json.load(open('gpurun_out/var.json')); r = d['roofline']
    print('%-50s kernel_ms %.3f step_ms %.3f frac %.3f' % (sys.argv[1], r['kernel_ms'], d['ms_per_step'], r['frac']))
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex, open('gpurun_out/var.err').read()[-500:])
PY
done
python -m pytransit_b200.build --force > /dev/null 2>&1

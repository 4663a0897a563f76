mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -k full_size --durations=5 2>&1 | tail -25 | tee gpurun_out/pytest_full.log

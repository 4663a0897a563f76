mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 --print-limit 20 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity.py::test_peer_memory_allgather_two_gpus 2>&1 | tail -30 > gpurun_out/sanitize_memcheck.log
tail -6 gpurun_out/sanitize_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 1 --print-limit 20 python -m pytest tests -m gpu -q -x -k "golden or edge or lnlike_blocks or graph or spectroscopy" 2>&1 | tail -30 > gpurun_out/sanitize_racecheck.log
tail -6 gpurun_out/sanitize_racecheck.log

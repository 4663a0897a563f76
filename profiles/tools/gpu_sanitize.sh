# compute-sanitizer evidence (run on the GPU box: bash profiles/tools/gpu_sanitize.sh): memcheck over the single-GPU
# parity suite (full-size property tests excluded: minutes each under the tool), racecheck + synccheck over the tests
# that exercise every kernel family (shared-memory queues, phase frames, TMA record slots, cluster sort).
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 --print-limit 20 python -m pytest tests -m gpu -q -x -k "not full_size and not pipelined_full_size" --deselect tests/test_gpu_parity.py::test_peer_memory_allgather_two_gpus 2>&1 | tail -30 > gpurun_out/sanitize_memcheck.log
tail -4 gpurun_out/sanitize_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 1 --print-limit 20 python -m pytest tests -m gpu -q -x -k "golden or edge or lnlike_blocks or graph or spectroscopy or supersampl or c3_shaped or c5_shaped" 2>&1 | tail -30 > gpurun_out/sanitize_racecheck.log
tail -4 gpurun_out/sanitize_racecheck.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 1 --print-limit 20 python -m pytest tests -m gpu -q -x -k "golden or supersampl or c3_shaped or c5_shaped" 2>&1 | tail -30 > gpurun_out/sanitize_synccheck.log
tail -4 gpurun_out/sanitize_synccheck.log

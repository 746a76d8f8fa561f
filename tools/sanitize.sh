# memory-safety / race evidence: run a subset of the kernel tests under compute-sanitizer
export PATH=/usr/local/cuda/bin:/usr/local/cuda/compute-sanitizer:$PATH
which compute-sanitizer
timeout 700 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_kernels_gpu.py -q -x -m gpu -k "test_gemm_plain or test_gemm_swiglu or test_gemm_rope or test_attention or test_rmsnorm or test_embed or test_align_softmax or test_ce" --timeout 600 --timeout-method=thread 2>&1 | tail -12
echo "memcheck rc=$?"
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_kernels_gpu.py -q -x -m gpu -k "test_rmsnorm or test_align_softmax or test_ce or test_patchify" --timeout 450 --timeout-method=thread 2>&1 | tail -8

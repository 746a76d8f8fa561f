timeout 900 python -m pytest tests/test_model_gpu.py -q -s -m gpu -k "real_width" --timeout 600 --timeout-method=thread 2>&1 | grep -E "parity|passed|failed|Error|error" | tail -8
python tools/profile_align.py --iters 3 2>&1 | grep profile_align
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gemm_bf16_kernel|align_softmax" -c 12 -f -o gpurun_out/prof_align_r1 python tools/profile_align.py --iters 1 2>&1 | grep -E "profile_align|Report"

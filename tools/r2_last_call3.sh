# verification of the mm_gemm_plan refactor (gemm_dispatch) on the GPU: kernel + training tests, smoke, short bench
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_kernels_gpu.py tests/test_train_gpu.py -x -q -m gpu 2>&1 | tail -3 > gpurun_out/last3_tests.log
tail -2 gpurun_out/last3_tests.log
timeout 60 python __graft_entry__.py --smoke 2>&1 | grep -E "smoke|build ok|Error" | tail -3
timeout 80 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/last_bench3.err | grep "^{" > gpurun_out/last_bench3.json
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/last_bench3.json").read())
    print("bench ms/step", round(d["ms_per_step"], 2), "tok/s", round(d["value"]), d["clocks"])
except Exception as e:
    print("bench line missing:", e)
PY

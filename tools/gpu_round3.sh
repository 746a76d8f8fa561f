timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention" --timeout 300 --timeout-method=thread 2>&1 | tail -5
timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu --timeout 300 --timeout-method=thread 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:_kernel -c 1500 --csv --log-file gpurun_out/launches.csv python tools/profile_forward.py --batch 32 --warm 1 --steps 1 2>&1 | grep profile_forward
python tools/summarize_launches.py gpurun_out/launches.csv | head -12
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | grep "^{" > gpurun_out/bench_tmp.json; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_tmp.json").read())
print("ms/step", d["ms_per_step"], "tok/s", d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], d["clocks"])
print("gemm", d["roofline"]["achieved"], d["roofline"]["share_of_step"])
for k,v in d["roofline"]["by_section"].items(): print("   ",k, round(v["tflops"]), round(v["ms_per_step"],2), v["launches_per_step"])
PY

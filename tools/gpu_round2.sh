timeout 900 python -m pytest tests -q -m gpu --timeout 300 --timeout-method=thread 2>&1 | grep -v "Warning\|generative\|trust_remote\|owner" | tail -25
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | grep "^{" > gpurun_out/bench_tmp.json; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_tmp.json").read())
print("ms/step", d["ms_per_step"], "tok/s", d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], d["clocks"])
print("gemm", d["roofline"]["achieved"], d["roofline"]["share_of_step"])
for k,v in d["roofline"]["by_section"].items(): print("   ",k, round(v["tflops"]), round(v["ms_per_step"],2), v["launches_per_step"])
PY

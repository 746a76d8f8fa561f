# what the driver runs at round end, in one go
timeout 1200 python -m pytest tests/ -x -q -m gpu --timeout 600 --timeout-method=thread 2>&1 | tail -4
python __graft_entry__.py --smoke 2>&1 | grep -E "smoke|build ok|Error"
timeout 900 python bench.py 2>&1 | grep "^{" > gpurun_out/bench_final.json; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_final.json").read())
print("N=1 ms/step", round(d["ms_per_step"],2), "tok/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"], d["clocks"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
print("roofline", round(d["roofline"]["achieved"]), round(d["roofline"]["frac"],3), d["roofline"]["traffic"], round(d["roofline"]["share_of_step"],3))
PY

timeout 1200 python -m pytest tests/ -x -q -m gpu --timeout 600 --timeout-method=thread 2>&1 | tail -3
for cfg in "1 32" "0 32" "1 4" "0 4" "1 4" "0 4"; do set -- $cfg; MACAW_B200_PDL=$1 timeout 600 python bench.py --global-batch $2 --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | grep "^{" > gpurun_out/bench_tmp.json; python -c "
import json
d=json.loads(open(\"gpurun_out/bench_tmp.json\").read())
print(\"PDL=$1 B=$2 ms/step\", round(d[\"ms_per_step\"],3), \"tok/s\", round(d[\"value\"]), d[\"clocks\"][\"sm_mhz\"])
"; done

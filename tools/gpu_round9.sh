timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention" --timeout 300 --timeout-method=thread 2>&1 | tail -3
timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu --timeout 600 --timeout-method=thread 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:_kernel -c 1500 --csv --log-file gpurun_out/launches.csv python tools/profile_forward.py --batch 32 --warm 1 --steps 1 2>&1 | grep profile_forward
python tools/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt; head -12 gpurun_out/launches_summary.txt
for i in 1 2; do timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | grep "^{" > gpurun_out/bench_tmp.json; python -c "
import json
d=json.loads(open(\"gpurun_out/bench_tmp.json\").read())
print(\"ms/step\", round(d[\"ms_per_step\"],2), \"tok/s\", round(d[\"value\"]), d[\"clocks\"][\"sm_mhz\"], \"gemm\", round(d[\"roofline\"][\"achieved\"]), round(d[\"roofline\"][\"share_of_step\"],3), {k: round(v[\"tflops\"]) for k,v in d[\"roofline\"][\"by_section\"].items()})
"; done

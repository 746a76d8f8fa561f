timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "norm" --timeout 300 --timeout-method=thread 2>&1 | tail -3
for gm in 16 4 8 32 16; do MACAW_B200_GEMM_GROUPM=$gm timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | grep "^{" > gpurun_out/bench_tmp.json; python -c "
import json
d=json.loads(open(\"gpurun_out/bench_tmp.json\").read())
print(\"GROUPM=$gm ms/step\", round(d[\"ms_per_step\"],2), \"tok/s\", round(d[\"value\"]), d[\"clocks\"][\"sm_mhz\"], \"gemm\", round(d[\"roofline\"][\"achieved\"]), {k: round(v[\"tflops\"]) for k,v in d[\"roofline\"][\"by_section\"].items()})
"; done

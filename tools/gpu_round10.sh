timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention" --timeout 300 --timeout-method=thread 2>&1 | tail -2
timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu -k "golden or encoders" --timeout 600 --timeout-method=thread 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:fa_tcgen05 -c 100 --csv --log-file gpurun_out/launches_fa.csv python tools/profile_forward.py --batch 32 --warm 1 --steps 1 2>&1 | grep profile_forward
python tools/summarize_launches.py gpurun_out/launches_fa.csv

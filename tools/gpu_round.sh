# one GPU round-trip: kernel + model parity tests, then a kernel launch list of one cfg4 forward
timeout 900 python -m pytest tests -q -m gpu --timeout 300 --timeout-method=thread 2>&1 | grep -v "Warning\|generative\|trust_remote\|owner" | tail -15
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:_kernel -c 1500 --csv --log-file gpurun_out/launches.csv python tools/profile_forward.py --batch 32 --warm 1 --steps 1 2>&1 | grep profile_forward
python tools/summarize_launches.py gpurun_out/launches.csv | tee gpurun_out/launches_summary.txt

# ncu evidence for profiles/: launch list of one cfg4 forward + full captures of the top kernels
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:_kernel -c 1500 --csv --log-file gpurun_out/launches_r1.csv python tools/profile_forward.py --batch 32 --warm 1 --steps 1 2>&1 | grep profile_forward
python tools/summarize_launches.py gpurun_out/launches_r1.csv | tee gpurun_out/launches_r1_summary.txt
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_bf16_kernel -s 140 -c 6 -f -o gpurun_out/prof_gemm_r1 python tools/profile_forward.py --batch 32 --warm 1 --steps 1 2>&1 | tail -1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:fa_tcgen05_kernel -s 31 -c 2 -f -o gpurun_out/prof_fa_r1 python tools/profile_forward.py --batch 32 --warm 1 --steps 1 2>&1 | tail -1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_bf16_kernel -s 124 -c 4 -f -o gpurun_out/prof_align_r1 python tools/profile_forward.py --batch 32 --warm 1 --steps 1 --layers 1 2>&1 | tail -1

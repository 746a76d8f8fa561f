# ncu evidence for profiles/ (round 1, final kernels): launch list of one cfg4 forward + full captures of the top kernels
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:_kernel -c 1500 --csv --log-file gpurun_out/launches_r1.csv python tools/profile_forward.py --batch 32 --warm 1 --steps 1 2>&1 | grep profile_forward
python tools/summarize_launches.py gpurun_out/launches_r1.csv > gpurun_out/launches_r1_summary.txt
# LLaMA layer 0: rms_rstd, QKV+RoPE GEMM, attention, o_proj, rms_rstd, gate/up+SwiGLU, down  (kernel-id filter skips encoders)
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gemm_bf16_kernel|fa_tcgen05" -s 189 -c 5 -f -o gpurun_out/prof_llama_layer_r1 python tools/profile_forward.py --batch 32 --warm 1 --steps 1 2>&1 | tail -1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gemm_bf16_kernel|align_softmax" -c 12 -f -o gpurun_out/prof_align_r1 python tools/profile_align.py --iters 1 2>&1 | grep -E "Report"

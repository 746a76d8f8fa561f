set -x
python - <<'PY'
import os, time, torch
print("affinity", len(os.sched_getaffinity(0)), "cpu_count", os.cpu_count())
try: print("cpu.max", open("/sys/fs/cgroup/cpu.max").read().strip())
except Exception as e: print("no cpu.max", e)
a = torch.randn(4096, 4096); b = torch.randn(4096, 4096)
for nt in (8, 16, 32, 64, 128):
    torch.set_num_threads(nt); a @ b
    t = time.perf_counter(); [a @ b for _ in range(5)]; dt = (time.perf_counter() - t) / 5
    print(f"threads {nt}: {2*4096**3/dt/1e12:.2f} TFLOP/s")
PY
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:mm:: -c 1500 --csv --log-file gpurun_out/launches_r1.csv python tools/profile_forward.py --batch 32 --warm 1 --steps 1 2>&1 | grep -v "Warning\|generative\|trust_remote\|owner" | tail -3
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_bf16_kernel -s 240 -c 2 -f -o gpurun_out/prof_gemm_r1 python tools/profile_forward.py --batch 32 --warm 1 --steps 1 2>&1 | tail -2
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:flash_attn_kernel -s 40 -c 1 -f -o gpurun_out/prof_attn_r1 python tools/profile_forward.py --batch 32 --warm 1 --steps 1 2>&1 | tail -2
ls -la gpurun_out/

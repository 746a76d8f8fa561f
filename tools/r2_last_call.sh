# Last GPU call of round 2 (14 GPU-minutes left): re-check the final tree where the last kernel commits touched it,
# a fresh default bench line, and the one ncu capture round 2 lacks (head_dim-96 video-long attention on tcgen05).
# Every step has its own timeout; nothing here can hold the box past ~9.5 minutes.
mkdir -p gpurun_out
# 1. tests that exercise the last kernel changes (cta_group::2 for MN-major B = training dX GEMMs; GEMM / attention kernels)
timeout 270 python -m pytest tests/test_train_gpu.py tests/test_kernels_gpu.py -x -q -m gpu --durations=8 2>&1 | tail -16 > gpurun_out/last_tests_a.log
tail -3 gpurun_out/last_tests_a.log
# 2. default bench line on the final tree (CPU arm skipped here: the driver runs it at round end)
timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/last_bench.err | grep "^{" > gpurun_out/last_bench.json
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/last_bench.json").read())
    print("bench ms/step", round(d["ms_per_step"], 2), "tok/s", round(d["value"]), "e2e", round(d["e2e"]["value"]),
          "launches", d["gpu_launches"], d["clocks"], "roofline", round(d["roofline"]["frac"], 3))
except Exception as e:
    print("bench line missing:", e)
PY
# 3. ncu --set full of the video-long self-attention (cfg5 shape: F=16 -> 4096 + 2 keys, 8 heads of 96) — the HD=128
#    instantiation is the first `fa_tcgen05_kernel<128,...>` launch of a video forward (CLIP / Whisper use <64,...>)
timeout 140 ncu --set full --import-source on --clock-control none --kernel-name-base demangled \
  -k regex:'fa_tcgen05_kernel<128' -c 1 -f -o gpurun_out/attn_hd96 \
  python tools/profile_forward.py --video-frames 16 --batch 2 --layers 1 > gpurun_out/last_ncu.log 2>&1
tail -2 gpurun_out/last_ncu.log
# 4. the rest of the GPU tier for as long as the budget allows
timeout 200 python -m pytest tests/test_model_gpu.py tests/test_fp16_gpu.py tests/test_inputs.py tests/test_wire.py -x -q -m gpu 2>&1 | tail -4 > gpurun_out/last_tests_b.log
tail -2 gpurun_out/last_tests_b.log

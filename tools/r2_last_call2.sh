# Follow-up to r2_last_call.sh (its ncu filter missed: demangled names print template arguments as `(int)128`).
mkdir -p gpurun_out
# 1. ncu --set full of the video-long self-attention (cfg5 shape: 4096 + 2 keys, 8 heads of 96, HD=128 instantiation)
timeout 150 ncu --set full --import-source on --clock-control none --kernel-name-base demangled \
  -k regex:'fa_tcgen05_kernel<.int.128' -c 1 -f -o gpurun_out/attn_hd96 \
  python tools/profile_forward.py --video-frames 16 --batch 2 --layers 1 > gpurun_out/last_ncu.log 2>&1
grep -E "PROF|profile_forward|WARNING" gpurun_out/last_ncu.log | tail -4
# 2. launch list of one cfg5 forward at the per-GPU batch of the 8-GPU run (B=2), full depth
timeout 140 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_cfg5_b2.csv \
  python tools/profile_forward.py --video-frames 16 --batch 2 --warm 0 > gpurun_out/last_ncu2.log 2>&1
tail -1 gpurun_out/last_ncu2.log
# 3. a second default bench line (the first call's box sat at 1061 MHz)
timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/last_bench2.err | grep "^{" > gpurun_out/last_bench2.json
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/last_bench2.json").read())
    print("bench ms/step", round(d["ms_per_step"], 2), "tok/s", round(d["value"]), "e2e", round(d["e2e"]["value"]),
          "launches", d["gpu_launches"], d["clocks"], "roofline", round(d["roofline"]["frac"], 3))
except Exception as e:
    print("bench line missing:", e)
PY

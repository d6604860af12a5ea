#!/usr/bin/env bash
# The ncu captures behind profiles/ (run on a B200 box, e.g. `gpurun --timeout 900 -- 'bash tools/profile.sh r02a'`).
# Numbers printed by a run under ncu are never bench values; bench.py times with CUDA events.
set -euo pipefail
tag=${1:-rXX}
out=gpurun_out
mkdir -p "$out"
if [ "${2:-}" != "decode" ]; then
# 1. launch list of ONE steady-state attack iteration (bench.py --ncu-step brackets it with cudaProfilerStart/Stop), with
#    DRAM bytes per launch; single stream so that the two vision towers do not interleave in the list
VLA_SINGLE_STREAM=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --profile-from-start off --csv --log-file "$out/${tag}_launches.csv" python bench.py --ncu-step > "$out/ncu_step.log" 2>&1
python tools/summarize_launches.py "$out/${tag}_launches.csv" "$out/${tag}_gemm_dram_traffic.json" > "$out/${tag}_launches_one_step.txt"
# 2. the dominant kernel, once per change: the CTA-pair GEMM on Llama's gate|up shape and on the DINOv2 fc1+GELU shape
ncu --set full --clock-control none --import-source on -k regex:gemm -s 2 -c 1 -f -o "$out/${tag}_prof_gemm_llama" \
    python tools/one_gemm.py 2304 22016 4096 plain 2 256 > "$out/ncu_gemm.log" 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm -s 2 -c 1 -f -o "$out/${tag}_prof_gemm_gelu" \
    python tools/one_gemm.py 2088 4096 1024 gelu 2 256 >> "$out/ncu_gemm.log" 2>&1
# 3. the attention kernels of one Llama layer (forward, delta, dQ, dK|dV)
ncu --set full --clock-control none --import-source on -k regex:attn -s 4 -c 4 -f -o "$out/${tag}_prof_attn" \
    python tools/one_attn.py 8 288 32 128 1 3 > "$out/ncu_attn.log" 2>&1
# read the reports in the authoring container:  ncu -i <file>.ncu-rep --page raw --csv   /   --page source --csv
echo "wrote $out/${tag}_*"
fi
# 4. (bash tools/profile.sh <tag> decode) the greedy action decode: launch list of two predict_action calls (prefill + 6 recorded
#    steps each) and a full capture of the skinny projections of one layer (q|k|v, o, gate|up, down) of a replayed step
if [ "${2:-}" = "decode" ]; then
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
      --log-file "$out/${tag}_decode_launches.csv" python tools/decode_bench.py --profile > "$out/ncu_decode.log" 2>&1
  ncu --set full --clock-control none --import-source on -k regex:gemv_kernel -s 300 -c 4 -f -o "$out/${tag}_prof_gemv" \
      python tools/decode_bench.py --profile >> "$out/ncu_decode.log" 2>&1
  ncu --set full --clock-control none --import-source on -k regex:attn_decode -s 40 -c 1 -f -o "$out/${tag}_prof_attn_decode" \
      python tools/decode_bench.py --profile >> "$out/ncu_decode.log" 2>&1
  echo "wrote $out/${tag}_decode_*"
fi

#!/usr/bin/env bash
# compute-sanitizer over the kernel-level parity tests, the tiny engine step and the graph-replayed attack step
# (run on a B200 box: `gpurun --timeout 1700 -- 'bash tools/sanitize.sh'`; slow, hence the selections and per-tool limits).
# memcheck: out-of-bounds / misaligned accesses (the 256-bit epilogue stores, TMA boxes at the sequence edges);
# racecheck: shared-memory hazards between the producer / MMA / softmax warps; synccheck: barrier misuse.
set -uo pipefail
out=gpurun_out
mkdir -p "$out"
limit=${SAN_LIMIT:-500}
sel=${1:-"gemm or gemv or attention or layernorm or rmsnorm or loss_head or frontend or patch_update or engine_step_vs_oracle or graph_replay or greedy"}
for tool in memcheck racecheck synccheck; do
  timeout "$limit" compute-sanitizer --tool "$tool" --error-exitcode 1 --log-file "$out/sanitizer_${tool}.log" \
      python -m pytest tests/test_kernels_gpu.py tests/test_decode_kernels_gpu.py tests/test_engine_gpu.py tests/test_attack_step_gpu.py -q -p no:cacheprovider -k "$sel" \
      > "$out/sanitizer_${tool}.out" 2>&1
  rc=$?
  echo "$tool: exit $rc (124 = stopped at the ${limit}s limit); $(grep -h 'ERROR SUMMARY' "$out/sanitizer_${tool}.log" 2>/dev/null | tail -1); $(tail -1 "$out/sanitizer_${tool}.out")"
done

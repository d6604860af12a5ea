#!/usr/bin/env bash
# compute-sanitizer over the kernel-level parity tests (run on a B200 box; slow: keep to the kernel tests and the toy engine).
# memcheck: out-of-bounds / misaligned accesses (the 256-bit epilogue stores, TMA boxes at the sequence edges);
# racecheck: shared-memory hazards between the producer / MMA / softmax warps; synccheck: barrier misuse.
set -uo pipefail
out=gpurun_out
mkdir -p "$out"
sel=${1:-"gemm or attention or layernorm or rmsnorm or loss_head or frontend"}
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool "$tool" --error-exitcode 1 --log-file "$out/sanitizer_${tool}.log" \
      python -m pytest tests/test_kernels_gpu.py -x -q -k "$sel" > "$out/sanitizer_${tool}.out" 2>&1
  echo "$tool: exit $? ($(grep -c 'ERROR SUMMARY' "$out/sanitizer_${tool}.log" 2>/dev/null) summaries)"
done

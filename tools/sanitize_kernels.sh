#!/usr/bin/env bash
# compute-sanitizer pass over the sm_100a kernels (needs a B200; minutes per tool): memcheck (out-of-bounds / misaligned),
# racecheck (shared-memory hazards between the warp roles), synccheck (barrier misuse), initcheck (uninitialised global reads).
# The reference has no sanitizer story (single-threaded Python); SURVEY.md section 5.2 asks for this on the native side.
#   tools/sanitize_kernels.sh [memcheck|racecheck|synccheck|initcheck ...]        logs -> gpurun_out/sanitize_<tool>.log
# TMA / tcgen05 traffic is only partially visible to racecheck; the mbarrier protocol errors it cannot see are caught by the
# bounded waits (every blocking wait traps with its barrier tag after a wall-clock limit).
set -uo pipefail
cd "$(dirname "$0")/.."
TOOLS=("${@:-memcheck}")
[ "$#" -eq 0 ] && TOOLS=(memcheck racecheck synccheck)
mkdir -p gpurun_out
TARGETS=(
  "tests/test_elementwise_gpu.py"
  "tests/test_gemm_gpu.py -k 'not grouped and not fp8' --maxfail=3"
  "tests/test_attention_gpu.py --maxfail=3"
)
rc=0
for tool in "${TOOLS[@]}"; do
  log="gpurun_out/sanitize_${tool}.log"
  : > "$log"
  for t in "${TARGETS[@]}"; do
    echo "== compute-sanitizer --tool $tool : pytest $t" | tee -a "$log"
    eval timeout 1500 compute-sanitizer --tool "$tool" --error-exitcode 66 --print-limit 5 --target-processes all \
         python -m pytest $t -q -m gpu -x >> "$log" 2>&1
    code=$?
    [ "$code" -ne 0 ] && { echo "   exit code $code (66 = sanitizer errors)" | tee -a "$log"; rc=1; }
  done
  grep -c "ERROR SUMMARY: 0 errors" "$log" | xargs -I{} echo "   $tool: {} clean runs" | tee -a "$log"
done
exit $rc

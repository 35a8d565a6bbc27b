#!/bin/bash
# compute-sanitizer over small cases of the kernels that synchronise through mbarriers, global flags and proxy
# fences (SURVEY section 5 "Race detection"): umma_trunk_kernel, umma_conv3x3_kernel, deform_umma_kernel,
# local_trunk_kernel<fwd/bwd>, flat_conv_kernel, flat_chain_kernel, flat_wgrad_kernel (scripts/sanitize_cases.py).
# Only the library's own kernels are instrumented (they live in namespace dbm); torch's allocator kernels are not.
#   usage: [CASES="split chain64"] scripts/sanitize.sh [outdir]        logs: <outdir>/sanitize_<tool>_<case>.log + sanitize_summary.txt
out=${1:-gpurun_out}
mkdir -p "$out"
: > "$out/sanitize_summary.txt"
cd "$(dirname "$0")/.."
for tool in memcheck racecheck synccheck; do
  for c in ${CASES:-trunk split local train chain chain64}; do
    log="$out/sanitize_${tool}_${c}.log"
    timeout 900 compute-sanitizer --tool $tool --kernel-name kns=dbm --print-limit 20 --error-exitcode 9 \
      python scripts/sanitize_cases.py $c > "$log" 2>&1
    rc=$?
    line=$(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" "$log" | tail -1)
    echo "$tool $c rc=$rc ${line}" | tee -a "$out/sanitize_summary.txt"
  done
done

#!/usr/bin/env bash
# compute-sanitizer over the kernels whose correctness rests on hand-rolled synchronisation: the mbarrier pipelines and
# TMEM hand-offs of the tcgen05 GEMM / convolutions, the flag-chained persistent recurrence (global release/acquire counters),
# the cluster split-K reduction through DSMEM and the ticketed reductions.  Reduced shapes: the sanitizer slows kernels 10-100x.
#   scripts/sanitize.sh [memcheck|racecheck|synccheck|initcheck ...]     (default: memcheck racecheck synccheck)
# Output: gpurun_out/sanitize_<tool>.log and a one-line verdict per tool in gpurun_out/sanitize_summary.txt
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TOOLS=${@:-memcheck racecheck synccheck}
# memcheck / synccheck: every family except the persistent recurrence, whose CTAs spin on each other's global flags (the tool's slowdown
# trips the kernel's own spin watchdog; HULC_B200_PERSISTENT_RNN=0 runs the step's recurrences as per-step products instead).
# racecheck (shared-memory hazards, ~100x slower): the two GEMM kernels and the convolutions at one frame.
cases_for() { case "$1" in racecheck) echo "gemm gemm_bf16 conv small";; *) echo "gemm gemm_bf16 conv step";; esac; }
export HULC_B200_PERSISTENT_RNN=0
: > gpurun_out/sanitize_summary.txt
for tool in $TOOLS; do
  log=gpurun_out/sanitize_${tool}.log
  HULC_B200_SANITIZE=1 timeout ${HULC_SANITIZE_TIMEOUT:-900} compute-sanitizer --tool "$tool" --error-exitcode 86 --print-limit 20 \
      python scripts/sanitize_cases.py $(cases_for "$tool") > "$log" 2>&1
  rc=$?
  errs=$(grep -c "^========= \(Invalid\|Race\|Error\|Barrier\|Uninitialized\|Program hit\)" "$log" || true)
  echo "$tool rc=$rc sanitizer_reports=$errs $(grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ALL OK|Error|assert" "$log" | tail -3 | tr '\n' ' ')" | tee -a gpurun_out/sanitize_summary.txt
done

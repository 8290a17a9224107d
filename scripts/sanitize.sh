#!/usr/bin/env bash
# compute-sanitizer over every kernel family of libcilantro_b200.so at small sizes (run on a GPU box:
#   gpurun --timeout 1500 -- 'bash scripts/sanitize.sh'   -> gpurun_out/sanitize_{memcheck,racecheck,synccheck}.log
# and a one-line verdict per tool on stdout; profiles/r02_sanitizer.md is the committed summary of the last run).
# memcheck: out-of-bounds / misaligned global, shared and local accesses, leaks of device allocations.
# racecheck: shared-memory hazards between threads of a block (the warp-pooled search queues, the barrier-free
#            block reduction). synccheck: divergent / invalid use of __syncthreads / __syncwarp / *_sync shuffles.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SAN=${SAN:-/usr/local/cuda/bin/compute-sanitizer}
rc_all=0
for tool in memcheck racecheck synccheck; do
  extra=""
  [ "$tool" = memcheck ] && extra="--leak-check full"
  log=gpurun_out/sanitize_${tool}.log
  SANITIZE_N=${SANITIZE_N:-6000} timeout 1200 "$SAN" --tool "$tool" $extra --error-exitcode 3 --print-limit 400 \
      python tools/sanitize_target.py > "$log" 2>&1
  rc=$?
  summary=$(grep -E "ERROR SUMMARY|RACECHECK SUMMARY|LEAK SUMMARY" "$log" | tr '\n' ' ')
  echo "[$tool] rc=$rc ${summary}"
  grep -q "all checks passed" "$log" || echo "[$tool] the target did not finish"
  [ $rc -ne 0 ] && rc_all=1
done
exit $rc_all

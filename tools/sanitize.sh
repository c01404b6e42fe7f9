#!/bin/bash
# compute-sanitizer over every kernel family (tools/sanitize_target.py).  memcheck always; racecheck / initcheck / synccheck on the
# environment half (shared-memory staging, mbarriers and TMA live in the foothold kernel) when asked:  tools/sanitize.sh [all]
# Exit code 0 = every tool reported 0 errors.  Logs: gpurun_out/sanitizer_<tool>.log
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CS=${COMPUTE_SANITIZER:-/usr/local/cuda/bin/compute-sanitizer}
rc=0
run() {  # tool, target section
  timeout 1500 "$CS" --tool "$1" --error-exitcode 9 --print-limit 20 python tools/sanitize_target.py "$2" > "gpurun_out/sanitizer_$1.log" 2>&1
  local r=$?
  tail -3 "gpurun_out/sanitizer_$1.log"
  grep -q "sanitize target ok" "gpurun_out/sanitizer_$1.log" || r=8
  [ $r -ne 0 ] && rc=$r
}
run memcheck all
if [ "${1:-}" = "all" ]; then
  run racecheck env
  run synccheck env
  run initcheck env
fi
exit $rc

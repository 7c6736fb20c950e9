#!/bin/bash
# usage (under gpurun, ONE tool per call -- B200_PROFILING.md): scripts/sanitize.sh memcheck|racecheck|synccheck|initcheck
# Runs scripts/sanitize_target.py (tiny VLM, every kernel family, results checked against the oracle) first plain, then
# under compute-sanitizer.  Logs: gpurun_out/sanitize_<tool>.log; copy the summary to profiles/.
tool=${1:-memcheck}
mkdir -p gpurun_out
python scripts/sanitize_target.py > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/sanitize_plain.log; exit 1; }
tail -1 gpurun_out/sanitize_plain.log
timeout ${SANITIZE_TIMEOUT:-600} compute-sanitizer --tool "$tool" --error-exitcode 99 --print-limit 30 python scripts/sanitize_target.py > "gpurun_out/sanitize_$tool.log" 2>&1
rc=$?
echo "compute-sanitizer $tool exit=$rc"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize target ok|Invalid|Race|hazard|Barrier error" "gpurun_out/sanitize_$tool.log" | head -40
exit $rc

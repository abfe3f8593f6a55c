#!/usr/bin/env bash
# Dev helper for the GPU box: run every -m gpu test in its own process (a trapped kernel poisons the CUDA context of
# the process, so isolation keeps the rest of the suite informative). Usage: scripts/run_gpu_tests_isolated.sh [pytest -k expr]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LOG=gpurun_out/isolated.log
: > "$LOG"
KEXPR=${1:-}
if [[ -n "$KEXPR" ]]; then
  ids=$(python -m pytest tests --collect-only -q -m gpu -k "$KEXPR" 2>/dev/null | grep "::")
else
  ids=$(python -m pytest tests --collect-only -q -m gpu 2>/dev/null | grep "::")
fi
pass=0; fail=0
for id in $ids; do
  out=$(timeout 240 python -m pytest -q -x "$id" 2>&1)
  rc=$?
  if [[ $rc -eq 0 ]]; then pass=$((pass+1)); echo "PASS $id" >> "$LOG";
  else fail=$((fail+1)); echo "FAIL($rc) $id" >> "$LOG"; echo "$out" | tail -25 >> "$LOG"; fi
done
echo "isolated: $pass passed, $fail failed" | tee -a "$LOG"
grep -E "^(PASS|FAIL)" "$LOG"

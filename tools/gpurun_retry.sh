#!/usr/bin/env bash
# gpurun with retries while the pod answers "busy" (exit code 3: nothing charged).
# usage: tools/gpurun_retry.sh <timeout_s> '<command>' [gpus]
t=$1; cmd=$2; g=${3:-1}
for i in $(seq 1 30); do
  if [[ $g == 1 ]]; then /usr/local/graft/bin/gpurun --timeout "$t" -- "$cmd"; else /usr/local/graft/bin/gpurun --gpus "$g" --timeout "$t" -- "$cmd"; fi
  rc=$?
  if grep -q '"status": "transient"' gpurun_out/.last_call.json 2>/dev/null || [[ $rc == 3 ]]; then sleep 60; continue; fi
  exit $rc
done
exit 3

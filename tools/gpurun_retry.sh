#!/bin/bash
# gpurun with retries while the pod is busy (exit code 3 / "transient"):  tools/gpurun_retry.sh <timeout> <command...>
T=$1; shift
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$out"; exit $rc
done
echo "gpurun: still busy after 20 tries"; exit 3

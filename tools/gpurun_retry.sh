#!/bin/bash
# gpurun with retries while the pod has no free slot (exit code 3 / "transient"): [GPUS=2] tools/gpurun_retry.sh <timeout> '<command>'
t=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout "$t" -- "$@" > /tmp/gpurun_last.log 2>&1
  if ! grep -q "status=transient" /tmp/gpurun_last.log; then break; fi
  sleep 100
done
cat /tmp/gpurun_last.log

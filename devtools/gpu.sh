#!/bin/bash
# usage: devtools/gpu.sh <logname> <timeout_s> [--gpus N] -- '<command>'   (retries while the pod answers "busy": exit code 3)
name=$1; shift; to=$1; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to "$@" > gpurun_out/$name.log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && [ $rc -ne 2 ]; then echo "rc=$rc attempt=$attempt" >> gpurun_out/$name.log; exit $rc; fi
  sleep 90
done

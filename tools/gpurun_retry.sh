#!/bin/bash
# gpurun_retry.sh LOGFILE GPUS TIMEOUT CMD... : retries while the pod answers "busy" (exit code 3)
LOG=$1; G=$2; T=$3; shift 3
for i in 1 2 3 4 5 6 7 8; do
  if [ "$G" == "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "$@" > $LOG 2>&1; else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@" > $LOG 2>&1; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done

#!/bin/bash
# usage: tools/gpurun_retry.sh OUTFILE TIMEOUT [--gpus N] -- 'command'   : retries while the pod answers "busy" (exit code 3)
out=$1; to=$2; shift 2
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $to "$@" > $out 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3

#!/bin/bash
# usage: tools/gpu_retry_n.sh <gpus> <logfile> <timeout> <command...>
n=$1; shift; log=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --gpus $n --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3

#!/bin/bash
# usage: tools/gpu/retry_n.sh <gpus> <logfile> <timeout> <command...>
n=$1; shift; log=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  gpurun --gpus $n --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" $log; then exit $rc; fi
  sleep 45
done

#!/bin/bash
# usage: tools/gpu.sh <timeout_s> <logfile> [--gpus N] -- '<command>'   (retries while the pod answers "busy")
T=$1; LOG=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $T "$@" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" $LOG; then echo "gpurun rc=$rc" >> $LOG; exit $rc; fi
  sleep 90
done
echo "gave up" >> $LOG

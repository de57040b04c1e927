#!/bin/bash
# gr.sh <log name> <timeout s> <command...>: gpurun with retries while the pod answers "busy" (nothing is charged then)
# GPUS=N in the environment asks for an N-GPU box.
name=$1; to=$2; shift 2
log=gpurun_out/${name}_gpurun.log
for try in 1 2 3 4 5 6 7 8; do
  gpurun ${GPUS:+--gpus $GPUS} --timeout $to -- "$@" > $log 2>&1
  if grep -q "status=transient\|already running" $log; then sleep 60; continue; fi
  break
done
echo done >> $log

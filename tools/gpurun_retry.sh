#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <timeout_s> [--gpus N] -- <command>
# retries a gpurun call while the pod answers "transient" (nothing charged)
log=$1; shift; tmo=$1; shift
extra=()
while [ "$1" != "--" ]; do extra+=("$1"); shift; done
shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$tmo" "${extra[@]}" -- "$@" > "$log" 2>&1
  if grep -q "status=transient" "$log" || grep -q "exit code 3" "$log"; then sleep 150; continue; fi
  break
done
echo done >> "$log"

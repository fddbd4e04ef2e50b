#!/bin/bash
# usage: tools/gpu/submit.sh <log> <timeout_s> [--gpus N] -- <command...>   (retries while the pod answers "busy")
log=$1; to=$2; shift 2
extra=()
while [ "$1" != "--" ]; do extra+=("$1"); shift; done
shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$to" "${extra[@]}" -- "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then echo "rc=$rc attempt=$attempt"; exit $rc; fi
  sleep 90
done
echo "gave up"; exit 3

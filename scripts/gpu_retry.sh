#!/bin/bash
# scripts/gpu_retry.sh <timeout> '<command>': gpurun, retried every 2 minutes while the pod answers busy (exit 3)
cd /root/repo
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3

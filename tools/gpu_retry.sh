#!/bin/bash
# tools/gpu_retry.sh JOB [TIMEOUT] [GPUS]: run tools/jobs/JOB.sh on a GPU box, retrying while the pod answers busy (exit 3 / "draining")
job=$1; to=${2:-1500}; gpus=${3:-1}
for i in $(seq 1 40); do
  if [ "$gpus" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $to -- "bash tools/jobs/$job.sh" > gpurun_out/${job}_call.log 2>&1; rc=$?
  else /usr/local/graft/bin/gpurun --gpus $gpus --timeout $to -- "bash tools/jobs/$job.sh" > gpurun_out/${job}_call.log 2>&1; rc=$?; fi
  if grep -q "status=ok\|status=fail\|status=timeout" gpurun_out/${job}_call.log; then exit $rc; fi
  sleep 60
done
exit 3

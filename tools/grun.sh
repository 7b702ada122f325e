#!/bin/bash
# build, then run tools/gpu_job.sh on the B200 box; log under gpurun_out/<tag>_call.log
tag=$1; shift
make -C /root/repo -j8 2>&1 | grep -E "error|warning" && { echo "build problem"; exit 1; }
make -C /root/repo 2>&1 | tail -1
gpurun --timeout ${GPU_TIMEOUT:-1500} "$@" -- 'bash tools/gpu_job.sh' > gpurun_out/${tag}_call.log 2>&1
tail -${TAIL:-14} gpurun_out/${tag}_call.log

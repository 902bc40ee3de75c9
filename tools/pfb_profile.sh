#!/bin/bash
# ncu --set full capture of the channelizer kernel (configs[1] workload); usage: gpurun -- 'bash tools/pfb_profile.sh TAG'
tag=$1
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --device-only --no-extra"
$CMD > gpurun_out/${tag}_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pfb -s 3 -c 1 -o gpurun_out/${tag}_pfb $CMD > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log

#!/bin/bash
# ncu --set full capture of the PSK kernel in the c4fm workload; usage: gpurun -- 'bash tools/psk_profile.sh TAG'
tag=$1
CMD="python bench.py --workload c4fm --steps 1 --warmup 3 --no-cpu-baseline --device-only"
$CMD > gpurun_out/${tag}_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:psk -s 3 -c 1 -o gpurun_out/${tag}_psk $CMD > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log

#!/bin/bash
# One ncu --set full capture of a named kernel (first launch after `skip`), exported as raw + source CSV.
#   tools/ncu_kernel.sh <tag> <kernel-regex> <skip> -- <command...>
tag=$1; kern=$2; skip=$3; shift 4
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$kern -s $skip -c 1 -f -o gpurun_out/$tag "$@" > gpurun_out/${tag}_ncu.log 2>&1
ncu -i gpurun_out/$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/$tag.ncu-rep --page source --csv > gpurun_out/${tag}_source.csv 2>/dev/null
ls -la gpurun_out/$tag.ncu-rep | awk '{print $5, $9}'

#!/bin/bash
# ncu --set full of ONE kernel of one step: tools/profile_kernel.sh <tag> <scene> <kernel name regex> [lines]
# (summary + hottest source lines are written on the GPU box; the report stays in /tmp)
tag=$1; scene=$2; kern=$3; lines=${4:-40}
out=gpurun_out
mkdir -p $out
SMENV_STEP_RANGES=1 ncu --profile-from-start off --set full --import-source on --clock-control none -f \
    -k regex:$kern -o /tmp/${tag}_${scene}_${kern} python tools/profile_step.py $scene > $out/${tag}_ncu_${scene}_${kern}.log 2>&1
python tools/ncu_summary.py /tmp/${tag}_${scene}_${kern}.ncu-rep > $out/${tag}_step_${scene}_${kern}.txt 2>&1
python tools/ncu_lines.py /tmp/${tag}_${scene}_${kern}.ncu-rep $kern $lines > $out/${tag}_lines_${scene}_${kern}.txt 2>&1 || true

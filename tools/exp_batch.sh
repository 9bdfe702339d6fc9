#!/bin/bash
out=gpurun_out; tag=${1:-exp}
run() { echo "== $*"; env "$@" python tools/step_time.py $scene; }
{
scene=space_bm; run SMENV_LUT_TINY_MIN=4; run SMENV_STEP_RANGES=3; run SMENV_STEP_RANGES=4
scene=human; run SMENV_LUT_TINY_MIN=4; run SMENV_STEP_RANGES=3; run SMENV_LUT_CONFIG=2
scene=space; run SMENV_LUT_TINY_MIN=4
} 2>&1 | tee $out/${tag}_cmp.txt

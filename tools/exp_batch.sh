#!/bin/bash
# One GPU-box call of an experiment batch: parity tests that exercise the GJK kernel, then step times of table configurations.
out=gpurun_out; tag=${1:-exp}
python -m pytest tests/test_gpu_parity.py tests/test_gpu_human.py -m gpu -q -x -k "distances or golden or from_device_pools or counters or scene_variants or human or replays or independent" 2>&1 | tail -4 > $out/${tag}_tests.txt
cat $out/${tag}_tests.txt
run() { echo "== $*"; env "$@" python tools/step_time.py $scene; }
{
scene=space_bm; run A=1; run SMENV_LUT_BIG_RES=12; run SMENV_LUT_BIG_RES=8; run SMENV_LUT_BUDGET_KB=113; run SMENV_LUT_FINE=0 SMENV_LUT_BIG_RES=8
scene=space; run A=1; run SMENV_LUT_BIG_RES=12; run SMENV_LUT_BUDGET_KB=113; run SMENV_LUT_BUDGET_KB=90
scene=ball; run A=1; run SMENV_LUT_BUDGET_KB=113; run SMENV_LUT_BUDGET_KB=80
scene=human; run A=1; run SMENV_LUT_CONFIG=4; run SMENV_LUT_CONFIG=5
} 2>&1 | tee $out/${tag}_cmp.txt
python tools/gjk_counters.py space_bm human 2>&1 | tee $out/${tag}_counters.txt

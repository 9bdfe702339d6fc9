#!/bin/bash
# The bench lines profiles/ keeps for a build (run WITHOUT ncu):  gpurun --timeout 2400 -- 'bash tools/bench_lines_round2.sh r02'
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
python bench.py > $out/${tag}_bench_default.json 2> $out/${tag}_bench_default.err
python bench.py --scene space_task_bm --risk-gate --no-scenes --no-cpu-baseline > $out/${tag}_bench_gate.json 2> $out/${tag}_bench_gate.err
python bench.py --impl reference --steps 10 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
python tools/e2e_sweep.py human space_bm > $out/${tag}_e2e_sweep.txt 2>&1
python tools/step_time.py space space_bm ball human space_task_bm ball_bm > $out/${tag}_step_time.txt 2>&1
python tools/envs_sweep.py human space > $out/${tag}_envs_sweep.txt 2>&1
tail -c 600 $out/${tag}_bench_default.json; cat $out/${tag}_e2e_sweep.txt $out/${tag}_step_time.txt

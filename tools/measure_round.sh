#!/bin/bash
# One GPU-box call that produces everything profiles/ and DESIGN.md quote for a build:
#   gpurun --timeout 1500 -- 'bash tools/measure_round.sh r01h'
# then here:  python tools/collect_profiles.py r01h
# Bench numbers come from the runs WITHOUT ncu; the ncu passes run afterwards.
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
python bench.py > $out/${tag}_bench_ball.json 2> $out/${tag}_bench_ball.err
python bench.py --scene space --no-cpu-baseline > $out/${tag}_bench_space.json 2> $out/${tag}_bench_space.err
python bench.py --scene space_task_bm --risk-gate --no-cpu-baseline > $out/${tag}_bench_gate.json 2> $out/${tag}_bench_gate.err
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err
# launch list of the bench command itself (cold-cache, serialised per-launch times: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches_ball.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $out/${tag}_ncu_launches.log 2>&1
for scene in ball space; do
  ncu --profile-from-start off --set full --import-source on --clock-control none -f -o $out/${tag}_step_${scene} \
      python tools/profile_step.py $scene > $out/${tag}_ncu_${scene}.log 2>&1
done
ls -la $out | grep ${tag}

#!/bin/bash
# One GPU-box call that produces the ncu evidence profiles/ quotes for a build:
#   gpurun --timeout 2400 -- 'bash tools/measure_round2.sh r02'
# then here:  python tools/collect_profiles.py r02   (summaries -> profiles/)
# Bench numbers come from runs WITHOUT ncu; the ncu passes run afterwards.  gpurun copies back at most 64 MiB: the
# reports are summarised on the box (tools/ncu_summary.py, raw CSV of the metrics collect_profiles.py needs) and deleted.
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
python tools/step_time.py space space_bm ball human > $out/${tag}_step_time.txt 2>&1
# launch list of the bench command itself (cold-cache, serialised per-launch times: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/${tag}_launches_human.csv \
    python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-scenes > $out/${tag}_ncu_launches.log 2>&1
for scene in ${SCENES:-human space_bm space ball}; do
  # one env range per step: every kernel appears once in the report
  SMENV_STEP_RANGES=1 ncu --profile-from-start off --set full --import-source on --clock-control none -f \
      -o /tmp/${tag}_step_${scene} python tools/profile_step.py $scene > $out/${tag}_ncu_${scene}.log 2>&1
  python tools/ncu_summary.py /tmp/${tag}_step_${scene}.ncu-rep > $out/${tag}_step_${scene}.txt 2>&1
  ncu -i /tmp/${tag}_step_${scene}.ncu-rep --page raw --csv \
      --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum > $out/${tag}_raw_${scene}.csv 2>/dev/null
done
# source-level hot spots of the kernels that lead the step (SASS lines with the most stall samples)
for k in human_brake_plan_kernel human_brake_traj_kernel distance_plan_kernel gjk_kernel hcontact_plan_kernel finish_kernel; do
  python tools/ncu_lines.py /tmp/${tag}_step_human.ncu-rep $k 30 > $out/${tag}_lines_human_${k}.txt 2>&1 || true
done
for k in distance_plan_kernel gjk_kernel contact_plan_kernel contact_coarse_kernel joint_solve_kernel; do
  python tools/ncu_lines.py /tmp/${tag}_step_space_bm.ncu-rep $k 30 > $out/${tag}_lines_space_bm_${k}.txt 2>&1 || true
done
ls -la $out | grep ${tag}; ls -la /tmp/*.ncu-rep

#!/bin/bash
# Records of a build in one GPU-box call: bench lines first (WITHOUT ncu), then the ncu evidence.
#   gpurun --timeout 2400 -- 'bash tools/final_round2.sh r02b'   then here: python tools/collect_profiles.py r02b r02b
tag=${1:-r02b}
out=gpurun_out
mkdir -p $out
python bench.py > $out/${tag}_bench_default.json 2> $out/${tag}_bench_default.err
python bench.py --scene space_task_bm --risk-gate --no-scenes --no-cpu-baseline > $out/${tag}_bench_gate.json 2> $out/${tag}_bench_gate.err
python bench.py --impl reference --steps 10 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
tail -c 400 $out/${tag}_bench_default.json
SCENES="human space_bm space" bash tools/measure_round2.sh $tag

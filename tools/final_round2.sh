#!/bin/bash
# Records of a build in one GPU-box call: tests, smoke, bench lines (WITHOUT ncu), then the ncu evidence.
#   gpurun --timeout 2400 -- 'bash tools/final_round2.sh r02c'   then here: python tools/collect_profiles.py r02c r02c
tag=${1:-r02c}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -q 2>&1 | tail -6 > $out/${tag}_tests.txt; cat $out/${tag}_tests.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/${tag}_smoke.txt 2>&1; tail -2 $out/${tag}_smoke.txt
python bench.py > $out/${tag}_bench_default.json 2> $out/${tag}_bench_default.err
python bench.py --scene space_task_bm --risk-gate --no-scenes --no-cpu-baseline > $out/${tag}_bench_gate.json 2> $out/${tag}_bench_gate.err
if [ -n "$WITH_REFERENCE" ]; then
  python bench.py --impl reference --steps 10 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
fi
tail -c 300 $out/${tag}_bench_default.json
SCENES="${SCENES:-human}" bash tools/measure_round2.sh $tag

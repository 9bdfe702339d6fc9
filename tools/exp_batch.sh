#!/bin/bash
out=gpurun_out; tag=${1:-exp}
{
python tools/lib_compare.py human build/exp/base.so build/exp/hcp4.so build/exp/fin5.so build/exp/fin6.so build/exp/hcc4.so
python tools/lib_compare.py space_bm build/exp/base.so build/exp/fin5.so build/exp/fin6.so
} 2>&1 | tee $out/${tag}_cmp.txt

#!/bin/bash
# ncu --set full (+ source) of the kernels matching a regex in one eager B=32 DenseBox step (run under gpurun, 1 GPU).
# usage: tools/prof_kernels.sh <tag> <regex> [count]
tag=$1; re=$2; cnt=${3:-4}
ncu --set full --clock-control none --import-source on -k "regex:$re" -s $cnt -c $cnt -f -o gpurun_out/${tag} \
    python tools/one_step.py densebox 32 2 > gpurun_out/${tag}.log 2>&1
ncu -i gpurun_out/${tag}.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}.ncu-rep --page source --csv > gpurun_out/${tag}_src.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${tag}_raw.csv

#!/bin/bash
# ncu --set full of the HBM-bound kernels of one B=32 DenseBox training step (run under gpurun, 1 GPU).
# usage: tools/prof_elementwise.sh <tag>     -> gpurun_out/<tag>_elem.ncu-rep, <tag>_elem_raw.csv
tag=${1:-r2}
K='regex:maxpool|upsample|im2col|colsum|loss_kernel|sgd_step|transpose_dgrad|fold_pairs|blockdiag'
# matching launches per step: 3 pool_fwd + up_fwd + im2col + loss + 3 pool_bwd + up_bwd + 3 colsum + fold + blockdiag + transpose + sgd = 18
ncu --set full --clock-control none --import-source on -k "$K" -s 18 -c 18 -f -o gpurun_out/${tag}_elem \
    python tools/one_step.py densebox 32 2 > gpurun_out/${tag}_elem.log 2>&1
ncu -i gpurun_out/${tag}_elem.ncu-rep --page raw --csv > gpurun_out/${tag}_elem_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${tag}_elem_raw.csv

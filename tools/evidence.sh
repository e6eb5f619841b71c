#!/bin/bash
# Round evidence on one B200 (run under gpurun): GPU tests, bench line, per-layer step profile, ncu launch list and
# ncu --set full pages of every kernel family.  usage: tools/evidence.sh <tag>   -> gpurun_out/<tag>_*
tag=${1:-r2}
o=gpurun_out
python -m pytest tests -m gpu -q --timeout=1200 > $o/${tag}_tests.log 2>&1; tail -3 $o/${tag}_tests.log
python bench.py > $o/${tag}_bench.json 2> $o/${tag}_bench.err; tail -c 200 $o/${tag}_bench.err
python tools/profile_step.py densebox 32 > $o/${tag}_step.txt 2>&1
python tools/profile_step.py lm 64 > $o/${tag}_step_lm64.txt 2>&1
# launch list of 2 eager steps after 3 warm-up steps (cold-cache, serialised: compare shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $o/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --records main --no-graph --no-cpu-baseline > $o/${tag}_launch_bench.log 2>&1
# ncu --set full, second eager step of tools/one_step.py: tensor kernels (fprop/dgrad/halo: 24 per step, wgrad: 13)
ncu --set full --clock-control none --import-source on -k "regex:conv_fprop|halo" -s 24 -c 24 -f -o $o/${tag}_fprop \
    python tools/one_step.py densebox 32 2 > $o/${tag}_fprop.log 2>&1
ncu -i $o/${tag}_fprop.ncu-rep --page raw --csv > $o/${tag}_fprop_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k "regex:conv_wgrad" -s 13 -c 13 -f -o $o/${tag}_wgrad \
    python tools/one_step.py densebox 32 2 > $o/${tag}_wgrad.log 2>&1
ncu -i $o/${tag}_wgrad.ncu-rep --page raw --csv > $o/${tag}_wgrad_raw.csv 2>/dev/null
# every other kernel of the step (HBM-bound family), both steps captured; the summary uses the second
ncu --set full --clock-control none --import-source on \
    -k "regex:maxpool|upsample|im2col|colsum|loss_kernel|sgd_step|transpose_dgrad|fold_pairs|blockdiag|heads2_dgrad|params_xfer|count_" \
    -c 60 -f -o $o/${tag}_elem python tools/one_step.py densebox 32 2 > $o/${tag}_elem.log 2>&1
ncu -i $o/${tag}_elem.ncu-rep --page raw --csv > $o/${tag}_elem_raw.csv 2>/dev/null
ncu -i $o/${tag}_elem.ncu-rep --page source --csv > $o/${tag}_elem_src.csv 2>/dev/null
rm -f $o/${tag}_fprop.ncu-rep $o/${tag}_wgrad.ncu-rep   # the raw pages are kept; the reports are ~50 MB each
ls -la $o | grep ${tag}_ | awk '{print $5, $9}'

#!/bin/bash
# usage: tools/ab/run_ab.sh "<python tool + args>" lib1 lib2 ...   (each lib: tools/ab/lib_<name>.so)
cmd=$1; shift
for l in "$@"; do echo "=== $l"; DBX_LIB=$PWD/tools/ab/lib_$l.so timeout 200 python $cmd 2>&1 | tail -40; done

#!/bin/bash
# usage: tools/ncu_one.sh <kernel-regex> <out-tag> <cmd...>   (one `ncu --set full` capture of a kernel; plain run first)
pat=$1; tag=$2; shift 2
"$@" > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${tag}_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:${pat} -s 3 -c 2 -f -o gpurun_out/${tag} "$@" > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/${tag}_ncu.log

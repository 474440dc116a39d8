#!/bin/bash
# ncu --set full on selected kernels of one warm step: bash scripts/gpu_ncu.sh TAG REGEX [npart] [skip] [count]
TAG=$1; RE=$2; NP=${3:-4000000}; SK=${4:-0}; CNT=${5:-6}
ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SK -c $CNT -o gpurun_out/prof_$TAG -f \
    python bench.py --steps 1 --warmup 3 --npart-per-gpu $NP --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_$TAG.log 2>&1
tail -n 3 gpurun_out/ncu_full_$TAG.log

#!/bin/bash
# Runs on the GPU box (under gpurun): the ncu launch list and full captures of the heavy kernels at the
# bench size.  Outputs land in gpurun_out/ (summaries are copied into profiles/ afterwards).
#   bash scripts/gpu_profile.sh TAG [npart]
set -x
TAG=${1:-r1}
NP=${2:-16777216}
mkdir -p gpurun_out
# every launch of one warm step with its device time (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 3 --npart-per-gpu $NP --no-cpu-baseline --no-e2e --no-parity-check > gpurun_out/ncu_list_$TAG.log 2>&1
# full captures of the heavy kernels (one launch each, taken in the 3rd step)
ncu --set full --clock-control none --import-source on \
    -k regex:'group_walk|euclid_cull|neigh_lists|h_solve|av_operators|force_cfl|onesweep_pass|leaf_aabb' \
    -s 33 -c 12 -o gpurun_out/prof_$TAG -f \
    python bench.py --steps 1 --warmup 3 --npart-per-gpu $NP --no-cpu-baseline --no-e2e --no-parity-check > gpurun_out/ncu_full_$TAG.log 2>&1
# (11 matching launches per step: 4 sort passes, the AABB pass, walk, cull, lists and the three loops; -s 33 skips
#  three steps, the capture is the fourth)
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$TAG.csv 2> /dev/null
ls -la gpurun_out

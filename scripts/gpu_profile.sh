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
    python bench.py --steps 1 --warmup 3 --npart-per-gpu $NP --no-cpu-baseline --no-e2e > gpurun_out/ncu_list_$TAG.log 2>&1
# full captures of the heavy kernels (one launch each, taken in the 3rd step)
ncu --set full --clock-control none --import-source on \
    -k regex:'group_walk|neigh_lists|h_solve|av_operators|force_cfl|radix_scatter|sort_gather' \
    -s 31 -c 12 -o gpurun_out/prof_$TAG -f \
    python bench.py --steps 1 --warmup 3 --npart-per-gpu $NP --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out

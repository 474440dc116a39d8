#!/bin/bash
# Runs on the GPU box (under gpurun): tests, the bench line, the ncu launch list and full captures
# of the heavy kernels.  Outputs land in gpurun_out/ (copied into profiles/ by hand afterwards).
set -x
TAG=${1:-r1}
NP=${2:-4000000}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -n 40 > gpurun_out/pytest_$TAG.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 3000 gpurun_out/bench_$TAG.json
# every launch of one warm step with its device time (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 3 --npart-per-gpu $NP --no-cpu-baseline --no-e2e > gpurun_out/ncu_list_$TAG.log 2>&1
# full captures of the heavy kernels (one launch each, taken in the 3rd step)
ncu --set full --clock-control none --import-source on \
    -k regex:'leaf_ranges|neigh_mask|neigh_fill|h_solve|av_operators|force_cfl|radix_scatter|sort_gather' \
    -s 27 -c 12 -o gpurun_out/prof_$TAG -f \
    python bench.py --steps 1 --warmup 3 --npart-per-gpu $NP --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out

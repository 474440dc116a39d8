#!/bin/bash
# quick device timing of the step stages (under gpurun): bash scripts/quick.sh [npart] [fp]
NP=${1:-4000000}; FP=${2:-fast}
python bench.py --npart-per-gpu $NP --steps 3 --no-cpu-baseline --no-e2e --fp $FP 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('%.1f Mpart/s  %.2f ms/step  N=%d' % (d['value']/1e6, d['ms_per_step'], d['config']['npart_total']))
        print(' '.join('%s=%.2f' % kv for kv in d['roofline']['stage_ms'].items()))
        print(' '.join('%s:%s=%.2f' % (k, v['bound'], v['frac']) for k, v in d['roofline']['stages'].items()), 'fp64 peak %.1f TF' % d['roofline']['peak_source']['fp64_tflops'], 'copy %.0f GB/s' % d['roofline']['peak_source']['copy_gbs_here'])
    else: print(l, end='')
"

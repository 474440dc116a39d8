#!/bin/bash
# compute-sanitizer over one small step of every code path (under gpurun): bash scripts/sanitize.sh
mkdir -p gpurun_out
cat > /tmp/san.py <<'P'
import sys
sys.path.insert(0, ".")
from tests import scenarios as S
for sc, fp in ((S.periodic_box(3000, "M4", "cd10", jitter=0.1), "fast"), (S.periodic_box(3000, "M6", "mm97", jitter=0.1), "strict"),
               (S.disc(3000, "M4"), "fast"), (S.periodic_box(4000, "M4", "cd10", jitter=0.1, grid=(2, 2, 1)), "fast")):
    m = S.make_cuda(sc, fp_mode=fp, keep_step_data=False)
    for _ in range(2):
        st = m.evolve_once()
    print(sc["name"], fp, st["npart"], st["dt"])
    m.close()
P
for tool in memcheck racecheck; do
  timeout 420 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|periodic_|disc_" gpurun_out/sanitizer_$tool.log | head -20
done

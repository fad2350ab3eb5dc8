#!/bin/bash
# sweep tile widths at a given size; prints per-pass timings
N=${1:-1024}
for cz in 4 8 16; do
  echo "== FB_CZ_COLS=$cz"; FB_CZ_COLS=$cz python tools/ncu_case.py $N noise 3 | tail -2
done
for cz in 8 16 32; do
  echo "== FB_CZ_X=$cz"; FB_CZ_X=$cz python tools/ncu_case.py $N noise 3 | tail -2
done

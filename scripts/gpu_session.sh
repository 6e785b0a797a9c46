#!/bin/bash
P=${1:-r02_s25}
O=gpurun_out
mkdir -p $O
: > $O/${P}_c3full_minctas.log
for m in 0 5 6 8; do
  if [ $m = 0 ]; then D=""; else D="-DCYRK_USER_MINCTAS=$m"; fi
  echo "== CYRK_USER_MINCTAS=$m" >> $O/${P}_c3full_minctas.log
  CYRK_B200_DEBUG=1 CYRK_B200_NVRTC_DEFINES="$D" python scripts/quick_c3full.py 2>&1 | grep -v "events_two_pass=0\|events_static=0" | grep "C3full\|CTAs/SM" | sort -u | tail -4 >> $O/${P}_c3full_minctas.log
done
cat $O/${P}_c3full_minctas.log

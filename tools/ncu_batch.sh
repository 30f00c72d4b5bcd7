#!/bin/bash
# ncu --set full captures of single kernels (one launch each).  The reports embed the whole module (20 MB each) and
# gpurun_out/ is capped at 64 MiB, so the pages are exported as CSV on the box (raw metrics + per-SASS-instruction source page)
# and the reports themselves stay behind, except the ones named in KEEP.
# usage: tools/ncu_batch.sh <tag> [target ...]   targets: blur_h_k7 blur_h_k51 blur_v_k7 blur_v_k51 polar landscape spikey_distant ...
TAG=${1:-r02}; shift
TARGETS=${@:-"blur_h_k7 blur_h_k51 blur_v_k7 blur_v_k51 polar landscape"}
NCU="ncu --set full --import-source on --clock-control none --launch-skip 1 --launch-count 1 -f"
OUT=gpurun_out/ncu; mkdir -p $OUT; TMP=/tmp/ncu_reps; mkdir -p $TMP
for T in $TARGETS; do
  case $T in
    blur_h_k*) K=${T#blur_h_k}; CMD="python tools/blur_one.py h $(python -c "print(($K+0.25)/255)")"; PAT="regex:old_blur";;
    blur_v_k*) K=${T#blur_v_k}; CMD="python tools/blur_one.py v $(python -c "print(($K+0.25)/255)")"; PAT="regex:old_blur";;
    polar) CMD="python tools/effect_one.py ball"; PAT="regex:polar_blit_kernel";;
    landscape) CMD="python tools/effect_one.py landscape"; PAT="regex:landscape_kernel";;
    ball) CMD="python tools/effect_one.py ball"; PAT="regex:ball_kernel";;
    ball_beams) CMD="python tools/effect_one.py ball_beams"; PAT="regex:ball_kernel";;
    tunnel) CMD="python tools/effect_one.py tunnel"; PAT="regex:tunnel_kernel";;
    chain_*) CMD="python tools/demo_one.py ${T#chain_}"; PAT="regex:blend_chain";;
    *) CMD="python tools/effect_one.py $T"; PAT="regex:raymarch_kernel";;
  esac
  $NCU -k $PAT -o $TMP/${TAG}_$T $CMD > $OUT/${TAG}_$T.log 2>&1
  ncu -i $TMP/${TAG}_$T.ncu-rep --page raw --csv > $OUT/${TAG}_${T}_raw.csv 2>> $OUT/${TAG}_$T.log
  ncu -i $TMP/${TAG}_$T.ncu-rep --page source --print-source sass --csv > $OUT/${TAG}_${T}_sass.csv 2>> $OUT/${TAG}_$T.log
  gzip -f $OUT/${TAG}_${T}_sass.csv
done
ls -la $OUT | tail -30

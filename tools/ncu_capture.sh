#!/bin/bash
# ncu captures of the DMFB step kernel at benchmark size (run under gpurun; reports land in gpurun_out/).
# usage: tools/ncu_capture.sh <tag> [c1 c2 c3 c2r c3r ...]; c1 = fused auto-reset as bench.py issues it, c2 / c3 without
# reset, c2r / c3r with fused auto-reset and staggered episodes
set -u
tag=$1; shift
for c in "$@"; do
  mode=""; [ "$c" != "c1" ] && mode="noreset"
  cfg=$c; steps=14; skip=10
  case "$c" in *r) cfg=${c%r}; mode=""; steps=230; skip=220;; esac    # steady state of the run-ahead task search
  python tools/prof_step.py $cfg $steps $mode > gpurun_out/plain_$c.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:dmfb_step_kernel -s $skip -c 2 \
      -f -o gpurun_out/${tag}_step_$c python tools/prof_step.py $cfg $steps $mode > gpurun_out/ncu_$c.log 2>&1
  rc=$?
  # summarise on the box (gpurun brings back at most 64 MiB): text summary kept, report dropped unless KEEP_REP=1
  python tools/ncu_summary.py gpurun_out/${tag}_step_$c.ncu-rep > gpurun_out/${tag}_step_${c}_ncu_full.txt 2>/dev/null
  [ "${KEEP_REP:-0}" = "1" ] || rm -f gpurun_out/${tag}_step_$c.ncu-rep
  echo "$c rc=$rc"
done

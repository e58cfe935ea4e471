#!/bin/bash
# CUDA-event timings of the entry points at 64K envs (run under gpurun) -> gpurun_out/<tag>_kernel_timings.txt
tag=${1:-r02}
out=gpurun_out/${tag}_kernel_timings.txt
{
  echo "# CUDA-event timings at 65,536 envs on one B200 (tools/time_kernels.py, tools/time_meda.py, tools/time_variant.py)."
  echo "# 'one launch per step' = BatchedDMFB(sub_batches=1); 'sub_batches 4' = the batch stepped as 4 pipelined sub-batches"
  for K in 1 4; do
    echo "== DMFB, sub_batches $K"
    for c in c1 c2 c3; do TK_SUB=$K python tools/time_kernels.py $c 65536 $([ $K = 4 ] && echo short); done
  done
  echo "== DMFB C3, fraction of degraded cells 0.3 / 1.0 (one launch per step)"
  for f in 0.3 1.0; do python tools/time_variant.py 50 50 10 9 1 1 65536 $f; done
  for K in $([ -n "${SKIP_MEDA:-}" ] || echo 1 4); do
    echo "== MEDA C4 (30x60, 4 droplets, fov 19), sub_batches $K: obs version, degrade, usage counters, auto-reset"
    TK_SUB=$K python tools/time_meda.py 0 65536 0 0 0
    TK_SUB=$K python tools/time_meda.py 0 65536 0 0 1
    TK_SUB=$K python tools/time_meda.py 2 65536 0 0 0
    TK_SUB=$K python tools/time_meda.py 2 65536 0 0 1
    TK_SUB=$K python tools/time_meda.py 1 65536 0 0 0
    TK_SUB=$K python tools/time_meda.py 0 65536 0 1 0
    TK_SUB=$K python tools/time_meda.py 0 65536 1 1 0
    TK_SUB=$K python tools/time_meda.py 0 65536 1 1 1
  done
} > $out 2>&1
cat $out

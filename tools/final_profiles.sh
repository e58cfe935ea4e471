#!/bin/bash
# Round-end evidence (run under gpurun): launch list of the bench command, steady-state DRAM traffic of the bench
# instance stepped as 4 sub-batches, ncu --set full of the C1 / C2 / C3 step kernels.  Outputs in gpurun_out/.
set -u
tag=${1:-r02}
B="python bench.py --steps 20 --warmup 3 --quick --no-cpu --no-others --no-e2e --windows 1 --window-ms 1"
$B > gpurun_out/${tag}_bench_quick.json 2> gpurun_out/${tag}_bench_quick.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_bench_launches.csv $B > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
TK_SUB=4 python tools/prof_step.py c1 60 > gpurun_out/plain_traffic.log 2>&1 &&
TK_SUB=4 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none \
    -k regex:dmfb_step_kernel -s 56 -c 176 --csv --log-file gpurun_out/${tag}_traffic_c1_sub4.csv python tools/prof_step.py c1 60 > gpurun_out/ncu_traffic.log 2>&1
echo "traffic rc=$?"
tools/ncu_capture.sh ${tag}f c1 c2 c3 c2r

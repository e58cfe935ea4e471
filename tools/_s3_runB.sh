cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_dmfb.py -x -q -m gpu -k "task_search_kernel" 2>&1 | tail -3
python bench.py > gpurun_out/r02g_bench_n1.json 2> gpurun_out/r02g_bench_n1.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r02g_bench_n1.err
python tools/train_vdn.py dmfb --envs 32768 --iters 4 --json gpurun_out/r02g_config5_n1_fp32.json 2>&1 | tail -1 | cut -c1-900
python tools/train_vdn.py dmfb --envs 32768 --iters 4 --bf16-rollout --json gpurun_out/r02g_config5_n1_bf16.json 2>&1 | tail -1 | cut -c1-900
tools/ncu_capture.sh r02g c2r > gpurun_out/r02g_ncu.log 2>&1; echo "ncu rc=$?"
ncu --set full --clock-control none --import-source on -k regex:dmfb_task_search -s 20 -c 2 -f -o gpurun_out/r02g_search_c2 python tools/prof_step.py c2 230 > gpurun_out/ncu_search.log 2>&1; echo "ncu search rc=$?"
python tools/ncu_summary.py gpurun_out/r02g_search_c2.ncu-rep > gpurun_out/r02g_search_c2_ncu_full.txt 2>/dev/null; rm -f gpurun_out/r02g_search_c2.ncu-rep

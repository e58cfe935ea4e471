cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02g_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02g_gputest.log
tail -3 gpurun_out/r02g_gputest.log
SKIP_MEDA=1 bash tools/final_timings.sh r02g
python tools/_s3_c3state.py 2>&1 | tee gpurun_out/r02g_c3_aged_chips.txt

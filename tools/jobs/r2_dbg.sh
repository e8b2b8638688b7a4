mkdir -p gpurun_out
( time PTB_KS_N=100000 PTB_DET_RUNS=2 PTB_LIB_PATH=pytestbeam_b200/_lib/var_dbg.so python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_gputest_dbg.log 2>&1; echo dbg rc=$?; tail -3 gpurun_out/r02_gputest_dbg.log

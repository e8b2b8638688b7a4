mkdir -p gpurun_out
PTB_LIB_PATH=pytestbeam_b200/_lib/var_queue.so python tools/prof_run.py c1 4e6 > /dev/null 2>&1; echo plain rc=$?
PTB_LIB_PATH=pytestbeam_b200/_lib/var_queue.so ncu --set full --clock-control none --import-source on -k regex:"^k_digitise$" --launch-skip 2 -c 1 -f -o gpurun_out/prof_queue python tools/prof_run.py c1 4e6 > gpurun_out/ncu_queue.log 2>&1; echo ncu rc=$?

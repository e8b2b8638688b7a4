mkdir -p gpurun_out
PTB_DIGEST_N=1e6 PTB_LIB_PATH=pytestbeam_b200/_lib/var_pool.so python tools/digest.py > gpurun_out/dig_pool.txt 2>/dev/null; echo pool rc=$?
PTB_DIGEST_N=1e6 PTB_LIB_PATH=pytestbeam_b200/_lib/var_queue.so timeout 300 python tools/digest.py > gpurun_out/dig_queue.txt 2>gpurun_out/dig_queue.err; echo queue rc=$?
diff gpurun_out/dig_pool.txt gpurun_out/dig_queue.txt > gpurun_out/dig_diff.txt && echo DIGESTS IDENTICAL || { echo DIGESTS DIFFER; head -5 gpurun_out/dig_diff.txt; tail -3 gpurun_out/dig_queue.err; }
python tools/time_variants.py tools/stage_time.py pool queue 2>&1 | grep -v Warn
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_compact.py tests/test_gpu_threshold_scan.py tests/test_gpu_errors.py -x -q 2>&1 | tail -5

mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()"
( time python -m pytest tests/test_gpu_production.py -x -q -s -k distributions ) 2>&1 | tail -120 > gpurun_out/r2_ks_1e6.txt; tail -8 gpurun_out/r2_ks_1e6.txt
( time python -m pytest tests/test_gpu_parity.py -x -q -s -k "ten_million" ) 2>&1 | tail -8

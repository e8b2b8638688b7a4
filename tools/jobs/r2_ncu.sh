mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-entry --no-check --reference-tree none --cpu-sample 20000 > gpurun_out/r2_ncu_plain.json 2>/dev/null; echo plain rc=$?
ncu --set full --clock-control none --import-source on -k regex:"k_digitise|k_tracks|k_gather|k_scan" --launch-skip 18 -c 6 -f -o gpurun_out/prof_r02a python bench.py --steps 2 --warmup 1 --no-entry --no-check --reference-tree none --cpu-sample 20000 > gpurun_out/ncu_r02a.log 2>&1; echo ncu rc=$?
ls -la gpurun_out/prof_r02a.ncu-rep

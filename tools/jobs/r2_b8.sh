mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py tests/test_gpu_frontend.py -x -q 2>&1 | tail -3
python tools/entry_profile.py 2e7 2>&1 | grep -v "INFO\|Warn" | tail -7
python bench.py --no-check --cpu-sample 100000 --reference-tree none > gpurun_out/r02d_bench.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r02d_bench.json')); print(d['value'], d['e2e']['value'], d['e2e_entry'])"

python tools/stage_time.py 2>&1 | grep -v Warn
python -m pytest tests/test_gpu_compact.py tests/test_gpu_fullsize.py tests/test_gpu_errors.py -x -q 2>&1 | tail -3
python bench.py --no-entry --cpu-sample 100000 --reference-tree none > gpurun_out/r02e_bench.json 2>gpurun_out/r02e_bench.err; echo bench rc=$?; tail -2 gpurun_out/r02e_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r02e_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['stage_ms_per_step'], d['roofline']['frac'], d['shard_check']['equal'])"

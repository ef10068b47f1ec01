set -x
python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/x3_pytest.log 2>&1
tail -5 gpurun_out/x3_pytest.log
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/x3_bench.json 2> gpurun_out/x3_bench.err
python tools/fuzz.py --help > /dev/null 2>&1
grep -h -o '"ms_per_step": [0-9.]*' gpurun_out/x3_bench.json

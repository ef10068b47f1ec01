#!/bin/bash
ms() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],2), round(d['roofline']['avg_launch_ms'],2), d['config'].get('passes'), d['check'], round(d['e2e']['ms_per_step'],1))"; }
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/record_cost.py 2>&1 >gpurun_out/record_cost.json | sed 's/{.*}//'
python bench.py --no-cpu 2>/dev/null | ms qft_shears
python bench.py --no-cpu --no-shears 2>/dev/null | ms qft_noshears
python bench.py --no-cpu --workload layers 2>/dev/null | ms layers_shears
python bench.py --no-cpu --workload layers --no-shears 2>/dev/null | ms layers_noshears
python bench.py --no-cpu --workload qaoa 2>/dev/null | ms qaoa_shears
python bench.py --no-cpu --workload qaoa --no-shears 2>/dev/null | ms qaoa_noshears

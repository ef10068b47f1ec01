set -x
timeout 900 python -m pytest tests/test_sharded_gpu.py -m gpu -q -x -rs > gpurun_out/r02f_pytest_sharded_n4.log 2>&1
tail -5 gpurun_out/r02f_pytest_sharded_n4.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/r02f_bench_n4.json 2> gpurun_out/r02f_bench_n4.err
tail -c 1500 gpurun_out/r02f_bench_n4.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02f_bench_n2.json 2> gpurun_out/r02f_bench_n2.err

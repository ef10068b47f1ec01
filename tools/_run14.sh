set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/r02f_bench_n8.json 2> gpurun_out/r02f_bench_n8.err
tail -c 600 gpurun_out/r02f_bench_n8.json

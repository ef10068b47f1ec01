for wl in layers qft; do
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-extras --workload $wl --row-bits 3 > gpurun_out/x10_${wl}_c3.json 2> gpurun_out/x10_${wl}_c3.err
echo "$wl c3 $(grep -h -o '"ms_per_step": [0-9.]*' gpurun_out/x10_${wl}_c3.json | head -1) $(grep -h -o '"passes": [0-9]*' gpurun_out/x10_${wl}_c3.json | head -1)"
done
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-extras --workload qaoa --qubits 28 --row-bits 3 > gpurun_out/x10_qaoa_c3.json 2> gpurun_out/x10_qaoa_c3.err
echo "qaoa c3 $(grep -h -o '"ms_per_step": [0-9.]*' gpurun_out/x10_qaoa_c3.json | head -1) $(grep -h -o '"passes": [0-9]*' gpurun_out/x10_qaoa_c3.json | head -1)"
tail -2 gpurun_out/x10_layers_c3.err

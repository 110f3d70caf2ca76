nvidia-smi -L | wc -l
python -m pytest tests/test_gpu_multi.py -x -q -s 2>&1 | tail -4
for N in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -c 1800 gpurun_out/bench_n$N.json | grep -oE '"ms_per_step": [0-9.]+|"exchange": "[a-z_]+"|"value": [0-9.e+]+|"unavailable": "[^"]+"' | tr '\n' ' '; echo; tail -3 gpurun_out/bench_n$N.err | cut -c1-300
done

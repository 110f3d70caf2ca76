set -x
nvidia-smi -L
nvidia-smi topo -m | head -14
[ -n "$SKIPTEST" ] || python -m pytest tests/test_gpu_multi.py -x -q -s 2>&1 | tail -15
N=${N:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02h_bench_n$N.json 2> gpurun_out/r02h_bench_n$N.err; tail -c 5000 gpurun_out/r02h_bench_n$N.json; tail -8 gpurun_out/r02h_bench_n$N.err
